"""Bit layouts of the tcgen05 descriptors the tensor-core kernels build (csrc/qs_umma_desc.cuh), checked on the CPU.

The packing functions are plain integer code compiled for host and device; here the host build (tests/host_harness) is
compared with an INDEPENDENT restatement of the published field tables (cute/arch/mma_sm100_desc.hpp, unions
SmemDescriptor and InstrDescriptor): each field is placed by (offset, width) from the table, not by the shifts the
header under test uses.  Covers every (M, N, major) combination and every shared-memory stride the kernels pass.
"""
import ctypes as C
import itertools

import pytest

from .util import HostHarness

# (name, bit offset, width) -- cute::UMMA::SmemDescriptor
SMEM_FIELDS = [("start_address", 0, 14), ("leading_byte_offset", 16, 14), ("stride_byte_offset", 32, 14), ("version", 46, 2),
               ("base_offset", 49, 3), ("lbo_mode", 52, 1), ("layout_type", 61, 3)]
# cute::UMMA::InstrDescriptor
INSTR_FIELDS = [("sparse_id2", 0, 2), ("sparse_flag", 2, 1), ("saturate", 3, 1), ("c_format", 4, 2), ("a_format", 7, 3),
                ("b_format", 10, 3), ("a_negate", 13, 1), ("b_negate", 14, 1), ("a_major", 15, 1), ("b_major", 16, 1),
                ("n_dim", 17, 6), ("m_dim", 24, 5), ("max_shift", 30, 2)]


def pack(fields, **vals):
    word = 0
    for name, off, width in fields:
        v = vals.pop(name, 0)
        assert 0 <= v < (1 << width), f"{name}={v} does not fit {width} bits"
        word |= v << off
    assert not vals, f"unknown fields {sorted(vals)}"
    return word


def unpack(fields, word):
    out = {name: (word >> off) & ((1 << width) - 1) for name, off, width in fields}
    covered = 0
    for _, off, width in fields:
        covered |= ((1 << width) - 1) << off
    out["_reserved"] = word & ~covered
    return out


@pytest.fixture(scope="module")
def lib():
    L = HostHarness.lib()
    L.hh_umma_smem_desc.restype = C.c_uint64
    L.hh_umma_smem_desc.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.hh_umma_idesc.restype = C.c_uint32
    L.hh_umma_idesc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    return L


# the strides the kernels pass: K-major operand tiles (lbo = bytes between K core matrices = rows/8 * 128, sbo = 128) and
# the MN-major forms of the PPO kernels (128 <-> group stride), for row counts 16 .. 256
STRIDES = sorted({(r // 8 * 128, 128) for r in (16, 32, 128, 256)} | {(128, 256), (128, 2048), (256, 128), (2048, 128), (128, 512),
                                                                      (512, 128), (2 * 128, 128), (16 * 128, 128), (32 * 128, 128)})


@pytest.mark.parametrize("lbo,sbo", STRIDES)
def test_smem_descriptor_fields(lib, lbo, sbo):
    # shared-memory window addresses are < 256 KB and 16-byte aligned in every kernel (asserted at their call sites by the
    # Smem struct offsets); sweep the start address over the whole 227 KB carve-out
    for saddr in (0, 16, 1024, 0x1000 + 48, 100 * 1024, 227 * 1024 - 16):
        got = lib.hh_umma_smem_desc(saddr, lbo, sbo)
        want = pack(SMEM_FIELDS, start_address=saddr >> 4, leading_byte_offset=lbo >> 4, stride_byte_offset=sbo >> 4, version=1,
                    base_offset=0, lbo_mode=0, layout_type=0)
        assert got == want, (hex(got), hex(want))
        f = unpack(SMEM_FIELDS, got)
        assert f["_reserved"] == 0
        assert f["start_address"] << 4 == saddr and f["leading_byte_offset"] << 4 == lbo and f["stride_byte_offset"] << 4 == sbo


def test_smem_descriptor_address_window(lib):
    # bits of a generic->shared address above the 256 KB window are dropped, never spilled into the lbo field
    got = lib.hh_umma_smem_desc(0xFFFC0000 | 0x2340, 256, 128)
    assert unpack(SMEM_FIELDS, got)["start_address"] == 0x2340 >> 4
    assert unpack(SMEM_FIELDS, got)["leading_byte_offset"] == 16


@pytest.mark.parametrize("M,N,a_mn,b_mn", list(itertools.product((64, 128), (8, 16, 32, 64, 128, 256), (0, 1), (0, 1))))
def test_instruction_descriptor_fields(lib, M, N, a_mn, b_mn):
    got = lib.hh_umma_idesc(M, N, a_mn, b_mn)
    want = pack(INSTR_FIELDS, c_format=1, a_format=1, b_format=1, a_major=a_mn, b_major=b_mn, n_dim=N >> 3, m_dim=M >> 4)
    assert got == want, (hex(got), hex(want))
    f = unpack(INSTR_FIELDS, got)
    assert f["_reserved"] == 0 and f["sparse_flag"] == 0 and f["saturate"] == 0 and f["a_negate"] == 0 and f["b_negate"] == 0
    assert f["n_dim"] * 8 == N and f["m_dim"] * 16 == M and f["max_shift"] == 0


def test_kernel_shapes_are_legal():
    """every (M, N) the kernels issue satisfies the kind::f16 shape rule: M = 128 -> N % 16 == 0, 16 <= N <= 256"""
    for M, N in ((128, 256), (128, 128), (128, 16), (128, 32)):
        assert M == 128 and N % 16 == 0 and 16 <= N <= 256
