#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched quadrotor hot path on N B200s (one process per GPU).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU port of the reference path

A "step" is one pass of the hot path over one batch of synthetic input: one fused env-step
launch (mixer -> MuJoCo-equivalent dynamics -> obs/reward/done -> Philox auto-reset) over
2^20 hover envs per GPU (BASELINE.json configs[1]).  Envs shard across ranks with no data-path
collective ("scaling": "weak"); NCCL is used only for the barrier and the max-over-ranks time.

Printed JSON (one line, rank 0):  value = whole-job env-steps/s with inputs resident in HBM;
e2e = the same through the host-buffer C-ABI call (H2D actions, D2H obs/reward/done inside the
timed region); roofline = HBM roofline of the step kernel; cpu_baseline = oracle/cpu_ref.c timed on
the host cores; extra keys `resident` (T-step state-resident dyn-only kernel) and `rollout`
(2x128 actor-critic policy rollout + GAE, BASELINE.json configs[2]).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (dyn-only hover step, 2^20 envs/GPU)"
UNIT = "env-steps/s"
NUM_ENVS = 1 << 20


def measured_peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, source): MEASURED_PEAKS.json (driver-written), else B200_PROFILING.md's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1400.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


LAUNCHES_PER_STEP = 16      # one bench "step" = 16 env-step launches: 4 independent state sets x 4 action batches
STATE_SETS = 4              # 4 x 294 MB of state/obs/action traffic rotate, so every launch reads lines the 126 MB L2 no longer holds


def source_fingerprint():
    """sha256 over the STEP kernel's sources (its include closure + the launcher + the ABI header): ties a committed ncu
    traffic measurement to the code that was measured."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "uav_reinforcement_learning_control_b200", "csrc")
    files = [os.path.join(csrc, f) for f in ("qs_dynamics.cuh", "qs_env.cuh", "qs_kernels.cuh", "qs_math.cuh", "qs_pack2.cuh",
                                             "qs_philox.cuh", "qs_step2.cuh", "qs_traj.cuh", "quadsim.cu")]
    for f in files + [os.path.join(ROOT, "include", "quadsim_abi.h")]:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per step-kernel launch from the ncu capture of THIS source tree (tools/measure_traffic.py writes it in
    the same gpurun call as the ncu summary); None when the committed measurement belongs to other sources."""
    tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        with open(tp) as f:
            d = json.load(f)
        if d.get("source_fingerprint") == source_fingerprint():
            return float(d["dram_bytes_per_launch"]), d
    except Exception:
        pass
    return None, None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            for ts, line in self.samples[-3:]:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[0])); mx = float(parts[1])
                except Exception:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(cfg, n_envs=1 << 16, target_seconds=12.0, threads=0):
    """oracle/cpu_ref.c (kind "port") on a bounded sample of the same workload, all host threads."""
    from oracle import cpu_ref
    from uav_reinforcement_learning_control_b200 import model as M
    tree = M.load_mjcf(M.default_model_path())
    hr = cpu_ref.HoverRollout(tree, cfg, n_envs)
    cores = threads or hr.max_threads()
    hr.run(1, seed=0, threads=cores)                       # warm-up + reset
    steps_done, t0 = 0, time.perf_counter()
    while True:
        hr.run(4, seed=steps_done + 1, threads=cores)
        steps_done += 4
        el = time.perf_counter() - t0
        if el >= target_seconds or steps_done >= 512:
            break
    return {"value": n_envs * steps_done / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_envs} hover envs x {steps_done} steps, random actions, auto-reset, float64 "
                      f"MuJoCo-pipeline restatement (oracle/cpu_ref.c), {el:.1f} s"}, el, steps_done


def cpu_baseline_c1(cfg):
    """the reference's own CPU-runnable case, 16 envs x 512 steps, random actions -- single thread and
    all host threads (oracle/cpu_ref.c stands in for the uninstallable MuJoCo / MJX stack)."""
    from oracle import cpu_ref
    from uav_reinforcement_learning_control_b200 import model as M
    tree = M.load_mjcf(M.default_model_path())
    out = {"workload": "16 hover envs x 512 steps, random actions, auto-reset (BASELINE.json configs[0]); float64 CPU port",
           "unit": UNIT, "kind": "port"}
    for key, thr in (("single_thread", 1), ("all_threads", 0)):
        hr = cpu_ref.HoverRollout(tree, cfg, 16)
        cores = thr or min(hr.max_threads(), 16)
        hr.run(8, seed=0, threads=cores)
        reps, t0 = 0, time.perf_counter()
        while True:
            hr.run(512, seed=reps + 1, threads=cores)
            reps += 1
            if time.perf_counter() - t0 > 1.0 or reps >= 64:
                break
        el = time.perf_counter() - t0
        out[key] = {"value": 16 * 512 * reps / el, "cores": cores, "episodes_of_512_steps": reps}
    # the REAL reference paths, if somebody ran tools/dump_reference_vectors.py where mujoco + jax exist and committed the fixture
    gp = os.path.join(ROOT, "tests", "golden", "mujoco_vectors.json")
    if os.path.exists(gp):
        try:
            from tests.golden import vector_io
            t = vector_io.load(gp).get("timing")
            if t:
                out["reference_measured_elsewhere"] = t
        except Exception:
            pass
    return out


def run_reference(args):
    """--impl reference: the CPU port of the reference path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from uav_reinforcement_learning_control_b200 import config as Q
    cfg = Q.EnvConfig.north_star(seed=0)
    n = 1 << 16
    from oracle import cpu_ref
    from uav_reinforcement_learning_control_b200 import model as M
    tree = M.load_mjcf(M.default_model_path())
    hr = cpu_ref.HoverRollout(tree, cfg, n)
    cores = hr.max_threads()
    sub = 4                                                # one "step" here = a bounded sample: n envs x sub env-steps
    for _ in range(max(args.warmup, 1)):
        hr.run(sub, seed=1, threads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        hr.run(sub, seed=2 + k, threads=cores)
    el = time.perf_counter() - t0
    value = n * sub * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "hover env dynamics-only step, 1M envs on 1xB200 (BASELINE.json configs[1]); "
                               "reference arm = CPU port on a bounded sample", "num_envs_per_gpu": NUM_ENVS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} hover envs x {sub} env-steps per step x {args.steps} steps (oracle/cpu_ref.c; "
                                   "mujoco/jax are not installable here, so the port stands in for the reference)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def roofline_policy(tflops_bf16, traj_gbs):
    """Where a tcgen05 policy rollout stands against both rooflines that bound it."""
    hbm, tc, src = measured_peaks()
    return {"tensor": {"achieved": tflops_bf16, "peak": tc, "unit": "TFLOP/s", "frac": tflops_bf16 / tc,
                       "note": "73 984 FLOP of actor + critic forward per env-step / bf16_tflops_sustained"},
            "hbm": {"achieved": traj_gbs, "peak": hbm, "unit": "GB/s", "frac": traj_gbs / hbm,
                    "note": "80 B of trajectory written per env-step / hbm_gbs"},
            "peak_source": src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=NUM_ENVS, help="envs per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip resident / rollout / cpu legs")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.num_envs
    K, W = args.steps, args.warmup
    # global env ids: rank r owns [r*n, (r+1)*n) -> results identical for any sharding
    cfg = Q.EnvConfig.north_star(seed=0, env_id_offset=rank * n)
    eng = Engine(cfg, n, device=local)
    # STATE_SETS independent copies of the whole per-launch working set (state planes, obs, reward, done) and as many action
    # batches: consecutive launches never touch the same lines, so every launch streams from / to HBM (cold L2)
    states = [eng.new_state() for _ in range(STATE_SETS)]
    obss = [torch.empty(n, 12, device=dev) for _ in range(STATE_SETS)]
    rews = [torch.empty(n, device=dev) for _ in range(STATE_SETS)]
    dones = [torch.empty(n, device=dev) for _ in range(STATE_SETS)]
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    acts = [torch.rand(n, 4, device=dev, generator=gen) * 2 - 1 for _ in range(STATE_SETS)]
    for j in range(STATE_SETS):
        eng.reset(states[j], obs=obss[j])
        if j:
            states[j][26] = torch.full((n,), j * 1000, dtype=torch.int32, device=dev).view(torch.float32)   # distinct episode streams
            eng.reset(states[j], obs=obss[j])
    state, obs, rew, done = states[0], obss[0], rews[0], dones[0]
    stream = torch.cuda.current_stream(dev)

    def bench_step(i):
        """one bench step = LAUNCHES_PER_STEP launches rotating over the state sets (every env advances 4 steps)"""
        for l in range(LAUNCHES_PER_STEP):
            j = l % STATE_SETS
            eng.step(states[j], acts[(j + l // STATE_SETS + i) % STATE_SETS], obs=obss[j], reward=rews[j], done=dones[j])

    # ---------------------------------------------------------------- value: HBM-resident API-mode step
    for i in range(W):
        bench_step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record(stream)
    for i in range(K):
        bench_step(i)
    e1.record(stream)
    barrier()
    wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sum_over_ranks(eng.launch_count() - l0)
    # keep the GPU busy a little longer if the timed region was too short for nvidia-smi to see it
    clocks = None
    if sampler:
        if wall1 - wall0 < 0.5:
            tl0 = time.time()
            while time.time() - tl0 < 0.6:
                for i in range(4):
                    bench_step(i)
                torch.cuda.synchronize()
            wall1 = time.time()
        clocks = sampler.stop(wall0, wall1)
    value = world * n * LAUNCHES_PER_STEP * K / (ms * 1e-3)
    finite = bool(all(torch.isfinite(o).all().item() for o in obss))

    # bytes per env-step actually required by the algorithm for this config (DESIGN.md "Traffic"):
    # state words in/out (21 qpos/qvel + 3 target + step_count + episode [+ voltage]) + action 16 + obs 48 + reward 4 + done 4
    words = 21 + 3 + 1 + 1 + (1 if cfg.battery else 0)
    bytes_per = 2 * 4 * words + 16 + 48 + 8
    peak, peak_tc, peak_src = measured_peaks()
    us_per_launch = ms * 1e3 / (K * LAUNCHES_PER_STEP)
    achieved = bytes_per * n / (us_per_launch * 1e-6) / 1e9       # per GPU (max-over-ranks time)
    traffic, traffic_rec = measured_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_steady_state": None if traffic_rec is None else traffic_rec.get("steady_state", {}).get("dram_bytes_per_launch"),
                "frac_dram": None if traffic is None else (traffic / (us_per_launch * 1e-6) / 1e9) / peak,
                "frac_dram_steady_state": None if (traffic_rec is None or "steady_state" not in traffic_rec) else
                (traffic_rec["steady_state"]["dram_bytes_per_launch"] / (us_per_launch * 1e-6) / 1e9) / peak,
                "traffic_note": "traffic = dram__bytes of ONE cold launch (ncu flushes caches first): reads match the algorithmic "
                                "126 MB, writes are short of the algorithmic 168 MB because dirty lines still sit in the 126 MB L2 "
                                "when the kernel ends; traffic_steady_state = the same launches without cache flushes, where every "
                                "launch also absorbs its predecessor's write-backs",
                "kernel": "qs::step2_kernel (two envs per thread, packed f32x2; QS_STEP2=0 selects qs::step_kernel<QS_MODE_HOVER_GYM, FeatLean>)", "bytes_per_env_step": bytes_per,
                "bytes_per_launch": bytes_per * n, "us_per_launch": us_per_launch,
                "traffic_source": None if traffic_rec is None else traffic_rec.get("how"),
                "peak_source": peak_src}

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI
    pin = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory()
    h_act = [(torch.rand(n, 4) * 2 - 1).pin_memory() for _ in range(2)]
    # done comes back as bytes (qs_step_host_bytes: the bool array the reference's step returns), unless the shard is odd
    D2H = 48 + 4 + (1 if n % 4 == 0 else 4)
    h_obs, h_rew = pin(n, 12), pin(n)
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory() if n % 4 == 0 else pin(n)
    Ke = max(5, min(K, 40))
    a_np = [t.numpy() for t in h_act]                     # the caller's arrays (NumPy views of the pinned buffers)
    o_np, r_np, d_np = h_obs.numpy(), h_rew.numpy(), h_done.numpy()
    for i in range(3):
        eng.step_host(state, a_np[i % 2], o_np, r_np, d_np)
    barrier()
    e0.record(stream)
    for i in range(Ke):
        eng.step_host(state, a_np[i % 2], o_np, r_np, d_np)
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    # what bounds it: the D2H of obs / reward / done.  Measure the plain pinned D2H copy rate of this box for the same
    # number of bytes (one cudaMemcpyAsync, all ranks at once, like the e2e loop) so the PCIe fraction is explicit.
    d_probe = torch.empty(D2H * n, dtype=torch.uint8, device=dev)
    h_probe = torch.empty(D2H * n, dtype=torch.uint8).pin_memory()
    h_probe.copy_(d_probe, non_blocking=True)
    barrier()
    e0.record(stream)
    for _ in range(5):
        h_probe.copy_(d_probe, non_blocking=True)
    e1.record(stream)
    barrier()
    ms_probe = max_over_ranks(e0.elapsed_time(e1)) / 5
    pcie_d2h_peak = D2H * n / (ms_probe * 1e-3) / 1e9
    del d_probe, h_probe
    e2e = {"value": world * n * Ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 * n,
           "d2h_bytes_per_step": D2H * n, "steps": Ke, "ms_per_step": ms_e2e / Ke,
           "pcie_d2h_gbs_per_gpu": D2H * n * Ke / (ms_e2e * 1e-3) / 1e9,      # what bounds it: obs/reward/done over PCIe
           "pcie_h2d_gbs_per_gpu": 16 * n * Ke / (ms_e2e * 1e-3) / 1e9,
           "pcie_d2h_copy_peak_gbs_per_gpu": pcie_d2h_peak,                       # measured: same bytes, one plain pinned copy
           "pcie_frac": (D2H * n * Ke / (ms_e2e * 1e-3) / 1e9) / pcie_d2h_peak,
           "api": "qs_step_host_bytes (C ABI, pinned host buffers: float32 actions in, float32 obs / reward and byte done flags out; what HoverVecEnv.step(numpy) calls)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hover env dynamics-only step, 1M envs on 1xB200 (BASELINE.json configs[1]): "
                               "HoverEnv semantics, random actions, Philox auto-reset, one fused launch per step",
                   "num_envs_per_gpu": n, "global_envs": world * n, "mode": "hover_gym/north_star",
                   "step": f"one bench step = {LAUNCHES_PER_STEP} env-step launches ({STATE_SETS} independent state sets x "
                           f"{LAUNCHES_PER_STEP // STATE_SETS} action batches) = {LAUNCHES_PER_STEP} x {n} env-steps per GPU",
                   "launches_per_step": LAUNCHES_PER_STEP,
                   "l2": f"{STATE_SETS} independent working sets of {bytes_per * n / 1e6:.0f} MB rotate between launches "
                         f"({STATE_SETS * bytes_per * n / 1e6:.0f} MB >> 126 MB L2): every launch reads and writes cold lines",
                   "bytes_per_env_step": f"{bytes_per} B = SURVEY 8d's 288 B minus the voltage plane (2 x 4 B), which the "
                                         "north-star config (battery sag off) neither reads nor writes",
                   "parallelism": f"env-shard x{world}, no per-step communication"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "outputs_finite": finite,
    }

    # ---------------------------------------------------------------- extras (rank-local, reported for context)
    if not args.no_extras:
        # (ii) state-resident dyn-only rollout: T steps per launch, in-kernel Philox actions
        T = 64
        stats = torch.zeros(4, n, device=dev)
        eng.rollout_random(state, T, t0=0, stats=stats)
        barrier()
        reps = 3
        e0.record(stream)
        for r in range(reps):
            eng.rollout_random(state, T, t0=(r + 1) * T, stats=stats)
        e1.record(stream)
        barrier()
        ms_r = max_over_ranks(e0.elapsed_time(e1))
        line["resident"] = {"value": world * n * T * reps / (ms_r * 1e-3), "unit": UNIT, "steps_per_launch": T,
                            "episodes_finished": sum_over_ranks(stats[1].sum().item()),
                            "note": "dyn-only, state in registers, obs/reward/done computed every step"}
        # policy rollout, BASELINE.json configs[2]: 8192 envs x 1024 steps, 2x128 ReLU actor-critic + GAE
        nb, Tp = 8192, 1024
        cfg_p = Q.EnvConfig.north_star(seed=1, env_id_offset=rank * nb)
        eng_p = Engine(cfg_p, nb, device=local)
        st_p = eng_p.new_state()
        eng_p.reset(st_p)
        params = make_policy_params(eng_p, torch, dev, seed=0)
        buf = eng_p.rollout_policy(st_p, params, T=Tp, t0=0, dist=0)
        adv, ret = eng_p.gae(buf["reward"], buf["value"], buf["done"], buf["trunc"], buf["last_value"], 0.99, 0.95)
        barrier()
        e0.record(stream)
        buf = eng_p.rollout_policy(st_p, params, T=Tp, t0=Tp, dist=0, buffers=buf)
        eg = torch.cuda.Event(enable_timing=True)
        eg.record(stream)
        eng_p.gae(buf["reward"], buf["value"], buf["done"], buf["trunc"], buf["last_value"], 0.99, 0.95, adv=adv, ret=ret)
        e1.record(stream)
        barrier()
        ms_p = max_over_ranks(e0.elapsed_time(e1))
        ms_roll = e0.elapsed_time(eg)
        flop = 73984.0
        line["rollout"] = {
            "value": world * nb * Tp / (ms_p * 1e-3), "unit": UNIT, "num_envs_per_gpu": nb, "T": Tp,
            "ms_rollout": ms_roll, "ms_gae": ms_p - ms_roll,
            "traj_hbm_gbs": (80.0 * nb * Tp) / (ms_roll * 1e-3) / 1e9,
            "policy_tflops_fp32": flop * nb * Tp / (ms_roll * 1e-3) / 1e12,
            "note": "2x128 ReLU actor + critic (fp32 FMA path), Gaussian sampling, trajectories written once, GAE scan"}
        # the same workload with the policy forward on tcgen05/TMEM (bf16 operands, fp32 accumulate)
        eng_p.rollout_policy(st_p, params, T=Tp, t0=2 * Tp, dist=0, buffers=buf, tensor_cores=True)
        barrier()
        e0.record(stream)
        eng_p.rollout_policy(st_p, params, T=Tp, t0=3 * Tp, dist=0, buffers=buf, tensor_cores=True)
        eg.record(stream)
        eng_p.gae(buf["reward"], buf["value"], buf["done"], buf["trunc"], buf["last_value"], 0.99, 0.95, adv=adv, ret=ret)
        e1.record(stream)
        barrier()
        ms_p = max_over_ranks(e0.elapsed_time(e1)); ms_roll = e0.elapsed_time(eg)
        line["rollout_tc"] = {
            "value": world * nb * Tp / (ms_p * 1e-3), "unit": UNIT, "num_envs_per_gpu": nb, "T": Tp,
            "ms_rollout": ms_roll, "ms_gae": ms_p - ms_roll, "traj_hbm_gbs": (80.0 * nb * Tp) / (ms_roll * 1e-3) / 1e9,
            "policy_tflops_bf16": flop * nb * Tp / (ms_roll * 1e-3) / 1e12,
            "roofline_policy": roofline_policy(flop * nb * Tp / (ms_roll * 1e-3) / 1e12, (80.0 * nb * Tp) / (ms_roll * 1e-3) / 1e9),
            "note": "tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM)"}
        # one full PPO iteration of configs[2] on the device -- tcgen05 rollout (8192 x 1024) -> GAE ->
        # 4 epochs x 8 minibatches of 2^20 samples through qs_ppo_grad / qs_ppo_adam (one flat NCCL all-reduce of the
        # 37 033-float gradient + statistics per minibatch when N > 1 -- or, by default, no collective at all: the gradients meet in
        # NVLink peer memory inside the optimiser kernel, qs_ppo_adam_peer); train.py:50-68 hyper-parameters
        from uav_reinforcement_learning_control_b200.parallel import DistContext
        from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer
        torch.manual_seed(1234)
        tr = PPOTrainer(eng_p, PPOConfig(n_steps=Tp, n_epochs=4, num_minibatches=8), ctx=DistContext(rank, world, local, None), seed=0,
                        peer=True)
        tr.collect(); tr.update()                                   # warm-up iteration (allocations, NCCL channels)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        tr.collect()
        ev[1].record(stream)
        tr.update()
        ev[2].record(stream)
        barrier()
        ms_it = max_over_ranks(ev[0].elapsed_time(ev[2])); ms_col = ev[0].elapsed_time(ev[1])
        n_mb = 4 * 8; mb = nb * Tp // 8
        line["train_iter"] = {
            "value": world * nb * Tp / (ms_it * 1e-3), "unit": "env-steps/s incl. PPO update", "num_envs_per_gpu": nb, "T": Tp,
            "ms_rollout_gae": ms_col, "ms_update": ms_it - ms_col, "minibatches": n_mb, "samples_per_minibatch": mb,
            "ms_per_minibatch": (ms_it - ms_col) / n_mb,
            "update_samples_per_s": world * mb * n_mb / ((ms_it - ms_col) * 1e-3),
            "update_tflops_bf16": 220672.0 * mb * n_mb / ((ms_it - ms_col) * 1e-3) / 1e12,
            "gradient_exchange": "none (1 GPU)" if world == 1 else ("NVLink peer memory inside the optimiser kernel (qs_ppo_adam_peer)"
                                                                    if tr.updater.comm is not None else "NCCL all-reduce"),
            "note": "qs_rollout_policy (tcgen05) + qs_gae + 32 x {qs_ppo_grad (tcgen05 forward + backward, weight gradients "
                    "in TMEM) + qs_ppo_adam}, issued as one qs_ppo_update_epoch call per epoch; includes qs_ppo_pack, the per-epoch "
                    "qs_ppo_permutation shuffles (computed on a side stream next to the rollout) and the statistics read-back"}
        tr.updater.close()
        del tr
        del eng_p
        # the same config in MJX-parity mode: JaxMJXQuadBraxEnv (21-D obs, Episode + AutoReset wrappers), tanh-normal
        # policy, brax GAE -- fp32 FMA path and tcgen05 path
        cfg_b = Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST, seed=3, env_id_offset=rank * nb)
        eng_b = Engine(cfg_b, nb, device=local)
        st_b = eng_b.new_state(); first_b = torch.zeros(21, nb, device=dev)
        eng_b.reset(st_b, first_state=first_b)
        params_b = make_policy_params(eng_b, torch, dev, seed=0, dist=1)
        Tb = 256
        for key, tcf in (("rollout_mjx_brax", False), ("rollout_mjx_brax_tc", True)):
            bufb = eng_b.rollout_policy(st_b, params_b, T=Tb, t0=0, dist=1, first_state=first_b, tensor_cores=tcf)
            advb, retb = eng_b.gae(bufb["reward"], bufb["value"], bufb["done"], bufb["trunc"], bufb["last_value"], 0.99, 0.95, brax_form=True)
            barrier()
            e0.record(stream)
            eng_b.rollout_policy(st_b, params_b, T=Tb, t0=Tb, dist=1, first_state=first_b, buffers=bufb, tensor_cores=tcf)
            eng_b.gae(bufb["reward"], bufb["value"], bufb["done"], bufb["trunc"], bufb["last_value"], 0.99, 0.95, brax_form=True,
                      adv=advb, ret=retb)
            e1.record(stream)
            barrier()
            ms_b = max_over_ranks(e0.elapsed_time(e1))
            line[key] = {"value": world * nb * Tb / (ms_b * 1e-3), "unit": UNIT, "num_envs_per_gpu": nb, "T": Tb,
                         "note": "JaxMJXQuadBraxEnv semantics, 21-D obs, tanh-normal 2x128 actor-critic, brax GAE"}
        # SURVEY 8f N4, Brax side: the same env trained with brax's recipe on the device -- tcgen05 rollout -> brax GAE -> running
        # observation normaliser -> fused tanh-normal PPO update (qs_ppo_grad for the 21-D policy + qs_ppo_adam).  Two
        # geometries: the reference's own (train_brax_ppo.py:431-455: unroll 10, 16 minibatches x 4 updates, batch_size x
        # num_minibatches = 16384 envs) and a throughput one (8192 envs x 256 steps, 8 minibatches x 4 epochs)
        line["train_iter_mjx_brax"] = {}
        for key, nenv, cfg_t in (("reference_geometry", 16384, PPOConfig.brax_reference()),
                                 ("throughput_geometry", nb, PPOConfig.brax_reference(n_steps=256, num_minibatches=8))):
            eng_t = Engine(Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST, seed=3, env_id_offset=rank * nenv),
                           nenv, device=local)
            trb = PPOTrainer(eng_t, cfg_t, ctx=DistContext(rank, world, local, None), seed=0, peer=True)
            trb.collect(); trb.update()
            barrier()
            reps = 4 if key == "reference_geometry" else 1
            ev[0].record(stream)
            for _ in range(reps):
                trb.collect()
                trb.update()
            ev[2].record(stream)
            barrier()
            ms_t = max_over_ranks(ev[0].elapsed_time(ev[2])) / reps
            line["train_iter_mjx_brax"][key] = {
                "value": world * nenv * cfg_t.n_steps / (ms_t * 1e-3), "unit": "env-steps/s incl. PPO update", "num_envs_per_gpu": nenv,
                "T": cfg_t.n_steps, "ms_per_iteration": ms_t, "minibatches": cfg_t.n_epochs * cfg_t.num_minibatches,
                "samples_per_minibatch": nenv * cfg_t.n_steps // cfg_t.num_minibatches, "fused": trb.fused,
                "params_finite": bool(torch.isfinite(trb.params).all().item())}
            trb.updater.close()
            del trb, eng_t
        del eng_b, bufb
        # big-batch policy rollout (per-GPU shard of configs[4]): 2^18 envs x 32 steps
        nb2, T2 = 1 << 18, 32
        cfg_q = Q.EnvConfig.north_star(seed=2, env_id_offset=rank * nb2)
        eng_q = Engine(cfg_q, nb2, device=local)
        st_q = eng_q.new_state()
        eng_q.reset(st_q)
        bufq = None
        for key, tcf in (("rollout_large", False), ("rollout_large_tc", True)):
            bufq = eng_q.rollout_policy(st_q, params, T=T2, t0=0, dist=0, buffers=bufq, tensor_cores=tcf)
            barrier()
            e0.record(stream)
            eng_q.rollout_policy(st_q, params, T=T2, t0=T2, dist=0, buffers=bufq, tensor_cores=tcf)
            e1.record(stream)
            barrier()
            ms_q = max_over_ranks(e0.elapsed_time(e1))
            line[key] = {"value": world * nb2 * T2 / (ms_q * 1e-3), "unit": UNIT, "num_envs_per_gpu": nb2, "T": T2,
                         ("policy_tflops_bf16" if tcf else "policy_tflops_fp32"): flop * nb2 * T2 / (ms_q * 1e-3) / 1e12,
                         "traj_hbm_gbs": (80.0 * nb2 * T2) / (ms_q * 1e-3) / 1e9}
            if tcf:
                line[key]["roofline_policy"] = roofline_policy(flop * nb2 * T2 / (ms_q * 1e-3) / 1e12, (80.0 * nb2 * T2) / (ms_q * 1e-3) / 1e9)
        del eng_q, bufq
        # BASELINE.json configs[3]: circle / figure-8 / square waypoint tracking (utils/trajectories.py tables, evaluate.py's
        # advance rule and lap reset fused in the step), 262 144 envs, policy rollout on tcgen05
        from uav_reinforcement_learning_control_b200 import policies, trajectories as TJ
        cfg_w = Q.EnvConfig.waypoint_eval(TJ.default_tables(0.5), auto_reset=Q.RESET_RESAMPLE, seed=4, env_id_offset=rank * nb2)
        eng_w = Engine(cfg_w, nb2, device=local)
        st_w = eng_w.new_state()
        eng_w.reset(st_w)
        # flown by a scripted cascaded-PD policy written into the 2x128 ReLU network's weights (policies.pd_waypoint_policy):
        # the timed leg really advances waypoints (a random-init policy leaves the bounds before reaching the first one)
        params_w = torch.from_numpy(policies.pd_waypoint_policy(log_std=-3.5)).to(dev)
        Tw = 128
        bufw = eng_w.rollout_policy(st_w, params_w, T=Tw, t0=0, dist=0, tensor_cores=True)
        barrier()
        reached0 = sum_over_ranks(st_w[29].view(torch.int32).sum().item())
        e0.record(stream)
        eng_w.rollout_policy(st_w, params_w, T=Tw, t0=Tw, dist=0, buffers=bufw, tensor_cores=True)
        e1.record(stream)
        barrier()
        ms_w = max_over_ranks(e0.elapsed_time(e1))
        reached1 = sum_over_ranks(st_w[29].view(torch.int32).sum().item())
        line["rollout_waypoint_tc"] = {"value": world * nb2 * Tw / (ms_w * 1e-3), "unit": UNIT, "num_envs_per_gpu": nb2, "T": Tw,
                                       "waypoints_reached": reached1, "waypoints_reached_in_timed_launch": reached1 - reached0,
                                       "envs_out_of_bounds_in_timed_launch": sum_over_ranks(bufw["done"].sum().item()),
                                       "policy_tflops_bf16": flop * nb2 * Tw / (ms_w * 1e-3) / 1e12,
                                       "note": "waypoint tables circle / eight / square (13 / 13 / 12 points), reach radius 0.25, battery sag on; "
                                               "scripted PD policy in the network weights, sigma = exp(-3.5) exploration noise"}
        del eng_w, bufw
        # BASELINE.json configs[4], this GPU's shard: 2^20 envs x 128 steps per PPO iteration (8 M envs across 8 GPUs), the
        # 37 033-float gradient all-reduce per minibatch being the only NCCL traffic
        nb3, T3 = 1 << 20, 128
        eng_s = Engine(Q.EnvConfig.north_star(seed=5, env_id_offset=rank * nb3), nb3, device=local)
        trs = PPOTrainer(eng_s, PPOConfig(n_steps=T3, n_epochs=4, num_minibatches=8), ctx=DistContext(rank, world, local, None), seed=0,
                         peer=True)
        trs.collect(); trs.update()
        barrier()
        ev[0].record(stream)
        trs.collect()
        ev[1].record(stream)
        trs.update()
        ev[2].record(stream)
        barrier()
        ms_s = max_over_ranks(ev[0].elapsed_time(ev[2])); ms_sc = ev[0].elapsed_time(ev[1])
        line["sharded_ppo"] = {"value": world * nb3 * T3 / (ms_s * 1e-3), "unit": "env-steps/s incl. PPO update",
                               "num_envs_per_gpu": nb3, "global_envs": world * nb3, "T": T3, "ms_rollout_gae": ms_sc,
                               "ms_update": ms_s - ms_sc, "samples_per_minibatch": nb3 * T3 // 8,
                               "update_samples_per_s": world * nb3 * T3 * 4 / ((ms_s - ms_sc) * 1e-3),
                               "collectives_per_iteration": 32 if (world > 1 and trs.updater.comm is None) else 0,
                               "gradient_exchange": "none (1 GPU)" if world == 1 else (
                                   "NVLink peer memory inside the optimiser kernel (qs_ppo_adam_peer)" if trs.updater.comm is not None
                                   else "NCCL all-reduce")}
        # do the ranks really hold the same policy after the timed update?  (all-gather AFTER the timed region)
        pbits = trs.params.view(torch.int32)
        if world > 1:
            gathered = [torch.empty_like(pbits) for _ in range(world)]
            dist.all_gather(gathered, pbits)
            in_sync = all(bool(torch.equal(g, gathered[0])) for g in gathered)
        else:
            in_sync = True
        line["sharded_ppo"]["params_bitwise_in_sync"] = in_sync
        line["sharded_ppo"]["params_finite"] = bool(torch.isfinite(trs.params).all().item())
        trs.updater.close()
        del trs, eng_s
        torch.cuda.empty_cache()
        if rank == 0:
            cb, _, _ = cpu_baseline(Q.EnvConfig.north_star(seed=0), target_seconds=args.cpu_seconds)
            line["cpu_baseline"] = cb
            line["cpu_baseline_c1"] = cpu_baseline_c1(Q.EnvConfig.north_star(seed=0))
    if rank == 0 and "cpu_baseline" not in line:
        cb, _, _ = cpu_baseline(Q.EnvConfig.north_star(seed=0), n_envs=1 << 14, target_seconds=3.0)
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def make_policy_params(eng, torch, dev, seed=0, dist=0):
    """Random-init 2x128 ReLU actor-critic in the packed layout of qs_policy_param_count (lecun-uniform)."""
    D, H, A = eng.obs_dim, 128, 4
    Ao = 2 * A if dist == 1 else A
    g = torch.Generator(device="cpu"); g.manual_seed(seed)

    def lin(i, o):
        lim = (3.0 / i) ** 0.5
        return (torch.rand(i, o, generator=g) * 2 - 1) * lim, torch.zeros(o)
    parts = []
    for out in (Ao, 1):
        for (i, o) in ((D, H), (H, H), (H, out)):
            w, b = lin(i, o)
            parts += [w.reshape(-1), b]
    if dist == 0:
        parts.append(torch.full((A,), -0.5))          # log_std
    parts += [torch.zeros(D), torch.ones(D)]           # obs mean, inv std
    p = torch.cat(parts).to(dev)
    assert p.numel() == eng.policy_param_count(dist), (p.numel(), eng.policy_param_count(dist))
    return p.contiguous()


if __name__ == "__main__":
    main()
