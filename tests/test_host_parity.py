"""CPU suite: the kernels' per-env source (g++ host build) against the oracle.

This is what keeps the closed-form dynamics and env semantics honest on the GPU-less dev box;
tests/test_gpu_parity.py repeats every case through the real library on the B200.
"""
import pytest

from . import parity_cases as pc
from .util import HostHarness


@pytest.mark.parametrize("name", list(pc.CONFIGS))
def test_single_step_vs_oracle(name):
    stats = pc.check_single_step(HostHarness, name, n=2048, seed=11)
    if name != "mjx_playground":        # (that env never terminates: jax_mjx_quad_env.py:170-172)
        assert stats["done"] > 0        # the synthetic states do straddle the bounds
    else:
        assert stats["truncated"] > 0


@pytest.mark.parametrize("name", ["hover_gym", "traj_gym", "mjx_brax", "hover_brax", "mjx_playground"])
def test_reset_vs_oracle(name):
    pc.check_reset(HostHarness, name, n=1024)


@pytest.mark.parametrize("name", ["hover_gym", "traj_gym", "mjx_brax", "hover_brax"])
def test_masks_bit_exact_on_injected_state(name):
    assert pc.check_observe_bit_exact_masks(HostHarness, name, n=4096) > 0


@pytest.mark.parametrize("auto_reset", [True, False])
def test_waypoint_advance_bit_exact(auto_reset):
    pc.check_waypoints(HostHarness, n=300, steps=5, auto_reset=auto_reset)


def test_waypoint_lap_completion():
    pc.check_lap_completion(HostHarness)


def test_traj_spline_info_vs_scipy():
    """TrajectoryFollowEnv spline reference (envs/trajectory_follow_env.py:176-218): the kernels' closed form
    (csrc/qs_traj.cuh, host build) against scipy.interpolate.CubicSpline(bc_type='natural') on the same draws."""
    pc.check_traj_info(HostHarness, n=256)
