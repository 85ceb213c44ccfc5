"""Golden vectors produced by the reference's own importable utilities (tests/golden/make_golden.py)
pin: the waypoint tables, the observation normalisation, the action denormalisation, the
quaternion <-> Euler convention, and the drone constants -- for the oracle AND for the kernel source."""
import json
import os

import numpy as np
import pytest

from oracle.envs import OracleEnv
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M
from uav_reinforcement_learning_control_b200 import trajectories as TJ

from .util import HostHarness, make_planes

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_utils.json")))


def test_drone_config_constants():
    d = G["drone_config"]
    assert (d["MAX_MOTOR_THRUST"], d["ARM_LENGTH"], d["YAW_TORQUE_COEFF"], d["MAX_TORQUE"]) == (
        Q.MAX_MOTOR_THRUST, Q.ARM_LENGTH, Q.YAW_TORQUE_COEFF, Q.MAX_TORQUE)
    assert d["MAX_TOTAL_THRUST"] == Q.MAX_TOTAL_THRUST and d["DT"] == M.load_default()[1].dt


@pytest.mark.parametrize("key", sorted(G["waypoints"]))
def test_waypoint_tables(key):
    want = np.array(G["waypoints"][key])
    name, rest = key.split("@")
    if "," in rest:
        got = TJ.circle(spacing=0.5, radius=0.7, center=(0.2, -0.1, 1.3))
    else:
        got = TJ.GENERATORS[name](spacing=float(rest))
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_waypoint_counts_and_first_points():
    e, c, s = TJ.default_tables(0.5)
    assert (len(e), len(c), len(s)) == (13, 13, 12)
    np.testing.assert_allclose(e[0], [1, 0, 1]); np.testing.assert_allclose(c[0], [1, 0, 1])
    np.testing.assert_allclose(s[0], [0.75, 0.75, 1])


def test_observation_normalisation_oracle_and_kernel_source():
    x = np.array(G["normalize"]["x"], dtype=np.float32); y = np.array(G["normalize"]["y"], dtype=np.float32)
    tree = M.load_mjcf(M.default_model_path())
    cfg = Q.EnvConfig.hover_gym()
    orc = OracleEnv(tree, cfg)
    got = orc.obs_gym(np.concatenate([-x[:, 0:3], x[:, 3:]], axis=1), np.zeros((len(x), 3), np.float32))
    np.testing.assert_array_equal(got, y)                      # the oracle follows the reference's float32 op order
    np.testing.assert_allclose(y[0], -1.0, atol=1e-6); np.testing.assert_allclose(y[1], 1.0, atol=1e-6)
    # kernel source: inject pos = -relpos (target 0), attitude via quaternion, velocities as is
    from scipy.spatial.transform import Rotation
    sel = (np.abs(x[:, 4]) < 1.5) & (np.abs(x[:, 3]) < 3.1) & (np.abs(x[:, 5]) < 3.1)    # representable Euler angles
    xs = x[sel]
    quat = Rotation.from_euler("xyz", xs[:, 3:6].astype(np.float64)).as_quat()[:, [3, 0, 1, 2]]
    qpos = np.zeros((len(xs), 11)); qpos[:, 0:3] = -xs[:, 0:3]; qpos[:, 3:7] = quat
    qvel = np.zeros((len(xs), 10)); qvel[:, 0:6] = xs[:, 6:12]
    st = make_planes(len(xs), qpos, qvel, target=np.zeros((len(xs), 3)))
    obs, _, _ = HostHarness(cfg).observe(st)
    np.testing.assert_allclose(obs, y[sel], rtol=0, atol=3e-6)


def test_action_denormalisation():
    a = np.array(G["denormalize"]["a"], dtype=np.float32); u = np.array(G["denormalize"]["u"])
    tree = M.load_mjcf(M.default_model_path())
    orc = OracleEnv(tree, Q.EnvConfig.hover_gym(battery=False))
    F, _ = orc.action_to_ctrl(a, np.full(len(a), 8.4))
    want = np.clip(u @ Q.EnvConfig().mixer()[1].T, 0.0, 13.0)
    np.testing.assert_allclose(F, want, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(u[0], [26.0, 0, 0, 0], atol=1e-6)      # action 0 -> 26 N (MIXING_MATRIX_CONTROL.md:49)


def test_euler_convention_matches_reference_state_class():
    d = G["set_from_mujoco"]
    qpos = np.array(d["qpos"]); qvel = np.array(d["qvel"]); want = np.array(d["state12"], dtype=np.float32)
    q11 = np.zeros((len(qpos), 11)); q11[:, :7] = qpos
    v10 = np.zeros((len(qpos), 10)); v10[:, :6] = qvel
    got = OracleEnv.state12(q11, v10)
    np.testing.assert_array_equal(got, want)
    # kernel source: normalised obs -> physical state12
    cfg = Q.EnvConfig.hover_gym()
    st = make_planes(len(qpos), q11, v10, target=np.zeros((len(qpos), 3)))
    obs, _, _ = HostHarness(cfg).observe(st)
    lo = np.float32(cfg.obs_lo); hi = np.float32(cfg.obs_hi)
    phys = (obs.astype(np.float64) + 1) / 2 * (hi - lo) + lo
    np.testing.assert_allclose(phys[:, 3:6], want[:, 3:6], atol=5e-6)
    g = G["get_mujoco_state"]
    rpy = np.array(g["rpy"], dtype=np.float32); quat = np.array(g["quat_wxyz"])
    # the engine's Euler->quaternion (reset path) must reproduce scipy's from_euler('xyz')
    from scipy.spatial.transform import Rotation
    mine = Rotation.from_euler("xyz", rpy.astype(np.float64)).as_quat()[:, [3, 0, 1, 2]]
    np.testing.assert_allclose(mine, quat, atol=1e-12)
