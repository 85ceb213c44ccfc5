"""Self-checks that pin the float64 oracle to everything that CAN be pinned without MuJoCo:
analytic known answers, conservation laws, symmetry of the closed-form assumptions, C == NumPy."""
import numpy as np
import pytest

from oracle import cpu_ref
from oracle.mujoco_pipeline import TreePipeline
from uav_reinforcement_learning_control_b200 import model as M

from .util import random_states


@pytest.fixture(scope="module")
def tree():
    return M.load_mjcf(M.default_model_path())


def _qacc(P, q, v, ctrl):
    f, Mm = P.forward(q, v, ctrl)
    return np.linalg.solve(Mm, f[..., None])[..., 0]


def test_known_accelerations(tree):
    """SURVEY 8a-11 scratch-verified answers."""
    P = TreePipeline(tree, fluid=False, gravity=False)
    q = np.zeros((1, 11)); q[0, 3] = 1; v = np.zeros((1, 10))
    _, Mm = P.forward(q, v, np.zeros((1, 4)))
    tau = np.zeros(10); tau[5] = 1e-3
    a = np.linalg.solve(Mm[0], tau)
    assert a[5] == pytest.approx(1.600277018, rel=1e-9)
    np.testing.assert_allclose(a[6:], -1.600277018, rtol=1e-9)       # rotors stay inertially fixed
    tau = np.zeros(10); tau[3] = 1e-3
    a = np.linalg.solve(Mm[0], tau)
    assert a[3] == pytest.approx(2.027267355, rel=1e-9) and a[1] == pytest.approx(8.231840e-3, rel=1e-6)
    Pg = TreePipeline(tree, fluid=False, gravity=True)
    np.testing.assert_allclose(_qacc(Pg, q, v, np.zeros((1, 4)))[0], [0, 0, -9.81] + [0] * 7, atol=1e-12)


def test_hover_thrust_balances_gravity(tree):
    _, c = M.load_default()
    P = TreePipeline(tree, fluid=True)
    q = np.zeros((1, 11)); q[0, 3] = 1; v = np.zeros((1, 10))
    a = _qacc(P, q, v, np.full((1, 4), c.hover_thrust_per_motor))
    np.testing.assert_allclose(a[0, :3], 0, atol=1e-9)
    # equal thrusts act through the site centroid, 4.06 mm below the COM offset -> no torque
    np.testing.assert_allclose(a[0, 3:6], 0, atol=1e-9)


def test_action_zero_gives_26N(tree):
    """MIXING_MATRIX_CONTROL.md:49-50: action 0 -> 26 N total, 6.5 N per motor."""
    from uav_reinforcement_learning_control_b200 import config as Q
    from oracle.envs import OracleEnv
    orc = OracleEnv(tree, Q.EnvConfig.hover_gym(battery=False))
    F, _ = orc.action_to_ctrl(np.zeros((1, 4), np.float32), np.array([8.4]))
    np.testing.assert_allclose(F, 6.5)
    A_inv = Q.EnvConfig().mixer()[1]
    np.testing.assert_allclose(A_inv[0], [0.25, -6.28156486, -6.28156486, 12.43781095], rtol=1e-8)


def test_momentum_conservation_first_order(tree):
    """No gravity, no fluid: energy and momenta drift only at the integrator's O(dt) and halve with dt."""
    rng = np.random.default_rng(0)
    drift = []
    for dt in (2e-5, 1e-5):
        P = TreePipeline(tree, fluid=False, gravity=False)
        P.dt = dt
        q = rng.normal(size=(3, 11)); q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
        v = rng.normal(size=(3, 10)) * 3; v[:, 6:] *= 50
        rng = np.random.default_rng(0)            # same state for both dt
        KE0, P0, L0 = P.momenta(q, v)
        for _ in range(int(round(4e-3 / dt))):
            q, v = P.step(q, v, np.zeros((3, 4)))
        KE1, P1, L1 = P.momenta(q, v)
        drift.append([np.abs((KE1 - KE0) / KE0).max(), np.abs(P1 - P0).max(), np.abs(L1 - L0).max()])
    d = np.array(drift)
    assert (d[0] < 1e-6).all()
    assert (d[1] < 0.75 * d[0] + 1e-12).all()      # first-order convergence


def test_mass_matrix_and_bias_independent_of_rotor_angle(tree):
    """Balanced axisymmetric rotors: M and the bias force do not depend on theta (fluid drag does)."""
    P = TreePipeline(tree, fluid=False)
    qpos, qvel = random_states(8, seed=2)
    q = qpos.astype(np.float64); v = qvel.astype(np.float64)
    _, M1, p1 = P.forward(q, v, np.zeros((8, 4)), return_parts=True)
    q2 = q.copy(); q2[:, 7:] += np.random.default_rng(1).uniform(-3, 3, (8, 4))
    _, M2, p2 = P.forward(q2, v, np.zeros((8, 4)), return_parts=True)
    # world-frame linear dofs: M depends on attitude only, not on theta
    np.testing.assert_allclose(M1, M2, atol=1e-18, rtol=1e-10)
    np.testing.assert_allclose(p1["bias"], p2["bias"], atol=1e-12, rtol=1e-9)
    Pf = TreePipeline(tree, fluid=True)
    f1, _ = Pf.forward(q, v, np.zeros((8, 4))); f2, _ = Pf.forward(q2, v, np.zeros((8, 4)))
    assert np.abs(f1 - f2).max() > 1e-9              # theta does enter through the prop drag


def test_c_restatement_equals_numpy(tree):
    qpos, qvel = random_states(256, seed=9)
    ctrl = np.random.default_rng(3).uniform(-2, 15, (256, 4))
    q1, v1 = TreePipeline(tree).step(qpos.astype(np.float64), qvel.astype(np.float64), ctrl)
    q2, v2 = cpu_ref.step_batch(tree, qpos, qvel, ctrl, threads=2)
    np.testing.assert_allclose(q2, q1, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(v2, v1, rtol=1e-11, atol=1e-12)


def test_zero_angular_velocity_quaternion_step(tree):
    P = TreePipeline(tree, fluid=False, gravity=False)
    q = np.zeros((1, 11)); q[0, 3:7] = [0.5, 0.5, 0.5, 0.5]
    q2, v2 = P.step(q, np.zeros((1, 10)), np.zeros((1, 4)))
    np.testing.assert_allclose(q2[0, 3:7], 0.5)
    assert np.isfinite(q2).all()
