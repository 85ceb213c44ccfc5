"""Backend-agnostic parity cases: the same checks run against
  * the g++ host build of the kernels' per-env source (CPU suite, `-m "not gpu"`), and
  * the real sm_100a library through the C ABI (GPU suite, `-m gpu`),
each compared with oracle/ on identical seeded inputs.

Bars (BASELINE.json north_star): done / truncated flags, step counters, episode (reset)
indices and waypoint indices bit-exact; float32 state within 1e-5 relative / 1e-6 absolute
of the float64 oracle after one step.
"""
from __future__ import annotations

import numpy as np

from oracle.envs import OracleEnv
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M

from .util import ATOL_OBS21, ATOL_QVEL, assert_close, make_planes, planes_view, random_states

_TREE = None


def tree():
    global _TREE
    if _TREE is None:
        _TREE = M.load_mjcf(M.default_model_path())
    return _TREE


CONFIGS = {
    "hover_gym": lambda: Q.EnvConfig.hover_gym(),
    "north_star": lambda: Q.EnvConfig.north_star(),
    "hover_gym_autoreset": lambda: Q.EnvConfig.hover_gym(auto_reset=Q.RESET_RESAMPLE, seed=1234, env_id_offset=77),
    "traj_gym": lambda: Q.EnvConfig.traj_gym(),
    "mjx_brax": lambda: Q.EnvConfig.mjx_brax(),
    "mjx_brax_wrapped": lambda: Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST),
    "hover_brax": lambda: Q.EnvConfig.hover_brax(),
    "hover_brax_wrapped": lambda: Q.EnvConfig.hover_brax(episode_length=100, auto_reset=Q.RESET_RESTORE_FIRST),
    "mjx_playground": lambda: Q.EnvConfig.mjx_playground(),
    # SB3 default training wrapper (train.py:31): CTBR action -> torque PI loop fused as an action pre-stage
    "hover_gym_rate": lambda: Q.EnvConfig.hover_gym(rate_wrapper=True, auto_reset=Q.RESET_RESAMPLE, seed=77),
}


def synth_inputs(cfg, n, seed):
    """Synthetic uniformly randomised states that straddle every termination bound."""
    gym = cfg.mode in (Q.MODE_HOVER_GYM, Q.MODE_TRAJ_GYM)
    pos = 2.3 if cfg.mode == Q.MODE_HOVER_GYM else 3.3
    qpos, qvel = random_states(n, seed=seed, pos=pos, vel=11.0 if gym else 21.0, omega=20.0)
    rng = np.random.default_rng(seed + 1)
    qpos[:, 2] = rng.uniform(-0.1, 2.2 if cfg.mode == Q.MODE_HOVER_GYM else 4.2, n).astype(np.float32)
    act = rng.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
    tgt = rng.uniform(-1.5, 1.8, (n, 3)).astype(np.float32)
    sc = rng.integers(0, cfg.max_episode_steps + 8, n).astype(np.int32)
    # a few exactly at the truncation edge
    sc[:8] = cfg.max_episode_steps - 1
    volt = rng.uniform(cfg.v_min, cfg.v_nominal, n).astype(np.float32)
    ep_steps = rng.integers(0, max(cfg.episode_length, 1), n).astype(np.int32)
    if cfg.episode_length > 0:
        ep_steps[:16] = cfg.episode_length - 1
    done_prev = (rng.uniform(size=n) < 0.2).astype(np.float32)
    st = make_planes(n, qpos, qvel, target=tgt, step_count=sc, voltage=volt, episode=rng.integers(0, 5, n),
                     ep_steps=ep_steps, done_prev=done_prev, rate_int=rng.uniform(-0.012, 0.012, (n, 3)))
    # non-finite injections: termination / NaN->0 paths
    st[0, 20] = np.nan; st[12, 21] = np.inf; st[4, 22] = np.nan; st[17, 23] = -np.inf
    first = None
    if cfg.auto_reset == Q.RESET_RESTORE_FIRST:
        fq, fv = random_states(n, seed=seed + 7, pos=0.01, vel=0.01, omega=0.01, spin=0.01, theta=0.01)
        first = np.ascontiguousarray(np.concatenate([fq.T, fv.T]).astype(np.float32))
    return st, act, first


def check_single_step(backend_factory, name, n=4096, seed=0):
    cfg = CONFIGS[name]()
    backend = backend_factory(cfg)
    orc = OracleEnv(tree(), cfg)
    st, act, first = synth_inputs(cfg, n, seed)
    prev = planes_view(st.copy())
    s = OracleEnv.from_planes(st)
    firsto = None if first is None else dict(qpos=first[:11].T.astype(np.float64), qvel=first[11:].T.astype(np.float64))
    o = orc.step(s, act, firsto)
    h = backend.step(st, act, first=first, want_term=True)
    pv = planes_view(st)

    # --- bit-exact items ---------------------------------------------------------------
    np.testing.assert_array_equal(h["done"], o["done"], err_msg=f"{name}: done flags")
    np.testing.assert_array_equal(h["truncated"], o["truncated"], err_msg=f"{name}: truncated flags")
    if cfg.mode != Q.MODE_HOVER_BRAX:
        np.testing.assert_array_equal(pv["step_count"], s["step_count"], err_msg=f"{name}: step_count")
    gym = cfg.mode in (Q.MODE_HOVER_GYM, Q.MODE_TRAJ_GYM)
    if gym:
        np.testing.assert_array_equal(pv["episode"], s["episode"], err_msg=f"{name}: episode (reset) index")
    if cfg.mode in (Q.MODE_MJX_BRAX, Q.MODE_HOVER_BRAX) and (cfg.episode_length > 0 or cfg.auto_reset):
        np.testing.assert_array_equal(pv["ep_steps"], s["ep_steps"], err_msg=f"{name}: ep_steps")
        if cfg.auto_reset == Q.RESET_RESTORE_FIRST:
            np.testing.assert_array_equal(pv["done_prev"], s["done_prev"], err_msg=f"{name}: done_prev")
    if h["terminal_obs"] is not None:               # (a lean backend requests neither terminal_obs nor metrics)
        term_written = ~np.isnan(h["terminal_obs"]).all(axis=1)
        finite_in = np.isfinite(o["terminal_obs"]).all(axis=1) | ~o["finished"]
        np.testing.assert_array_equal(term_written[finite_in], o["finished"][finite_in],
                                      err_msg=f"{name}: which envs wrote terminal_obs")

    # --- float items ---------------------------------------------------------------------
    ok = np.isfinite(s["qpos"]).all(axis=1) & np.isfinite(s["qvel"]).all(axis=1)
    reset = o["finished"] & (cfg.auto_reset != Q.RESET_NONE)
    keep = ok & ~reset
    assert keep.sum() > n // 10
    assert_close(pv["qpos"][keep], s["qpos"][keep], what=f"{name}: qpos", scale=prev["qpos"][keep])
    assert_close(pv["qvel"][keep], s["qvel"][keep], what=f"{name}: qvel", scale=prev["qvel"][keep], atol=ATOL_QVEL)
    if reset.any():
        # freshly reset / restored states: float32 op-for-op reproducible up to sincos rounding
        assert_close(pv["qpos"][reset], s["qpos"][reset], what=f"{name}: reset qpos", rtol=1e-6, atol=1e-7)
        assert_close(pv["qvel"][reset], s["qvel"][reset], what=f"{name}: reset qvel", rtol=1e-6, atol=1e-7)
        if gym:
            np.testing.assert_array_equal(pv["target"][reset], s["target"][reset], err_msg=f"{name}: reset target")
    fin_r = np.isfinite(o["reward"])
    assert_close(h["reward"][fin_r], o["reward"][fin_r], what=f"{name}: reward")
    # observations: angles near +-pi wrap, compare on the circle for the gym attitude entries
    oo, ho = o["obs"].astype(np.float64), h["obs"].astype(np.float64)
    # rows whose INPUT state was non-finite are excluded from float comparisons: the generic pipeline
    # (like MuJoCo's dense solve) spreads one NaN to every coordinate, the closed form keeps it
    # contained; both flag the env done (checked bit-exactly above), which is what the reference tests
    okobs = np.isfinite(oo).all(axis=1) & ok
    if gym:
        d = ho[:, 3:6] - oo[:, 3:6]
        d = (d + 1.0) % 2.0 - 1.0                # normalised angle lives on [-1, 1)
        ho = ho.copy(); ho[:, 3:6] = oo[:, 3:6] + d
    if gym:
        # normalised obs: the velocity entries are state / bound, so scale the pre-step state alike
        sc_obs = np.zeros_like(oo)
        sc_obs[:, 6:9] = prev["qvel"][:, 0:3] / 10.0; sc_obs[:, 9:12] = prev["qvel"][:, 3:6] / (6 * np.pi)
    else:
        sc_obs = np.concatenate([prev["qpos"], prev["qvel"]], axis=1)
    sc_obs = np.where(reset[:, None], 0.0, sc_obs)
    assert_close(ho[okobs], oo[okobs], what=f"{name}: obs", rtol=2e-5, atol=2e-6 if gym else 2 * ATOL_OBS21,
                 scale=sc_obs[okobs])
    if cfg.battery:
        assert_close(pv["voltage"], s["voltage"], what=f"{name}: voltage")
    if cfg.rate_wrapper:
        assert_close(pv["rate_int"], s["rate_int"], what=f"{name}: rate integral", rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(pv["prev_action"], s["prev_action"], err_msg=f"{name}: prev_action")
    for i, k in enumerate(["pos_error", "reward_hover", "reward_action", "reward"] if h["metrics"] is not None else []):
        f = np.isfinite(o[k])
        assert_close(h["metrics"][i][f], np.asarray(o[k])[f], what=f"{name}: metric {k}", rtol=2e-5, atol=2e-6)
    return dict(done=int(o["done"].sum()), truncated=int(o["truncated"].sum()), finished=int(o["finished"].sum()))


def check_reset(backend_factory, name, n=2048):
    cfg = CONFIGS[name]()
    backend = backend_factory(cfg)
    orc = OracleEnv(tree(), cfg)
    rng = np.random.default_rng(5)
    st = make_planes(n, episode=rng.integers(0, 1000, n))
    s = OracleEnv.from_planes(st)
    mask = (rng.uniform(size=n) < 0.7)
    orc.reset(s, mask)
    obs, first = backend.reset(st, mask.astype(np.uint8), want_first=True)
    pv = planes_view(st)
    m = mask
    gym = cfg.mode in (Q.MODE_HOVER_GYM, Q.MODE_TRAJ_GYM)
    # positions / velocities / targets are straight float32 Philox draws: bit-exact
    np.testing.assert_array_equal(pv["qpos"][m][:, 0:3], s["qpos"][m][:, 0:3].astype(np.float32))
    np.testing.assert_array_equal(pv["qvel"][m], s["qvel"][m].astype(np.float32))
    assert_close(pv["qpos"][m][:, 3:7], s["qpos"][m][:, 3:7], what=f"{name}: reset quat", rtol=1e-6, atol=2e-7)
    np.testing.assert_array_equal(pv["qpos"][m][:, 7:11], s["qpos"][m][:, 7:11].astype(np.float32))
    if gym:
        np.testing.assert_array_equal(pv["target"][m], s["target"][m])
    np.testing.assert_array_equal(pv["step_count"][m], 0)
    # untouched envs keep their (blank) state
    np.testing.assert_array_equal(pv["qpos"][~m][:, 3], 1.0)
    np.testing.assert_array_equal(pv["qpos"][~m][:, 0:3], 0.0)
    ev = orc.evaluate(s, None)
    assert_close(obs[m], ev["obs"][m], what=f"{name}: reset obs", rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(first[:11].T[m], pv["qpos"][m])
    np.testing.assert_array_equal(first[11:].T[m], pv["qvel"][m])


def check_observe_bit_exact_masks(backend_factory, name, n=8192, seed=3):
    """R5: masks decided on the SAME float32 state must be bit-exact, including at the bound edges."""
    cfg = CONFIGS[name]()
    backend = backend_factory(cfg)
    orc = OracleEnv(tree(), cfg)
    st, act, _ = synth_inputs(cfg, n, seed)
    pv = planes_view(st)
    # plant states exactly ON the bounds and one ulp either side
    if cfg.mode in (Q.MODE_HOVER_GYM, Q.MODE_TRAJ_GYM):
        hi = np.float32(cfg.term_hi[0]); lo_z = np.float32(cfg.term_lo[2])
        vals = [hi, np.nextafter(hi, np.float32(np.inf)), np.nextafter(hi, np.float32(-np.inf))]
        for k, v in enumerate(vals):
            st[0, 100 + k] = v; st[1, 110 + k] = -v
            st[11, 120 + k] = np.float32(10.0) + np.float32(k - 1) * np.float32(1e-6)
        st[2, 130] = lo_z; st[2, 131] = np.nextafter(lo_z, np.float32(-1))
    else:
        lim = np.float32(cfg.pos_limit_xy)
        for k, v in enumerate([lim, np.nextafter(lim, np.float32(np.inf)), np.nextafter(lim, np.float32(-np.inf))]):
            st[0, 100 + k] = v; st[1, 110 + k] = -v
        st[2, 130] = np.float32(cfg.z_low); st[2, 131] = np.nextafter(np.float32(cfg.z_low), np.float32(-1))
        st[11, 140] = np.float32(cfg.vel_limit); st[11, 141] = np.nextafter(np.float32(cfg.vel_limit), np.float32(np.inf))
    s = OracleEnv.from_planes(st)
    ev = orc.evaluate(s, act)
    obs, rew, done = backend.observe(st, act)
    np.testing.assert_array_equal(done, ev["done"], err_msg=f"{name}: done on injected state")
    f = np.isfinite(ev["reward"])
    assert_close(rew[f], ev["reward"][f], what=f"{name}: reward on injected state", rtol=2e-5, atol=2e-6)
    return int(ev["done"].sum())


def check_waypoints(backend_factory, n=600, steps=6, auto_reset=True):
    """evaluate.py:535-557 waypoint advance: indices, reach counters, laps and targets bit-exact vs the oracle,
    over several steps with states planted inside / outside / exactly on the reach radius."""
    from uav_reinforcement_learning_control_b200 import trajectories as TJ
    tables = TJ.default_tables(0.5)
    cfg = Q.EnvConfig.waypoint_eval(tables, auto_reset=Q.RESET_RESAMPLE if auto_reset else Q.RESET_NONE,
                                    env_id_offset=4, max_episode_steps=50)
    backend = backend_factory(cfg)
    orc = OracleEnv(tree(), cfg)
    st = make_planes(n)
    s = OracleEnv.from_planes(st)
    orc.reset(s)
    backend.reset(st)
    pv = planes_view(st)
    np.testing.assert_array_equal(pv["wp_idx"], s["wp_idx"]); np.testing.assert_array_equal(pv["target"], s["target"])
    np.testing.assert_array_equal(pv["qpos"][:, 0:3], s["qpos"][:, 0:3].astype(np.float32))
    rng = np.random.default_rng(8)
    total_reached = 0
    for t in range(steps):
        # teleport every env next to its current target (some inside the 0.25 m radius, some outside), at rest
        ids = np.arange(n) + cfg.env_id_offset
        off = rng.normal(size=(n, 3)); off /= np.linalg.norm(off, axis=1, keepdims=True)
        off *= rng.choice([0.05, 0.2, 0.2495, 0.2505, 0.4], size=(n, 1))
        wp = np.stack([tables[i % 3][pv["wp_idx"][k]] for k, i in enumerate(ids)])
        # last waypoint before the lap closes for a few envs, so laps complete
        pos = (wp + off).astype(np.float32)
        st[0:3] = pos.T; st[3] = 1; st[4:7] = 0; st[11:21] = 0
        s["qpos"][:, 0:3] = pos; s["qpos"][:, 3:7] = [1, 0, 0, 0]; s["qpos"][:, 7:] = 0; s["qvel"][:] = 0
        act = np.tile(np.array([[-0.958, 0, 0, 0]], np.float32), (n, 1))      # ~hover thrust: the env barely moves
        o = orc.step(s, act)
        h = backend.step(st, act, want_term=True)
        pv = planes_view(st)
        np.testing.assert_array_equal(h["done"], o["done"], err_msg=f"t={t} done")
        np.testing.assert_array_equal(h["truncated"], o["truncated"], err_msg=f"t={t} trunc")
        np.testing.assert_array_equal(pv["wp_idx"], s["wp_idx"], err_msg=f"t={t} wp_idx")
        np.testing.assert_array_equal(pv["wp_reached"], s["wp_reached"], err_msg=f"t={t} wp_reached")
        np.testing.assert_array_equal(pv["laps"], s["laps"], err_msg=f"t={t} laps")
        np.testing.assert_array_equal(pv["target"], s["target"], err_msg=f"t={t} target")
        np.testing.assert_array_equal(pv["episode"], s["episode"], err_msg=f"t={t} episode")
        np.testing.assert_array_equal(pv["step_count"], s["step_count"], err_msg=f"t={t} step_count")
        assert_close(h["obs"], o["obs"], rtol=2e-5, atol=2e-6, what=f"t={t} obs")
        total_reached = int(s["wp_reached"].sum())
    assert total_reached > n           # most envs advanced more than once
    return total_reached


def check_lap_completion(backend_factory):
    """An env walked around the whole table closes the lap: laps += 1, episode ends (evaluate.py:552-555)."""
    from uav_reinforcement_learning_control_b200 import trajectories as TJ
    tables = [TJ.square(0.5)]
    cfg = Q.EnvConfig.waypoint_eval(tables, auto_reset=Q.RESET_RESAMPLE)
    backend = backend_factory(cfg)
    orc = OracleEnv(tree(), cfg)
    n = 64
    st = make_planes(n); s = OracleEnv.from_planes(st)
    orc.reset(s); backend.reset(st)
    act = np.tile(np.array([[-0.958, 0, 0, 0]], np.float32), (n, 1))
    laps_seen = 0
    for t in range(len(tables[0]) + 2):
        pv = planes_view(st)
        pos = np.stack([tables[0][k] for k in pv["wp_idx"]]).astype(np.float32)
        st[0:3] = pos.T; st[3] = 1; st[4:7] = 0; st[11:21] = 0
        s["qpos"][:, 0:3] = pos; s["qpos"][:, 3:7] = [1, 0, 0, 0]; s["qpos"][:, 7:] = 0; s["qvel"][:] = 0
        o = orc.step(s, act); h = backend.step(st, act, want_term=True)
        pv = planes_view(st)
        np.testing.assert_array_equal(pv["laps"], s["laps"]); np.testing.assert_array_equal(pv["wp_idx"], s["wp_idx"])
        np.testing.assert_array_equal(pv["episode"], s["episode"])
        laps_seen = int(s["laps"].sum())
    assert laps_seen == n              # every env completed exactly one lap and was reset to waypoint 0 / target 1
    np.testing.assert_array_equal(planes_view(st)["episode"], 1)


def check_traj_info(backend_factory, n=256):
    """TrajectoryFollowEnv info target / target_vel / target_acc: engine vs scipy (oracle/traj_spline.py)."""
    from oracle import traj_spline
    for kw in (dict(), dict(seed=99, env_id_offset=1000, spline_duration=None, max_episode_steps=300)):
        cfg = Q.EnvConfig.traj_gym(**kw)
        backend = backend_factory(cfg)
        rng = np.random.default_rng(3)
        episode = rng.integers(0, 50, n).astype(np.uint32)
        N = cfg.max_episode_steps
        idx = rng.integers(0, N, n).astype(np.int32)
        idx[:8] = [0, 0, N - 1, N - 1, 1, N - 2, N // 2, N // 3]
        got = backend.traj_info(episode, idx)
        ids = np.arange(n, dtype=np.uint32) + np.uint32(cfg.env_id_offset)
        want = traj_spline.info(cfg, ids, episode, idx)
        # float64 spline on both sides, cast to float32: agreement to float32 rounding
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
        # the trajectory starts at the drone's start position (trajectory_follow_env.py:208-209, 241-243) ...
        start, _, n_wp, _ = traj_spline.draws(cfg, ids, episode)
        at0 = backend.traj_info(episode, np.zeros(n, np.int32))
        np.testing.assert_allclose(at0[:, 0:3], start, rtol=0, atol=1e-6)
        # ... with zero curvature at both ends (natural boundary condition)
        atN = backend.traj_info(episode, np.full(n, N - 1, np.int32))
        assert np.abs(at0[:, 6:9]).max() < 1e-5 and np.abs(atN[:, 6:9]).max() < 1e-4
        assert set(np.unique(n_wp)) <= {3, 4, 5} and len(np.unique(n_wp)) == 3
        # out-of-range indices clamp like min(step_count - 1, N - 1) / the reset info at index 0
        np.testing.assert_array_equal(backend.traj_info(episode, np.full(n, -1, np.int32)), at0)
        np.testing.assert_array_equal(backend.traj_info(episode, np.full(n, N + 7, np.int32)), atN)
