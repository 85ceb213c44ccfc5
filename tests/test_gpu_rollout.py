"""GPU suite (-m gpu): fused policy rollout (qs_rollout_policy) and GAE (qs_gae) against the oracle."""
import numpy as np
import pytest

from oracle import ppo_ref
from oracle.envs import OracleEnv
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M

from .util import ATOL_OBS21, assert_close, planes_view

pytestmark = pytest.mark.gpu


def _tree():
    return M.load_mjcf(M.default_model_path())


def _random_policy(obs_dim, dist, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    n = ppo_ref.param_count(obs_dim, dist)
    parts = []
    Ao = 8 if dist == 1 else 4
    for out in (Ao, 1):
        for (i, o) in ((obs_dim, 128), (128, 128), (128, out)):
            lim = scale * (3.0 / i) ** 0.5
            parts += [rng.uniform(-lim, lim, i * o), rng.uniform(-0.05, 0.05, o)]
    if dist == 0:
        parts.append(rng.uniform(-1.5, -0.5, 4))
    parts += [rng.uniform(-0.2, 0.2, obs_dim), rng.uniform(0.5, 1.5, obs_dim)]
    p = np.concatenate(parts).astype(np.float32)
    assert p.size == n
    return p


@pytest.mark.parametrize("brax_form", [0, 1])
def test_gae_matches_oracle(brax_form):
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    T, B = 37, 1000
    rng = np.random.default_rng(brax_form)
    r = rng.normal(size=(T, B)).astype(np.float32); v = rng.normal(size=(T, B)).astype(np.float32)
    done = (rng.uniform(size=(T, B)) < 0.1).astype(np.float32)
    trunc = ((rng.uniform(size=(T, B)) < 0.05) * (1 if brax_form == 0 else done)).astype(np.float32)
    if brax_form == 0:
        trunc = trunc * (1 - done)
    lv = rng.normal(size=B).astype(np.float32)
    eng = Engine(Q.EnvConfig.north_star(), B, device=0)
    up = lambda a: torch.from_numpy(a).cuda()
    adv, ret = eng.gae(up(r), up(v), up(done), up(trunc), up(lv), 0.99, 0.95, brax_form=bool(brax_form))
    torch.cuda.synchronize()
    f = ppo_ref.gae_brax if brax_form else ppo_ref.gae_sb3
    a0, r0 = f(r.astype(np.float64), v.astype(np.float64), done.astype(np.float64), trunc.astype(np.float64), lv, 0.99, 0.95)
    assert_close(adv.cpu().numpy(), a0, rtol=1e-5, atol=1e-5, what="gae adv")
    assert_close(ret.cpu().numpy(), r0, rtol=1e-5, atol=1e-5, what="gae ret")


@pytest.mark.parametrize("B,T", [(100, 48), (32, 8), (4801, 4)])
def test_rollout_mjx_brax_teacher_forced(B, T):
    """mjx_brax obs IS the full state, so every transition in the trajectory can be re-derived by the
    oracle from the recorded obs_t and action_t: policy head/value/log-prob, next state, reward and
    flags are checked step by step at single-step tolerance; flags and counters bit-exact."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    cfg = Q.EnvConfig.mjx_brax(episode_length=20, auto_reset=Q.RESET_RESTORE_FIRST, seed=99, env_id_offset=5)
    eng = Engine(cfg, B, device=0)
    st = eng.new_state()
    first = torch.zeros(21, B, device="cuda")
    eng.reset(st, first_state=first)
    params = _random_policy(21, 1, seed=3, scale=0.7)
    buf = eng.rollout_policy(st, torch.from_numpy(params).cuda(), T=T, t0=7, dist=1, first_state=first)
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    st_np = st.cpu().numpy(); first_np = first.cpu().numpy()
    pp = ppo_ref.unpack(params, 21, 1)
    orc = OracleEnv(_tree(), cfg)
    ids = np.arange(B, dtype=np.uint32) + np.uint32(cfg.env_id_offset)
    ep_steps = np.zeros(B, np.int32); done_prev = np.zeros(B, np.float32)
    firsto = dict(qpos=first_np[:11].T.astype(np.float64), qvel=first_np[11:].T.astype(np.float64))
    n_done = 0
    for t in range(T):
        obs_t = b["obs"][t]
        head, value = ppo_ref.forward(pp, obs_t)
        eps = ppo_ref.policy_noise(cfg.seed, ids, 7 + t)
        raw, act, logp = ppo_ref.sample(pp, head, eps, 1)
        assert_close(b["act"][t], raw, rtol=2e-4, atol=2e-4, what=f"t={t} raw action")
        assert_close(b["value"][t], value, rtol=2e-4, atol=2e-4, what=f"t={t} value")
        assert_close(b["logp"][t], logp, rtol=1e-3, atol=2e-3, what=f"t={t} logp")
        # transition from the recorded state with the action the kernel actually applied
        s = OracleEnv.blank(B)
        s["qpos"] = obs_t[:, :11].astype(np.float64); s["qvel"] = obs_t[:, 11:].astype(np.float64)
        s["step_count"][:] = t; s["ep_steps"] = ep_steps.copy(); s["done_prev"] = done_prev.copy()
        act_gpu = np.tanh(b["act"][t].astype(np.float64)).astype(np.float32)
        o = orc.step(s, act_gpu, firsto)
        np.testing.assert_array_equal(b["done"][t], o["done"], err_msg=f"t={t} done")
        np.testing.assert_array_equal(b["trunc"][t], o["truncated"], err_msg=f"t={t} trunc")
        assert_close(b["reward"][t], o["reward"], rtol=2e-5, atol=2e-5, what=f"t={t} reward")
        nxt = b["obs"][t + 1] if t + 1 < T else b["last_obs"]
        assert_close(nxt, o["obs"], rtol=1e-5, atol=ATOL_OBS21, what=f"t={t} next obs", scale=obs_t)
        ep_steps, done_prev = s["ep_steps"], s["done_prev"]
        n_done += int(o["done"].sum())
    pv = planes_view(st_np)
    np.testing.assert_array_equal(pv["step_count"], T)
    np.testing.assert_array_equal(pv["ep_steps"], ep_steps)
    np.testing.assert_array_equal(pv["done_prev"], done_prev)
    _, v_last = ppo_ref.forward(pp, b["last_obs"])
    assert_close(b["last_value"], v_last, rtol=2e-4, atol=2e-4, what="last value")
    if T >= 20:
        assert n_done >= B          # every env hit the 20-step episode limit at least once


def test_rollout_hover_free_running_and_timeout_bootstrap():
    """north-star hover mode, SB3 Gaussian policy, ragged env count, short episodes so that truncation,
    termination, Philox auto-reset and the gamma*V(terminal_obs) bootstrap all occur."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    B, T, gamma = 203, 24, 0.97
    cfg = Q.EnvConfig.north_star(seed=11, env_id_offset=1000, max_episode_steps=6)
    eng = Engine(cfg, B, device=0)
    st = eng.new_state()
    eng.reset(st)
    torch.cuda.synchronize()
    s = OracleEnv.from_planes(st.cpu().numpy())
    params = _random_policy(12, 0, seed=5, scale=0.3)
    params[ppo_ref.param_count(12, 0) - 24 - 4: ppo_ref.param_count(12, 0) - 24] = -3.0      # small exploration noise
    buf = eng.rollout_policy(st, torch.from_numpy(params).cuda(), T=T, t0=0, dist=0, bootstrap_gamma=gamma)
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    pp = ppo_ref.unpack(params, 12, 0)
    orc = OracleEnv(_tree(), cfg)
    ids = np.arange(B, dtype=np.uint32) + np.uint32(cfg.env_id_offset)
    obs = orc.evaluate(s, None)["obs"]
    n_trunc = n_done = 0
    for t in range(T):
        tol = 2e-5 * (1.6 ** min(t, 5))      # fp32-vs-fp64 drift inside an episode (episodes last <= 6 steps)
        assert_close(b["obs"][t], obs, rtol=tol, atol=tol, what=f"t={t} obs")
        head, value = ppo_ref.forward(pp, obs)
        raw, act, logp = ppo_ref.sample(pp, head, ppo_ref.policy_noise(cfg.seed, ids, t), 0)
        assert_close(b["act"][t], raw, rtol=1e-3, atol=1e-3, what=f"t={t} action")
        assert_close(b["value"][t], value, rtol=1e-3, atol=1e-3, what=f"t={t} value")
        assert_close(b["logp"][t], logp, rtol=1e-4, atol=1e-4, what=f"t={t} logp")
        o = orc.step(s, act.astype(np.float32))
        np.testing.assert_array_equal(b["done"][t], o["done"], err_msg=f"t={t} done")
        np.testing.assert_array_equal(b["trunc"][t], o["truncated"], err_msg=f"t={t} trunc")
        boot = (o["truncated"] != 0) & (o["done"] == 0)
        rew = o["reward"].copy()
        if boot.any():
            _, vt = ppo_ref.forward(pp, o["terminal_obs"][boot])
            rew[boot] += gamma * vt
        assert_close(b["reward"][t], rew, rtol=1e-3, atol=1e-3, what=f"t={t} reward")
        n_trunc += int(boot.sum()); n_done += int(o["done"].sum())
        obs = o["obs"]
    assert n_trunc > 0          # (terminations are exercised by the teacher-forced and single-step cases)
    pv = planes_view(st.cpu().numpy())
    np.testing.assert_array_equal(pv["episode"], s["episode"])
    np.testing.assert_array_equal(pv["step_count"], s["step_count"])
    assert_close(pv["qpos"], s["qpos"], rtol=1e-3, atol=1e-3, what="final qpos")


def test_rollout_random_matches_stepwise():
    """The state-resident dyn-only kernel (qs_rollout_random) must equal T calls of qs_step fed with the
    same Philox actions: identical source, so bit-exact state."""
    import torch
    from oracle import philox
    from uav_reinforcement_learning_control_b200.engine import Engine
    B, T = 777, 20
    cfg = Q.EnvConfig.north_star(seed=3, env_id_offset=10)
    eng = Engine(cfg, B, device=0)
    st_a = eng.new_state(); eng.reset(st_a)
    st_b = st_a.clone()
    stats = torch.zeros(4, B, device="cuda")
    eng.rollout_random(st_a, T, t0=100, stats=stats)
    ids = np.arange(B, dtype=np.uint32) + np.uint32(10)
    fin = np.zeros(B); rsum = np.zeros(B)
    for t in range(T):
        raw = philox.draw_blocks(cfg.seed, ids, np.uint32(100 + t), 1, philox.STREAM_ACTION)
        act = np.stack([philox.uniform(raw[:, i], -1.0, 1.0) for i in range(4)], axis=1)
        trunc = torch.zeros(B, device="cuda")
        obs, rew, done = eng.step(st_b, torch.from_numpy(act).cuda(), truncated=trunc)
        fin += np.maximum(done.cpu().numpy(), trunc.cpu().numpy()); rsum += rew.cpu().numpy()
    torch.cuda.synchronize()
    a, b = st_a.cpu().numpy(), st_b.cpu().numpy()
    np.testing.assert_array_equal(a[:27].view(np.uint32), b[:27].view(np.uint32))
    np.testing.assert_array_equal(stats[1].cpu().numpy(), fin)
    assert_close(stats[0].cpu().numpy(), rsum, rtol=1e-5, atol=1e-5, what="reward sums")


@pytest.mark.parametrize("dist,B,T", [(0, 128, 6), (0, 300, 5), (1, 1000, 4)])
def test_rollout_tensor_core_path(dist, B, T):
    """tcgen05/TMEM policy forward: given the observations the kernel recorded, its actions / values / log-probs
    must match the oracle's bf16-operand, fp32-accumulate forward pass; bookkeeping must stay consistent."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    cfg = Q.EnvConfig.north_star(seed=21, env_id_offset=64, max_episode_steps=4)
    eng = Engine(cfg, B, device=0)
    st = eng.new_state(); eng.reset(st)
    params = _random_policy(12, dist, seed=17, scale=0.6)
    buf = eng.rollout_policy(st, torch.from_numpy(params).cuda(), T=T, t0=3, dist=dist, tensor_cores=True, bootstrap_gamma=0.9)
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    pp = ppo_ref.unpack(params, 12, dist)
    ids = np.arange(B, dtype=np.uint32) + np.uint32(64)
    for t in range(T):
        head, value = ppo_ref.forward(pp, b["obs"][t], bf16=True)
        raw, act, logp = ppo_ref.sample(pp, head, ppo_ref.policy_noise(cfg.seed, ids, 3 + t), dist)
        assert_close(b["value"][t], value, rtol=2e-3, atol=2e-3, what=f"t={t} value (tcgen05)")
        assert_close(b["act"][t], raw, rtol=2e-3, atol=2e-3, what=f"t={t} action (tcgen05)")
        assert_close(b["logp"][t], logp, rtol=5e-3, atol=5e-3, what=f"t={t} logp (tcgen05)")
        # and it is close to the exact fp32 network too (bf16 operand rounding only)
        head32, value32 = ppo_ref.forward(pp, b["obs"][t])
        assert np.abs(value - value32).max() < 0.1
    assert np.isfinite(b["obs"]).all() and np.isfinite(b["reward"]).all()
    _, v_last = ppo_ref.forward(pp, b["last_obs"], bf16=True)
    assert_close(b["last_value"], v_last, rtol=2e-3, atol=2e-3, what="last value (tcgen05)")
    pv = planes_view(st.cpu().numpy())
    fin = np.maximum(b["done"], b["trunc"]).sum(axis=0)
    np.testing.assert_array_equal(pv["episode"], fin.astype(np.uint32))       # one reset per finished episode
    assert (b["trunc"].sum() > 0)                                             # 4-step episodes did truncate
    # the same rollout on the fp32 FMA path sees the same first observation and similar values
    st2 = eng.new_state(); eng.reset(st2)
    buf2 = eng.rollout_policy(st2, torch.from_numpy(params).cuda(), T=1, t0=3, dist=dist)
    np.testing.assert_array_equal(buf2["obs"][0].cpu().numpy(), b["obs"][0])
    assert np.abs(buf2["value"][0].cpu().numpy() - b["value"][0]).max() < 0.1


@pytest.mark.parametrize("mode,B,T", [("mjx_brax", 300, 6), ("hover_brax", 130, 5), ("mjx_brax", 1100, 5)])
def test_rollout_tensor_core_path_21d(mode, B, T):
    """tcgen05/TMEM policy forward on the 21-D raw observation of the Brax envs (layer-1 K = 32, tanh-normal head):
    same checks as the 12-D case -- the recorded observations through the oracle's bf16-operand forward must give the
    kernel's actions / values / log-probs -- plus agreement of every recorded stream with the fp32 FMA kernel's
    bookkeeping (done / trunc flags are functions of the state, which the two paths evolve with the same actions only
    approximately, so they are compared on the first step, where both start from identical states)."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    mk = Q.EnvConfig.mjx_brax if mode == "mjx_brax" else Q.EnvConfig.hover_brax
    cfg = mk(episode_length=4, auto_reset=Q.RESET_RESTORE_FIRST, seed=31, env_id_offset=9)
    eng = Engine(cfg, B, device=0)
    st = eng.new_state(); first = torch.zeros(21, B, device="cuda")
    eng.reset(st, first_state=first)
    st0 = st.clone()
    params = _random_policy(21, 1, seed=5, scale=0.6)
    d_params = torch.from_numpy(params).cuda()
    buf = eng.rollout_policy(st, d_params, T=T, t0=11, dist=1, first_state=first, tensor_cores=True)
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    pp = ppo_ref.unpack(params, 21, 1)
    ids = np.arange(B, dtype=np.uint32) + np.uint32(9)
    for t in range(T):
        head, value = ppo_ref.forward(pp, b["obs"][t], bf16=True)
        raw, act, logp = ppo_ref.sample(pp, head, ppo_ref.policy_noise(cfg.seed, ids, 11 + t), 1)
        assert_close(b["value"][t], value, rtol=2e-3, atol=2e-3, what=f"t={t} value (tcgen05, 21-D)")
        assert_close(b["act"][t], raw, rtol=3e-3, atol=3e-3, what=f"t={t} action (tcgen05, 21-D)")
        assert_close(b["logp"][t], logp, rtol=1e-2, atol=1e-2, what=f"t={t} logp (tcgen05, 21-D)")
    _, v_last = ppo_ref.forward(pp, b["last_obs"], bf16=True)
    assert_close(b["last_value"], v_last, rtol=2e-3, atol=2e-3, what="last value (tcgen05, 21-D)")
    assert np.isfinite(b["obs"]).all() and np.isfinite(b["reward"]).all()
    # EpisodeWrapper: 4-step episodes -> an env that did not terminate earlier is done (truncated) at t = 3 and
    # restored to its first observation at t = 4 (AutoResetWrapper)
    full = b["done"][:3].sum(axis=0) == 0
    assert (b["done"][3][full] == 1).all() and (b["trunc"][3][full] == 1).all()
    if T > 4:
        was_done = b["done"][3] == 1
        np.testing.assert_array_equal(b["obs"][4][was_done], b["obs"][0][was_done])
    # against the fp32 FMA kernel from the same start: identical first observation, same flags on step 0
    buf2 = eng.rollout_policy(st0, d_params, T=1, t0=11, dist=1, first_state=first)
    np.testing.assert_array_equal(buf2["obs"][0].cpu().numpy(), b["obs"][0])
    np.testing.assert_array_equal(buf2["done"][0].cpu().numpy(), b["done"][0])
    assert np.abs(buf2["value"][0].cpu().numpy() - b["value"][0]).max() < 0.1
    # and the FMA path's own numbers against the exact fp32 oracle forward
    head32, value32 = ppo_ref.forward(pp, b["obs"][0])
    assert_close(buf2["value"][0].cpu().numpy(), value32, rtol=2e-4, atol=2e-4, what="value (fp32 FMA, 21-D)")
