"""GPU suite (-m gpu): BASELINE.json full sizes (2^20 envs; 8192 x 1024 rollout), checked through size-independent
properties because the float64 oracle cannot run a million envs in seconds:
determinism, shard invariance (Philox keyed by global env id), counter identities, unit quaternions,
resident-kernel == repeated step kernel, and oracle spot checks on a random subset."""
import numpy as np
import pytest

from oracle.envs import OracleEnv
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M

from .util import ATOL_QVEL, assert_close

pytestmark = pytest.mark.gpu
N = 1 << 20


def _run(cfg, n, steps, seed_actions=0):
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    eng = Engine(cfg, n, device=0)
    st = eng.new_state(); eng.reset(st)
    g = torch.Generator(device="cuda"); g.manual_seed(seed_actions)
    fin_total = torch.zeros(n, device="cuda")
    trunc = torch.zeros(n, device="cuda")
    acts = []
    for t in range(steps):
        a = torch.rand(n + 8, 4, device="cuda", generator=g)[:n] * 2 - 1
        acts.append(a)
        obs, rew, done = eng.step(st, a, truncated=trunc)
        fin_total += torch.maximum(done, trunc)
    torch.cuda.synchronize()
    return eng, st, fin_total, obs, acts


def test_full_size_step_properties():
    import torch
    cfg = Q.EnvConfig.north_star(seed=3)
    eng, st, fin, obs, acts = _run(cfg, N, 12)
    # determinism: the same launch sequence reproduces the state bit for bit
    _, st2, fin2, _, _ = _run(cfg, N, 12)
    assert torch.equal(st.view(torch.int32), st2.view(torch.int32)) and torch.equal(fin, fin2)
    # counter identity: episode index == number of finished episodes; step_count <= episode limit
    epi = st[26].view(torch.int32)
    assert torch.equal(epi.to(torch.float32), fin)
    sc = st[24].view(torch.int32)
    assert int(sc.min()) >= 0 and int(sc.max()) <= cfg.max_episode_steps
    # unit quaternions, finite everything, normalised observations in a sane range
    qn = (st[3:7] ** 2).sum(0).sqrt()
    assert float((qn - 1).abs().max()) < 2e-6
    assert bool(torch.isfinite(st[:27]).all()) and bool(torch.isfinite(obs).all())
    assert float(obs[:, 0:3].abs().max()) <= 1.0 + 1e-5
    # a healthy fraction of the million envs went through the Philox auto-reset path
    assert 0.2 * N < float((fin > 0).sum()) <= N


def test_shard_invariance_at_full_size():
    """Two shards of 2^19 envs (env_id_offset) reproduce one shard of 2^20 bit for bit, incl. auto-resets."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    steps = 6
    cfg = Q.EnvConfig.north_star(seed=9)
    eng, st, _, _, acts = _run(cfg, N, steps, seed_actions=5)
    half = N // 2
    parts = []
    for r in range(2):
        e2 = Engine(Q.EnvConfig.north_star(seed=9, env_id_offset=r * half), half, device=0)
        s2 = e2.new_state(); e2.reset(s2)
        for t in range(steps):
            e2.step(s2, acts[t][r * half:(r + 1) * half].contiguous())
        parts.append(s2)
    torch.cuda.synchronize()
    both = torch.cat(parts, dim=1)
    assert torch.equal(both[:27].view(torch.int32), st[:27].view(torch.int32))


def test_full_size_oracle_spot_check():
    """One step of 2^20 envs; 2048 randomly chosen envs are re-derived by the float64 oracle."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    cfg = Q.EnvConfig.north_star(seed=21)
    eng = Engine(cfg, N, device=0)
    st = eng.new_state(); eng.reset(st)
    a = torch.rand(N, 4, device="cuda") * 2 - 1
    for _ in range(3):
        eng.step(st, a)
    before = st.clone()
    obs, rew, done = eng.step(st, a)
    torch.cuda.synchronize()
    idx = torch.randperm(N, device="cuda")[:2048]
    sub = before[:, idx].cpu().numpy(); after = st[:, idx].cpu().numpy()
    cfg_sub = Q.EnvConfig.north_star(seed=21)             # resets of the subset are excluded below (ids differ)
    orc = OracleEnv(M.load_mjcf(M.default_model_path()), cfg_sub)
    s = OracleEnv.from_planes(sub)
    o = orc.step(s, a[idx].cpu().numpy())
    np.testing.assert_array_equal(done[idx].cpu().numpy(), o["done"])
    keep = ~o["finished"]
    assert keep.sum() > 1000
    assert_close(after[0:11].T[keep], s["qpos"][keep], what="qpos", scale=sub[0:11].T[keep])
    assert_close(after[11:21].T[keep], s["qvel"][keep], what="qvel", scale=sub[11:21].T[keep], atol=ATOL_QVEL)
    assert_close(rew[idx].cpu().numpy(), o["reward"], rtol=2e-5, atol=2e-6, what="reward")


def test_rollout_config_8192x1024_identities():
    """BASELINE.json configs[2] at full size: trajectory buffers are finite, flags are 0/1, the number of episode
    ends in the trajectory equals the episode counters, GAE returns satisfy ret = adv + value."""
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine
    from bench import make_policy_params
    B, T = 8192, 1024
    eng = Engine(Q.EnvConfig.north_star(seed=2), B, device=0)
    st = eng.new_state(); eng.reset(st)
    p = make_policy_params(eng, torch, torch.device("cuda"), seed=0)
    for tc in (False, True):
        epi0 = st[26].view(torch.int32).clone()
        b = eng.rollout_policy(st, p, T=T, t0=0, dist=0, tensor_cores=tc)
        adv, ret = eng.gae(b["reward"], b["value"], b["done"], b["trunc"], b["last_value"], 0.99, 0.95)
        torch.cuda.synchronize()
        for k in ("obs", "act", "logp", "value", "reward", "last_value"):
            assert bool(torch.isfinite(b[k]).all()), (k, tc)
        for k in ("done", "trunc"):
            assert bool(((b[k] == 0) | (b[k] == 1)).all())
        ends = torch.maximum(b["done"], b["trunc"]).sum(0)
        assert torch.equal((st[26].view(torch.int32) - epi0).to(torch.float32), ends)
        assert float((ret - (adv + b["value"])).abs().max()) < 1e-5
        assert float(b["obs"][:, :, 0:3].abs().max()) <= 1.0 + 1e-5 and float(b["reward"].min()) >= 0.0


@pytest.mark.parametrize("waypoints", [False, True], ids=["hover", "waypoint"])
def test_rollout_large_batch_forms_bitwise_equal_at_full_size(waypoints, monkeypatch):
    """The 2^18-env rollout legs (BASELINE.json configs[3] and the large-batch throughput leg) at full size: the compact
    two-tile CTAs the launcher picks there, the plain two-tile CTAs and the launcher's default record bitwise the same
    trajectories over 48 steps (two chained launches), with the same episode-end identities as the 8192-env test.  Together
    with the small-batch test in test_gpu_rollout.py (all forms == the one-tile form == the oracle lock-step) this is the
    parity statement for the batch sizes the oracle cannot step."""
    import torch
    from uav_reinforcement_learning_control_b200 import policies, trajectories as TJ
    from uav_reinforcement_learning_control_b200.engine import Engine
    from bench import make_policy_params
    B, T = 1 << 18, 24
    if waypoints:
        cfg = Q.EnvConfig.waypoint_eval(TJ.default_tables(0.5), auto_reset=Q.RESET_RESAMPLE, seed=3, battery=True)
    else:
        cfg = Q.EnvConfig.north_star(seed=3, max_episode_steps=17)
    outs = {}
    for form in ("2", "3", None):
        if form is None:
            monkeypatch.delenv("QS_TC_FORM", raising=False)
        else:
            monkeypatch.setenv("QS_TC_FORM", form)
        eng = Engine(cfg, B, device=0)
        st = eng.new_state(); eng.reset(st)
        if waypoints:
            p = torch.from_numpy(policies.pd_waypoint_policy(log_std=-3.5)).cuda()
        else:
            p = make_policy_params(eng, torch, torch.device("cuda"), seed=0)
        epi0 = st[26].view(torch.int32).clone()
        b = eng.rollout_policy(st, p, T=T, t0=0, dist=0, tensor_cores=True, bootstrap_gamma=0.9)
        ends = torch.maximum(b["done"], b["trunc"]).sum(0)
        b = eng.rollout_policy(st, p, T=T, t0=T, dist=0, tensor_cores=True, bootstrap_gamma=0.9, buffers=b)
        ends = ends + torch.maximum(b["done"], b["trunc"]).sum(0)
        torch.cuda.synchronize()
        if not waypoints:          # (a closed lap ends an episode without a done / truncated flag)
            assert torch.equal((st[26].view(torch.int32) - epi0).to(torch.float32), ends)
        outs[form] = ({k: v.clone() for k, v in b.items()}, st.clone())
        eng.close()
    ref_b, ref_st = outs["2"]
    for k in ("obs", "act", "logp", "value", "reward", "last_value"):
        assert bool(torch.isfinite(ref_b[k]).all()), k
    for form, (b, stn) in outs.items():
        for k in ref_b:
            assert torch.equal(b[k].view(torch.int32), ref_b[k].view(torch.int32)), (form, k)
        assert torch.equal(stn.view(torch.int32), ref_st.view(torch.int32)), form


def test_empty_and_bad_arguments_fail_loudly():
    import torch
    from uav_reinforcement_learning_control_b200.engine import Engine, QuadSimError
    with pytest.raises(QuadSimError):
        Engine(Q.EnvConfig.north_star(), 0, device=0)                     # empty batch
    eng = Engine(Q.EnvConfig.north_star(), 64, device=0)
    st = eng.new_state()
    with pytest.raises(QuadSimError):
        eng.step(st, torch.zeros(63, 4, device="cuda"))                   # ragged action batch
    with pytest.raises(QuadSimError):
        eng.step(st.double(), torch.zeros(64, 4, device="cuda"))          # wrong dtype
    with pytest.raises(QuadSimError):
        Engine(Q.EnvConfig.mjx_brax(auto_reset=Q.RESET_RESTORE_FIRST, episode_length=5), 8, device=0).step(
            torch.zeros(Q.NPLANES, 8, device="cuda"), torch.zeros(8, 4, device="cuda"))   # restore-first needs first_state
