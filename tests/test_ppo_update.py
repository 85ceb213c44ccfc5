"""PPO minibatch update: the hand-derived oracle against torch autograd / torch.optim.Adam (CPU), and the
sm_100a kernels (qs_ppo_grad: tcgen05 forward + backward, qs_ppo_adam) against the oracle (GPU)."""
import numpy as np
import pytest

from oracle import ppo_ref, ppo_update_ref as U

CLIP, VF, ENT = 0.1915, 0.5, 9.1e-5          # train.py:53-60


def _policy(seed, obs_dim=12):
    rng = np.random.default_rng(seed)
    parts = []
    for out in (4, 1):
        for (i, o) in ((obs_dim, 128), (128, 128), (128, out)):
            lim = (3.0 / i) ** 0.5
            parts += [rng.uniform(-lim, lim, i * o), rng.uniform(-0.1, 0.1, o)]
    parts.append(rng.uniform(-1.0, -0.3, 4))
    parts += [rng.uniform(-0.2, 0.2, obs_dim), rng.uniform(0.5, 1.5, obs_dim)]
    p = np.concatenate(parts).astype(np.float32)
    assert p.size == ppo_ref.param_count(obs_dim, 0)
    return p


def _batch(params, N, seed):
    """A synthetic rollout slice that is 'slightly off-policy' so that both clip branches are populated."""
    rng = np.random.default_rng(seed)
    obs = rng.uniform(-1, 1, (N, 12)).astype(np.float32)
    pp = ppo_ref.unpack(params, 12, 0)
    head, value = ppo_ref.forward(pp, obs)
    eps = rng.normal(size=(N, 4))
    act = (head + np.exp(pp["log_std"]) * eps).astype(np.float32)
    logp = np.sum(-0.5 * eps * eps - pp["log_std"] - ppo_ref.LOG_SQRT_2PI, axis=1)
    old_logp = (logp + rng.normal(scale=0.15, size=N)).astype(np.float32)      # ratio spread around 1 beyond the clip range
    # advantages that depend on the action noise and returns that depend on the observation, as in a real rollout: the
    # minibatch gradient then carries a signal instead of being the 1/sqrt(N) residue of cancelling per-sample terms
    adv = (2.0 * (eps @ np.array([1.0, -1.0, 0.5, 0.2])) + 0.5 * rng.normal(size=N) + 0.3).astype(np.float32)
    ret = (value + 0.5 + 0.5 * np.sin(3.0 * obs[:, 0]) + 0.3 * obs[:, 5] + 0.1 * rng.normal(size=N)).astype(np.float32)
    return obs, act, old_logp, adv, ret


def _torch_loss(ac, obs, act, old_logp, adv, ret, normalize=True):
    import torch
    a = adv
    if normalize:
        a = (a - a.mean()) / (a.std() + 1e-8)
    logp, value, ent = ac.evaluate(obs, act)
    ratio = torch.exp(logp - old_logp)
    pg = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - CLIP, 1 + CLIP)).mean()
    vl = torch.nn.functional.mse_loss(value, ret)
    return pg + VF * vl - ENT * ent, pg, vl


def _autograd_grad(params, batch, normalize=True):
    import torch
    from uav_reinforcement_learning_control_b200.ppo import ActorCritic
    ac = ActorCritic(12, "cpu")
    for group in (ac.actor, ac.critic, [ac.log_std]):
        for i, t in enumerate(group):
            group[i] = t.detach().double().requires_grad_(True)
    ac.log_std = ac.log_std.detach().double().requires_grad_(True)
    ac.obs_mean, ac.obs_inv_std = ac.obs_mean.double(), ac.obs_inv_std.double()
    ac.load_packed(torch.from_numpy(params.astype(np.float64)))
    t = [torch.from_numpy(np.asarray(b, dtype=np.float64)) for b in batch]
    loss, pg, vl = _torch_loss(ac, *t, normalize=normalize)
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in ac.parameters()]).numpy()
    return np.concatenate([g, np.zeros(24)]), float(pg.detach()), float(vl.detach())


@pytest.mark.parametrize("normalize", [True, False])
def test_oracle_gradient_matches_torch_autograd(normalize):
    params = _policy(1)
    batch = _batch(params, 777, 2)
    g_ref, pg, vl = _autograd_grad(params, batch, normalize)
    g, st = U.grad(params, *batch, CLIP, VF, ENT, normalize_adv=normalize)
    assert 0.05 < st["clip_frac"] < 0.95, st       # both branches of the clipped surrogate are exercised
    np.testing.assert_allclose(g, g_ref, rtol=1e-9, atol=1e-12)
    assert abs(st["pg_loss"] - pg) < 1e-12 and abs(st["v_loss"] - vl) < 1e-12


def test_oracle_adam_matches_torch():
    import torch
    rng = np.random.default_rng(3)
    P, n_train = 500, 470
    p0 = rng.normal(size=P); lr = 1.5e-4
    tp = torch.tensor(p0[:n_train], dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tp], lr=lr, eps=1e-5)
    p, m, v = p0.copy(), np.zeros(P), np.zeros(P)
    for step in range(1, 6):
        g = rng.normal(size=P) * (3.0 if step % 2 else 0.01)
        tp.grad = torch.tensor(g[:n_train] * 0.5, dtype=torch.float64)
        norm_t = float(torch.nn.utils.clip_grad_norm_([tp], 0.5))
        opt.step()
        p, m, v, norm = U.adam_step(p, g, m, v, step, lr, max_grad_norm=0.5, grad_scale=0.5, n_train=n_train)
        assert abs(norm - norm_t) < 1e-12
        np.testing.assert_allclose(p[:n_train], tp.detach().numpy(), rtol=1e-12, atol=1e-15)
        np.testing.assert_array_equal(p[n_train:], p0[n_train:])


def test_bf16_oracle_is_close_to_exact_oracle():
    """The rounding model of the tensor-core path perturbs the gradient by bf16-level noise only."""
    params = _policy(4)
    batch = _batch(params, 2048, 5)
    g, _ = U.grad(params, *batch, CLIP, VF, ENT)
    gb, _ = U.grad(params, *batch, CLIP, VF, ENT, bf16=True)
    a, b = U.split(g), U.split(gb)
    for k in ("aW1", "aW2", "aW3", "cW1", "cW2", "cW3", "ab1", "ab2", "cb1", "cb2", "log_std"):
        scale = np.abs(a[k]).max()
        # log_std's gradient, sum g (z^2 - 1), is a small residue of cancelling terms: looser
        assert np.abs(a[k] - b[k]).max() < (0.15 if k == "log_std" else 0.03) * scale, k


@pytest.mark.parametrize("n", [1, 2, 5, 128, 1000, 4096, 100003])
def test_oracle_feistel_permutation_is_a_permutation(n):
    p0 = U.feistel_permutation(n, 1234, 0)
    assert np.array_equal(np.sort(p0), np.arange(n))
    if n >= 128:
        p1 = U.feistel_permutation(n, 1234, 1)
        assert np.mean(p0 == p1) < 0.05 and np.mean(p0 == np.arange(n)) < 0.05        # epochs differ, no identity
        assert abs(np.corrcoef(p0, np.arange(n))[0, 1]) < 0.1


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,epoch", [(1, 0, 0), (5, 1, 2), (1000, 7, 0), (1 << 20, 99, 3), (8192 * 1024, (5 << 32) + 77, 1),
                                          (3000001, 123456789012345, 40)])
def test_fused_permutation_bit_exact(n, seed, epoch):
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    up = FusedUpdater("cuda:0")
    out = up.permutation(n, seed, epoch)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, U.feistel_permutation(n, seed, epoch))                # integer work: bit-exact
    assert np.array_equal(np.sort(got), np.arange(n))



def _report(g, g_ref, tol_rel, tol_log_std=None):
    a, b = U.split(g), U.split(g_ref)
    bad = []
    for k in U.PARAM_ORDER:
        scale = max(np.abs(b[k]).max(), 1e-12)
        err = np.abs(a[k] - b[k]).max()
        tol = tol_log_std if (k == "log_std" and tol_log_std is not None) else tol_rel
        if not err <= tol * scale + 1e-9:
            bad.append(f"{k}: max err {err:.3e} vs scale {scale:.3e}")
    assert not bad, "; ".join(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("N,n_idx,normalize", [(128, None, True), (1000, None, True), (77, None, False), (40000, None, True),
                                               (50000, 30011, True)])
def test_fused_gradient_matches_oracle(N, n_idx, normalize):
    """qs_ppo_grad vs the bf16-rounding oracle, tensor by tensor (a wrong UMMA descriptor shows up as ONE bad tensor)."""
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    params = _policy(10 + N % 7)
    batch = _batch(params, N, N)
    up = FusedUpdater("cuda:0")
    dev = [torch.from_numpy(b).cuda() for b in batch]
    idx = None
    sel = slice(None)
    if n_idx is not None:
        sel = np.random.default_rng(N).permutation(N)[:n_idx]
        idx = torch.from_numpy(sel.astype(np.int32)).cuda()
    for rep in range(2):                                  # second call: workspace re-armed, same answer
        g = up.grad(torch.from_numpy(params).cuda(), *dev, idx=idx, clip_range=CLIP, vf_coef=VF, ent_coef=ENT,
                    normalize_adv=normalize)
        torch.cuda.synchronize()
        g = g.cpu().numpy().astype(np.float64)
        sub = [b[sel] for b in batch]
        g_ref, st = U.grad(params, *sub, CLIP, VF, ENT, normalize_adv=normalize, bf16=True)
        n = len(sub[0])
        # fp32 accumulation order and __expf-level differences only: 2e-3 of each tensor's scale
        _report(g[:up.P], g_ref, 2e-3, tol_log_std=1e-2)
        stats = g[up.P:]
        assert stats[4] == n
        assert abs(stats[0] / n - st["pg_loss"]) < 2e-3 * max(1.0, abs(st["pg_loss"]))
        assert abs(stats[1] / n - st["v_loss"]) < 2e-3 * max(1.0, abs(st["v_loss"]))
        assert abs(stats[2] / n - st["clip_frac"]) < 5e-3
        assert abs(stats[3] / n - st["approx_kl"]) < 1e-3
        # and against exact arithmetic: bf16-level agreement (minibatches large enough for the gradient to be a mean)
        if n >= 1000:
            g_exact, _ = U.grad(params, *sub, CLIP, VF, ENT, normalize_adv=normalize)
            _report(g[:up.P], g_exact, 0.05, tol_log_std=0.2)


@pytest.mark.gpu
def test_fused_gradient_is_bitwise_reproducible():
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    params = _policy(3)
    batch = _batch(params, 100000, 8)
    up = FusedUpdater("cuda:0")
    dev = [torch.from_numpy(b).cuda() for b in batch]
    p = torch.from_numpy(params).cuda()
    g1 = up.grad(p, *dev, clip_range=CLIP, vf_coef=VF, ent_coef=ENT).clone()
    g2 = up.grad(p, *dev, clip_range=CLIP, vf_coef=VF, ent_coef=ENT).clone()
    assert torch.equal(g1, g2)


@pytest.mark.gpu
def test_fused_adam_matches_oracle():
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    rng = np.random.default_rng(0)
    params = _policy(5)
    up = FusedUpdater("cuda:0")
    P = up.P
    p_dev = torch.from_numpy(params).cuda()
    p, m, v = params.astype(np.float64), np.zeros(P), np.zeros(P)
    for step in range(1, 5):
        g = (rng.normal(size=P) * (1e-2 if step % 2 else 1e-4)).astype(np.float32)
        up.grad_buf[:P] = torch.from_numpy(g).cuda()
        norm_dev = up.adam(p_dev, lr=1.5478e-4, max_grad_norm=0.5, grad_scale=0.5)
        p, m, v, norm = U.adam_step(p, g, m, v, step, 1.5478e-4, max_grad_norm=0.5, grad_scale=0.5, n_train=P - 24)
        torch.cuda.synchronize()
        assert abs(float(norm_dev) - norm) < 1e-5 * norm
        np.testing.assert_allclose(p_dev.cpu().numpy(), p, rtol=2e-6, atol=2e-8)
    np.testing.assert_array_equal(p_dev.cpu().numpy()[P - 24:], params[P - 24:])       # the obs normaliser is untouched


@pytest.mark.gpu
def test_fused_update_follows_autograd_update():
    """Three full minibatch updates (same minibatches): fused kernels vs torch autograd + torch.optim.Adam on the GPU."""
    import torch
    from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater
    params = _policy(6)
    batch = _batch(params, 30000, 9)
    dev = [torch.from_numpy(b).cuda() for b in batch]
    up = FusedUpdater("cuda:0")
    p_fused = torch.from_numpy(params).cuda()
    ac = ActorCritic(12, "cuda:0")
    ac.load_packed(torch.from_numpy(params).cuda())
    opt = torch.optim.Adam(ac.parameters(), lr=3e-4, eps=1e-5)
    rng = np.random.default_rng(1)
    for it in range(3):
        idx = torch.from_numpy(rng.permutation(30000)[:10000].astype(np.int32)).cuda()
        up.grad(p_fused, *dev, idx=idx, clip_range=CLIP, vf_coef=VF, ent_coef=ENT)
        up.adam(p_fused, lr=3e-4, max_grad_norm=0.5)
        il = idx.long()
        loss, _, _ = _torch_loss(ac, *[d[il] for d in dev])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ac.parameters(), 0.5)
        opt.step()
    torch.cuda.synchronize()
    a, b = p_fused.cpu().numpy(), ac.pack().cpu().numpy()
    # Adam normalises the step to ~lr per element, so bf16-level gradient noise moves a parameter by at most a
    # fraction of 3 steps x lr
    assert np.abs(a - b).max() < 3 * 3e-4
    assert np.abs(a - params).max() > 2e-4            # ...and the parameters did move
    moved = np.abs(b - params) > 5e-4
    assert np.mean(np.sign(a - params)[moved] == np.sign(b - params)[moved]) > 0.98


@pytest.mark.gpu
def test_trainer_fused_improves_reward():
    """End to end on the device: tcgen05 rollout -> GAE -> fused update, a few iterations; the policy must improve."""
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine
    from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer
    torch.manual_seed(0)
    eng = Engine(Q.EnvConfig.north_star(seed=0), 16384, device=0)
    tr = PPOTrainer(eng, PPOConfig(n_steps=64, learning_rate=1e-3, ent_coef=0.0), seed=0)
    assert tr.fused and tr.tensor_cores
    tr.set_log_std(-1.0)
    log = tr.train(8)
    assert np.isfinite(log[-1]["mean_reward"]) and np.isfinite(log[-1]["pg_loss"])
    assert log[-1]["mean_reward"] > log[0]["mean_reward"]
    assert 0.0 <= log[-1]["clip_frac"] <= 1.0


@pytest.mark.gpu
def test_fused_update_respects_buffer_bounds():
    """compute-sanitizer is closed on the GPU pool, so bounds are checked by hand: every buffer the kernels write
    (workspace, gradient, Adam moments, parameters) is carved out of a canary-filled arena and the canaries must survive
    a ragged multi-tile update; the read-only inputs are exactly sized, so an out-of-range gather would fault."""
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    params = _policy(8)
    N = 33333
    batch = _batch(params, N, 12)
    dev = [torch.from_numpy(b).cuda() for b in batch]
    up = FusedUpdater("cuda:0")
    P, G = up.P, 4096                                        # G = guard floats between the carved buffers
    ws_floats = (up.workspace.numel() + 3) // 4
    sizes = {"ws": ws_floats, "grad": P + up.N_STATS, "m": P, "v": P, "params": P}
    canary = 1234567.0
    total = sum((n + G + 63) // 64 * 64 for n in sizes.values()) + G
    arena = torch.full((total,), canary, dtype=torch.float32, device="cuda")
    views, o = {}, G
    for k, n in sizes.items():
        views[k] = arena[o:o + n]
        o += (n + G + 63) // 64 * 64
    views["ws"].zero_(); views["grad"].zero_(); views["m"].zero_(); views["v"].zero_()
    views["params"].copy_(torch.from_numpy(params).cuda())
    up.workspace = views["ws"].view(torch.uint8)[:up.workspace.numel()]
    up.grad_buf, up.m, up.v = views["grad"], views["m"], views["v"]
    idx = up.permutation(N, 3, 0)[:20001].contiguous()
    for _ in range(2):
        up.grad(views["params"], *dev, idx=idx, clip_range=CLIP, vf_coef=VF, ent_coef=ENT)
        up.adam(views["params"], lr=1e-3)
    torch.cuda.synchronize()
    mask = torch.ones(total, dtype=torch.bool, device="cuda")
    o = G
    for k, n in sizes.items():
        mask[o:o + n] = False
        o += (n + G + 63) // 64 * 64
    assert bool((arena[mask] == canary).all()), "a PPO kernel wrote outside its buffers"
    assert torch.isfinite(views["params"]).all() and torch.isfinite(views["grad"]).all()


@pytest.mark.gpu
def test_trainer_fused_on_trajectory_follow_env_and_unsupported_policies():
    """The fused update serves every 12-D Gaussian-policy env (TrajectoryFollowEnv too) and the 21-D Brax policies; shapes
    outside {12, 21} x {Gaussian, tanh-normal} are refused loudly instead of being routed through some fallback."""
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine, QuadSimError
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater, PPOConfig, PPOTrainer
    eng = Engine(Q.EnvConfig.traj_gym(auto_reset=Q.RESET_RESAMPLE, seed=2), 4096, device=0)
    tr = PPOTrainer(eng, PPOConfig(n_steps=32, learning_rate=3e-4), seed=1)
    assert tr.fused
    tr.set_log_std(-1.0)
    before = tr.packed_params().clone()
    log = tr.train(3)
    assert all(np.isfinite(l["mean_reward"]) and np.isfinite(l["pg_loss"]) and np.isfinite(l["v_loss"]) for l in log)
    assert torch.isfinite(tr.packed_params()).all() and (tr.packed_params() != before).any()
    with pytest.raises(QuadSimError):
        FusedUpdater("cuda:0", obs_dim=20)
    with pytest.raises(QuadSimError):
        FusedUpdater("cuda:0", obs_dim=12, dist=2)
    eng21 = Engine(Q.EnvConfig.mjx_brax(), 256, device=0)
    with pytest.raises(ValueError):
        PPOTrainer(eng21, PPOConfig(n_steps=8), seed=0, fused=False)          # no silent torch fallback for the Brax policy
    assert PPOTrainer(eng21, PPOConfig(n_steps=8), seed=0).fused


@pytest.mark.gpu
def test_peer_memory_update_two_gpus():
    """qs_ppo_adam_peer (gradient exchange over NVLink peer memory inside the optimiser kernel) against the NCCL form, two
    ranks under torchrun: same parameters as reduce + all-reduce + Adam (bitwise at world 2, where the sum is commutative)
    and bitwise identical across the ranks.  Needs two GPUs on the box."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PPO_N="262144")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29577", os.path.join(root, "tools", "peer_update_check.py")],
                         env=env, capture_output=True, text=True, timeout=240)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-2000:]
    d = json.loads(lines[-1])
    assert d["world"] == 2 and d["peer_params_bitwise_in_sync_across_ranks"]
    assert d["max_abs_diff_peer_vs_nccl"] == 0.0 and d["max_abs_param_change"] > 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# Brax side of N4: 21-D observation, tanh-normal policy, brax's PPO loss and running observation
# normaliser.  The oracle extension is pinned to torch autograd on the CPU exactly like the SB3 one above.
# ---------------------------------------------------------------------------------------------------------------------
B_CLIP, B_VF, B_ENT = 0.3, 0.25, 1e-2            # brax ppo defaults: clipping_epsilon 0.3, v_loss 0.5 * 0.5, entropy_cost (train_brax_ppo.py:440-455)


def _policy_brax(seed, obs_dim=21):
    rng = np.random.default_rng(seed)
    parts = []
    for out in (8, 1):
        for (i, o) in ((obs_dim, 128), (128, 128), (128, out)):
            lim = (3.0 / i) ** 0.5
            parts += [rng.uniform(-lim, lim, i * o), rng.uniform(-0.1, 0.1, o)]
    parts += [rng.uniform(-0.2, 0.2, obs_dim), rng.uniform(0.5, 1.5, obs_dim)]
    p = np.concatenate(parts).astype(np.float32)
    assert p.size == ppo_ref.param_count(obs_dim, 1)
    return p


def _batch_brax(params, N, seed, obs_dim=21):
    rng = np.random.default_rng(seed)
    obs = rng.uniform(-1, 1, (N, obs_dim)).astype(np.float32)
    pp = ppo_ref.unpack(params, obs_dim, 1)
    head, value = ppo_ref.forward(pp, obs)
    eps = rng.normal(size=(N, 4))
    raw, _, logp = ppo_ref.sample(pp, head, eps, 1)
    act = raw.astype(np.float32)
    old_logp = (logp + rng.normal(scale=0.2, size=N)).astype(np.float32)
    adv = (2.0 * (eps @ np.array([1.0, -1.0, 0.5, 0.2])) + 0.5 * rng.normal(size=N) + 0.3).astype(np.float32)
    ret = (value + 0.5 + 0.5 * np.sin(3.0 * obs[:, 0]) + 0.3 * obs[:, 5] + 0.1 * rng.normal(size=N)).astype(np.float32)
    return obs, act, old_logp, adv, ret


def _autograd_grad_brax(params, batch, eps_ent, obs_dim=21, normalize=True):
    """brax compute_ppo_loss on precomputed advantages, written with torch ops (float64) and differentiated by autograd."""
    import torch
    pp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=k not in ("mean", "inv_std")) for k, v in ppo_ref.unpack(params, obs_dim, 1).items()}
    obs, act, old_logp, adv, ret = [torch.from_numpy(np.asarray(b, dtype=np.float64)) for b in batch]
    eps_ent = torch.from_numpy(np.asarray(eps_ent, dtype=np.float64))
    x = (obs - pp["mean"]) * pp["inv_std"]
    h = torch.relu(x @ pp["aW1"] + pp["ab1"]); h = torch.relu(h @ pp["aW2"] + pp["ab2"]); head = h @ pp["aW3"] + pp["ab3"]
    c = torch.relu(x @ pp["cW1"] + pp["cb1"]); c = torch.relu(c @ pp["cW2"] + pp["cb2"]); value = c @ pp["cW3"] + pp["cb3"][0]
    loc, scale = head[:, :4], torch.nn.functional.softplus(head[:, 4:]) + 0.001
    dist = torch.distributions.Normal(loc, scale)
    ldj = lambda t: 2.0 * (np.log(2.0) - t - torch.nn.functional.softplus(-2.0 * t))
    logp = (dist.log_prob(act) - ldj(act)).sum(-1)
    a = adv
    if normalize:
        a = (a - a.mean()) / (a.std(unbiased=False) + 1e-8)
    rho = torch.exp(logp - old_logp)
    pg = -torch.min(rho * a, torch.clamp(rho, 1 - B_CLIP, 1 + B_CLIP) * a).mean()
    vl = ((ret - value) ** 2).mean()
    ent = (dist.entropy() + ldj(loc + scale * eps_ent)).sum(-1).mean()
    (pg + B_VF * vl - B_ENT * ent).backward()
    order = ["aW1", "ab1", "aW2", "ab2", "aW3", "ab3", "cW1", "cb1", "cW2", "cb2", "cW3", "cb3"]
    g = np.concatenate([pp[k].grad.reshape(-1).numpy() for k in order] + [np.zeros(2 * obs_dim)])
    return g, float(pg.detach()), float(vl.detach()), float(ent.detach())


@pytest.mark.parametrize("obs_dim", [21, 12])
def test_oracle_brax_gradient_matches_torch_autograd(obs_dim):
    params = _policy_brax(11, obs_dim)
    batch = _batch_brax(params, 600, 12, obs_dim)
    eps_ent = U.entropy_noise(77, np.arange(600))
    assert abs(eps_ent.mean()) < 0.1 and abs(eps_ent.std() - 1.0) < 0.1
    g_ref, pg, vl, ent = _autograd_grad_brax(params, batch, eps_ent, obs_dim)
    g, st = U.grad(params, *batch, B_CLIP, B_VF, B_ENT, dist=1, eps_entropy=eps_ent)
    assert 0.05 < st["clip_frac"] < 0.95, st
    np.testing.assert_allclose(g, g_ref, rtol=1e-9, atol=1e-12)
    assert abs(st["pg_loss"] - pg) < 1e-12 and abs(st["v_loss"] - vl) < 1e-12 and abs(st["entropy"] - ent) < 1e-12


def test_oracle_running_obs_stats_match_direct_moments():
    rng = np.random.default_rng(5)
    state = (0.0, np.zeros(21), np.zeros(21))
    chunks = [rng.normal(loc=3.0, scale=[0.1 + 0.2 * k for k in range(21)], size=(n, 21)) for n in (1000, 37, 5000)]
    for c in chunks:
        state, mean, inv_std = U.running_obs_stats(state, c)
    allx = np.concatenate(chunks)
    np.testing.assert_allclose(mean, allx.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(1.0 / inv_std, allx.std(axis=0), rtol=1e-10)
    # a constant feature clips at std_min instead of dividing by zero
    st2, m2, i2 = U.running_obs_stats((0.0, np.zeros(2), np.zeros(2)), np.ones((10, 2)))
    assert np.all(i2 == 1e6)


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dim,dist,n", [(21, 1, 4096), (21, 1, 1000), (12, 1, 2048), (21, 0, 2048)])
def test_fused_gradient_generic_policies_match_oracle(obs_dim, dist, n):
    """qs_ppo_grad for the Brax policy family (21-D obs, tanh-normal head) and the mixed variants, through
    ppo_grad_tc_kernel<D, DIST> (qs_ppo_generic.cuh), against the bf16-operand oracle: 2e-3 per tensor."""
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    params = _policy_brax(21, obs_dim) if dist == 1 else _policy(21, obs_dim)
    if dist == 1:
        batch = _batch_brax(params, n, 22, obs_dim)
    else:
        rng = np.random.default_rng(22)
        obs = rng.uniform(-1, 1, (n, obs_dim)).astype(np.float32)
        pp = ppo_ref.unpack(params, obs_dim, 0)
        head, value = ppo_ref.forward(pp, obs)
        eps = rng.normal(size=(n, 4))
        act = (head + np.exp(pp["log_std"]) * eps).astype(np.float32)
        logp = np.sum(-0.5 * eps * eps - pp["log_std"] - ppo_ref.LOG_SQRT_2PI, axis=1)
        batch = (obs, act, (logp + rng.normal(scale=0.15, size=n)).astype(np.float32),
                 (2.0 * (eps @ np.array([1.0, -1.0, 0.5, 0.2])) + 0.3).astype(np.float32), (value + 0.5).astype(np.float32))
    up = FusedUpdater("cuda:0", obs_dim=obs_dim, dist=dist)
    dev = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in batch]
    clip, vf, ent = (B_CLIP, B_VF, B_ENT) if dist == 1 else (CLIP, VF, ENT)
    g = up.grad(torch.from_numpy(params).cuda(), *dev, clip_range=clip, vf_coef=vf, ent_coef=ent,
                normalize_adv=2 if dist == 1 else 1, sample_seed=1234)
    torch.cuda.synchronize()
    eps_ent = U.entropy_noise(1234, np.arange(n)) if dist == 1 else None
    g_ref, st = U.grad(params, *batch, clip, vf, ent, bf16=True, dist=dist, eps_entropy=eps_ent)
    got = g.cpu().numpy().astype(np.float64)
    a, b = U.split(got[:up.P], obs_dim, dist), U.split(g_ref, obs_dim, dist)
    for k in U.PARAM_ORDER:
        if k not in b or b[k].size == 0:
            continue
        tol = (1e-2 if k == "log_std" else 2e-3) * max(np.abs(b[k]).max(), 1e-12) + 1e-9
        assert np.abs(a[k] - b[k]).max() <= tol, (k, np.abs(a[k] - b[k]).max(), np.abs(b[k]).max())
    stats = got[up.P:]
    assert abs(stats[0] / n - st["pg_loss"]) < 2e-3 and abs(stats[1] / n - st["v_loss"]) < 2e-3 * max(1.0, st["v_loss"])
    assert stats[4] == n and abs(stats[2] / n - st["clip_frac"]) < 5e-3
    if dist == 1:
        assert abs(stats[5] / n - st["entropy"]) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dim", [21, 12])
def test_running_obs_normaliser_kernel_matches_oracle(obs_dim):
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    up = FusedUpdater("cuda:0", obs_dim=obs_dim, dist=1)
    params = torch.from_numpy(_policy_brax(3, obs_dim)).cuda()
    rng = np.random.default_rng(8)
    state = (0.0, np.zeros(obs_dim), np.zeros(obs_dim))
    for n in (5000, 777, 200000):
        x = (rng.normal(loc=np.linspace(-3, 3, obs_dim), scale=np.linspace(0.05, 4.0, obs_dim), size=(n, obs_dim))).astype(np.float32)
        x[:, 3] = 1.0                                   # a constant feature (the quaternion w of a level drone): std clips at 1e-6
        up.update_obs_stats(params, torch.from_numpy(x).cuda())
        state, mean, inv_std = U.running_obs_stats(state, x)
        torch.cuda.synchronize()
        got = params.cpu().numpy()
        L = ppo_ref.param_count(obs_dim, 1)
        np.testing.assert_allclose(got[L - 2 * obs_dim:L - obs_dim], mean, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(got[L - obs_dim:], inv_std, rtol=1e-5)
    assert got[L - obs_dim + 3] == np.float32(1e6)


@pytest.mark.gpu
def test_trainer_runs_fused_on_mjx_brax():
    """PPOTrainer on the MJX-parity env (JaxMJXQuadBraxEnv semantics, 21-D obs, Episode + AutoReset wrappers) with brax's
    reference hyper-parameters: tcgen05 rollout -> brax GAE -> running obs normaliser -> fused tanh-normal PPO update.
    The update must be the fused kernels (no torch autograd), keep everything finite and actually move the policy."""
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine
    from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer
    eng = Engine(Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST, seed=3), 2048, device=0)
    tr = PPOTrainer(eng, PPOConfig.brax_reference(), seed=0)
    assert tr.fused and tr.brax and tr.updater.dist == 1 and tr.updater.obs_dim == 21 and tr.policy is None
    p0 = tr.params.clone()
    launches0 = eng.launch_count()
    log = tr.train(3)
    torch.cuda.synchronize()
    assert eng.launch_count() - launches0 >= 3 * (1 + 1 + 1 + 4 * 16 * 2)       # rollout, GAE, obs stats, 64 x (grad, adam) per iteration
    assert torch.isfinite(tr.params).all()
    L = tr.params.numel()
    assert (tr.params[:L - 42] != p0[:L - 42]).float().mean() > 0.9               # the weights moved
    mean, inv_std = tr.params[L - 42:L - 21].cpu().numpy(), tr.params[L - 21:].cpu().numpy()
    assert 0.5 < mean[2] < 4.0 and np.all(np.abs(mean[3:7]) <= 1.0)               # z inside the bounds; an untrained policy tumbles, |quat| <= 1
    assert (inv_std > 0).all() and np.isfinite(inv_std).all() and (inv_std < 1e6).all()     # every feature varies in a rollout
    assert float(tr.updater.obs_running[0].item()) == 3 * 10 * 2048
    for s in log:
        assert np.isfinite(s["pg_loss"]) and np.isfinite(s["v_loss"]) and 0.0 <= s["clip_frac"] <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["hover_gym", "mjx_brax"])
def test_native_epoch_loop_is_bitwise_the_per_minibatch_loop(mode):
    """qs_ppo_update_epoch (one native call per epoch: statistics accumulated inside the optimiser kernel, minibatch
    slicing, seeds and optimiser steps counted in C) against the same update issued minibatch by minibatch from Python:
    parameters, Adam moments and the reported statistics must agree bit for bit, for both learners, incl. a minibatch
    count that does not divide the rollout (the last minibatch takes the remainder rows)."""
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine
    from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer
    out = []
    for native in (True, False):
        if mode == "hover_gym":
            eng = Engine(Q.EnvConfig.north_star(seed=2), 1000, device=0)
            cfg = PPOConfig(n_steps=16, n_epochs=3, num_minibatches=7, learning_rate=1e-3)
        else:
            eng = Engine(Q.EnvConfig.mjx_brax(episode_length=500, auto_reset=Q.RESET_RESTORE_FIRST, seed=3), 1024, device=0)
            cfg = PPOConfig.brax_reference()
        tr = PPOTrainer(eng, cfg, seed=1, native_epochs=native)
        assert tr.fused and tr.native_epochs == native
        log = tr.train(2)
        torch.cuda.synchronize()
        out.append((tr.params.clone(), tr.updater.m.clone(), tr.updater.v.clone(), tr.updater.step, log))
    (pa, ma, va, sa, la), (pb, mb_, vb, sb, lb) = out
    assert sa == sb and sa == 2 * cfg.n_epochs * cfg.num_minibatches
    assert torch.equal(pa, pb) and torch.equal(ma, mb_) and torch.equal(va, vb)
    for x, y in zip(la, lb):
        for k in ("pg_loss", "v_loss", "clip_frac", "approx_kl"):
            assert x[k] == y[k], (k, x[k], y[k])


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dim,dist,n,N", [(12, 0, 4096, 20000), (12, 0, 1000, 1000), (21, 1, 2048, 9000)])
def test_packed_rows_give_the_same_gradient_bitwise(obs_dim, dist, n, N):
    """qs_ppo_pack + qs_ppo_grad_packed (one 128-byte line per sample) against qs_ppo_grad on the five separate arrays, same
    minibatch indices: the kernels see identical numbers in identical order, so the gradients are bitwise equal."""
    import torch
    from uav_reinforcement_learning_control_b200.ppo import FusedUpdater
    params = _policy_brax(5, obs_dim) if dist == 1 else _policy(5, obs_dim)
    rng = np.random.default_rng(6)
    obs = rng.uniform(-1, 1, (N, obs_dim)).astype(np.float32); act = rng.normal(size=(N, 4)).astype(np.float32)
    logp = rng.normal(size=N).astype(np.float32) - 3; adv = rng.normal(size=N).astype(np.float32); ret = rng.normal(size=N).astype(np.float32)
    up = FusedUpdater("cuda:0", obs_dim=obs_dim, dist=dist)
    dev = [torch.from_numpy(a).cuda() for a in (obs, act, logp, adv, ret)]
    d_params = torch.from_numpy(params).cuda()
    idx = torch.from_numpy(rng.permutation(N)[:n].astype(np.int32)).cuda()
    packed = up.pack(*dev)
    torch.cuda.synchronize()
    pk = packed.cpu().numpy()
    np.testing.assert_array_equal(pk[:, :obs_dim], obs); np.testing.assert_array_equal(pk[:, obs_dim:obs_dim + 4], act)
    np.testing.assert_array_equal(pk[:, obs_dim + 4], logp); np.testing.assert_array_equal(pk[:, obs_dim + 5], adv)
    np.testing.assert_array_equal(pk[:, obs_dim + 6], ret); assert (pk[:, obs_dim + 7:] == 0).all()
    kw = dict(clip_range=0.2, vf_coef=0.5, ent_coef=1e-3, normalize_adv=2 if dist else 1, sample_seed=9)
    g0 = up.grad(d_params, *dev, idx=idx, **kw).clone()
    g1 = up.grad(d_params, adv=dev[3], idx=idx, packed=packed, **kw).clone()
    torch.cuda.synchronize()
    assert torch.equal(g0.view(torch.int32), g1.view(torch.int32))
