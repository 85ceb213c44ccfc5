"""ORACLE (test infrastructure, not product code): one PPO minibatch update in NumPy float64.

Restates what the caller of the rollout does per minibatch -- SB3 ``PPO.train`` as configured by the reference
(train.py:50-68: MlpPolicy pi=[128,128] vf=[128,128] ReLU, clip_range, ent_coef, vf_coef, max_grad_norm 0.5,
normalize_advantage, Adam lr / eps 1e-5); stable-baselines3 is not vendored or installable here, semantics are from its
published source:

    A      = (adv - mean(adv)) / (std(adv) + 1e-8)                      (unbiased std, per minibatch)
    ratio  = exp(logp(a | obs) - old_logp)
    loss   = -mean(min(ratio A, clip(ratio, 1-c, 1+c) A)) + vf_coef mean((ret - V(obs))^2) - ent_coef H
    H      = sum_k (0.5 + 0.5 log 2 pi + log_std_k)
    g      <- clip_by_global_norm(grad loss, max_grad_norm);   Adam step (torch.optim.Adam semantics)

The backward pass is written out by hand (no autograd here); tests/test_ppo_update.py additionally checks it against
torch autograd of uav_reinforcement_learning_control_b200.ppo.ActorCritic.evaluate.

``dist=1`` is the loss of brax ``ppo_train.train`` as the reference configures it (train_brax_ppo.py:589-620: tanh-normal
policy over the 21-D raw observation; brax.training.agents.ppo.losses.compute_ppo_loss, third party, restated from its
published source) on precomputed advantages / value targets:

    scale  = softplus(raw_scale) + 0.001,   logp = sum_k [N(a_raw; loc, scale) - ldj(a_raw)],  ldj(x) = 2 (log 2 - x - softplus(-2x))
    A      = (adv - mean(adv)) / (std(adv) + 1e-8)                      (population std, jnp.std)
    loss   = -mean(min(rho A, clip(rho, 1-c, 1+c) A)) + vf_coef mean((ret - V)^2) - ent_coef mean(H)     (brax: vf_coef = 0.25)
    H_i    = sum_k [1/2 + log sqrt(2 pi) + log scale_k + ldj(loc_k + scale_k eps_ik)]      eps: one fresh N(0,1) draw per row

``bf16=True`` models the rounding points of the tcgen05 kernel (csrc/qs_ppo.cuh): every MMA operand -- normalised
observation, weights, relu(H1), relu(H2), the head gradient, the masked hidden gradients -- is rounded to bfloat16,
products and sums are exact-ish (float64 here, fp32 in TMEM), hidden biases enter as bf16 hi + lo pairs.
"""
from __future__ import annotations

import numpy as np

from .ppo_ref import A, H, LOG_SQRT_2PI, bf16_round, bias_hilo, unpack

PARAM_ORDER = ["aW1", "ab1", "aW2", "ab2", "aW3", "ab3", "cW1", "cb1", "cW2", "cb2", "cW3", "cb3", "log_std", "mean", "inv_std"]


def adv_normalise(adv, ddof=1):
    adv = np.asarray(adv, dtype=np.float64)
    n = adv.size
    std = np.sqrt(np.sum((adv - adv.mean()) ** 2) / (n - ddof)) if n > ddof else 0.0
    return (adv - adv.mean()) / (std + 1e-8)


def softplus(x):
    return np.logaddexp(0.0, x)


def entropy_noise(sample_seed, rows):
    """The four standard normals the kernel draws for the entropy term of row j: Philox4x32-10 with key
    (sample_seed, 0x5eed0ea7), counter (j, 0, 0, 3), Box-Muller pairs (csrc/qs_ppo_generic.cuh)."""
    from . import philox
    rows = np.asarray(rows, dtype=np.uint32)
    ctr = np.stack([rows, np.zeros_like(rows), np.zeros_like(rows), np.full_like(rows, 3)], axis=1)
    raw = philox.philox4x32_10(ctr, np.array([int(sample_seed) & 0xFFFFFFFF, 0x5eed0ea7], dtype=np.uint32))
    e0, e1 = philox.normal_pair(raw[:, 0], raw[:, 1])
    e2, e3 = philox.normal_pair(raw[:, 2], raw[:, 3])
    return np.stack([e0, e1, e2, e3], axis=1)


def _mlp_fwd_bwd(x, W1, b1, W2, b2, W3, dout_fn, q, qb):
    """x [n, D] (already rounded); returns (out, grads dict) with dout = dout_fn(out) the loss gradient at the head."""
    h1p = x @ q(W1) + qb(b1)
    h1 = q(np.maximum(h1p, 0.0))
    h2p = h1 @ q(W2) + qb(b2)
    h2 = q(np.maximum(h2p, 0.0))
    out = h2 @ q(W3)
    dout = dout_fn(out)                       # float64, unrounded (bias / log_std sums use this)
    dq = q(dout)
    d2 = q((dq @ q(W3).T) * (h2 > 0))
    d1 = q((d2 @ q(W2).T) * (h1 > 0))
    g = {"W3": h2.T @ dq, "b3": dout.sum(0), "W2": h1.T @ d2, "b2": d2.sum(0), "W1": x.T @ d1, "b1": d1.sum(0)}
    return out, g


def grad(params, obs, act, old_logp, adv, ret, clip_range, vf_coef, ent_coef, normalize_adv=True, bf16=False, dist=0,
         ddof=None, eps_entropy=None):
    """-> (flat gradient in the packed layout, stats dict).  obs [n, D], act [n,4] raw samples, old_logp/adv/ret [n].
    dist 1: eps_entropy [n, 4] is the noise of the entropy sample (entropy_noise(sample_seed, rows) for the kernel's)."""
    obs_dim = np.asarray(obs).shape[1]
    pp = unpack(params, obs_dim, dist)
    ddof = (1 if dist == 0 else 0) if ddof is None else ddof
    n = len(obs)
    q = bf16_round if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    qb = bias_hilo if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    if bf16:
        x = q((np.asarray(obs, dtype=np.float32) - pp["mean"].astype(np.float32)) * pp["inv_std"].astype(np.float32))
    else:
        x = (np.asarray(obs, dtype=np.float64) - pp["mean"]) * pp["inv_std"]
    Ahat = adv_normalise(adv, ddof) if normalize_adv else np.asarray(adv, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    old_logp = np.asarray(old_logp, dtype=np.float64)
    ret = np.asarray(ret, dtype=np.float64)
    ls = pp["log_std"] if dist == 0 else None
    inv_sig = np.exp(-ls) if dist == 0 else None
    st = {}
    extra = {}
    Ao = A if dist == 0 else 2 * A

    def surrogate(logp):
        lr = logp - old_logp
        ratio = np.exp(lr)
        lo, hi = 1.0 - clip_range, 1.0 + clip_range
        unclipped, clipped = Ahat * ratio, Ahat * np.clip(ratio, lo, hi)
        inside = (ratio >= lo) & (ratio <= hi)
        active = inside | (unclipped < clipped)
        st["pg_loss"] = float(np.mean(-np.minimum(unclipped, clipped)))
        st["clip_frac"] = float(np.mean(~inside))
        st["approx_kl"] = float(np.mean((ratio - 1.0) - lr))
        return np.where(active, -Ahat * ratio, 0.0)          # d loss_i / d logp_i (before the 1/n)

    def actor_dout_tanh(out):
        loc = out[:, :A] + pp["ab3"][:A]
        rs = out[:, A:2 * A] + pp["ab3"][A:]
        scale = softplus(rs) + 0.001
        sig = 1.0 / (1.0 + np.exp(-rs))
        z = (act - loc) / scale
        ldj_a = 2.0 * (np.log(2.0) - act - softplus(-2.0 * act))
        logp = np.sum(-0.5 * z * z - np.log(scale) - LOG_SQRT_2PI - ldj_a, axis=1)
        g = surrogate(logp)
        xs = loc + scale * eps_entropy
        th = np.tanh(xs)
        ent = np.sum(0.5 + LOG_SQRT_2PI + np.log(scale) + 2.0 * (np.log(2.0) - xs - softplus(-2.0 * xs)), axis=1)
        st["entropy"] = float(np.mean(ent))
        d = np.zeros_like(out)
        d[:, :A] = g[:, None] * z / scale + (-ent_coef) * (-2.0 * th)
        d[:, A:2 * A] = (g[:, None] * (z * z - 1.0) / scale + (-ent_coef) * (1.0 / scale - 2.0 * th * eps_entropy)) * sig
        return d

    def actor_dout(out):
        mean = out[:, :A] + pp["ab3"]
        z = (act - mean) * inv_sig
        logp = np.sum(-0.5 * z * z - ls - LOG_SQRT_2PI, axis=1)
        lr = logp - old_logp
        ratio = np.exp(lr)
        lo, hi = 1.0 - clip_range, 1.0 + clip_range
        unclipped, clipped = Ahat * ratio, Ahat * np.clip(ratio, lo, hi)
        inside = (ratio >= lo) & (ratio <= hi)
        active = inside | (unclipped < clipped)
        g = np.where(active, -Ahat * ratio, 0.0)
        st["pg_loss"] = float(np.mean(-np.minimum(unclipped, clipped)))
        st["clip_frac"] = float(np.mean(~inside))
        st["approx_kl"] = float(np.mean((ratio - 1.0) - lr))
        extra["dls"] = np.sum(g[:, None] * (z * z - 1.0), axis=0)
        d = np.zeros_like(out)
        d[:, :A] = g[:, None] * z * inv_sig
        return d

    def critic_dout(out):
        v = out[:, 0] + pp["cb3"][0]
        err = v - ret
        st["v_loss"] = float(np.mean(err * err))
        d = np.zeros_like(out)
        d[:, 0] = 2.0 * vf_coef * err
        return d

    W3a = np.zeros((H, 16)); W3a[:, :Ao] = pp["aW3"]
    W3c = np.zeros((H, 16)); W3c[:, 0] = pp["cW3"]
    if dist == 1 and eps_entropy is None:
        eps_entropy = np.zeros((n, A))
    _, ga = _mlp_fwd_bwd(x, pp["aW1"], pp["ab1"], pp["aW2"], pp["ab2"], W3a, actor_dout if dist == 0 else actor_dout_tanh, q, qb)
    _, gc = _mlp_fwd_bwd(x, pp["cW1"], pp["cb1"], pp["cW2"], pp["cb2"], W3c, critic_dout, q, qb)
    s = 1.0 / n
    parts = {
        "aW1": ga["W1"] * s, "ab1": ga["b1"] * s, "aW2": ga["W2"] * s, "ab2": ga["b2"] * s,
        "aW3": ga["W3"][:, :Ao] * s, "ab3": ga["b3"][:Ao] * s,
        "cW1": gc["W1"] * s, "cb1": gc["b1"] * s, "cW2": gc["W2"] * s, "cb2": gc["b2"] * s,
        "cW3": gc["W3"][:, 0] * s, "cb3": gc["b3"][:1] * s,
        "log_std": (extra["dls"] * s - ent_coef) if dist == 0 else np.zeros(0),
        "mean": np.zeros(obs_dim), "inv_std": np.zeros(obs_dim),
    }
    flat = np.concatenate([np.asarray(parts[k], dtype=np.float64).reshape(-1) for k in PARAM_ORDER])
    return flat, st


def split(flat, obs_dim=12, dist=0):
    """Packed vector -> dict of named views (for per-tensor error reports)."""
    return unpack(flat, obs_dim, dist)


def running_obs_stats(state, obs, std_min=1e-6, std_max=1e6):
    """brax.training.acme.running_statistics.update + the normaliser it implies (third party, restated): merge the batch
    into state = (count, mean[D], summed_variance[D]); returns (state', mean, inv_std) with
    std = clip(sqrt(max(summed_variance / count, 0)), std_min, std_max)."""
    count, mean, m2 = state
    obs = np.asarray(obs, dtype=np.float64)
    n = obs.shape[0]
    tot = count + n
    diff0 = obs - mean
    mean_new = mean + diff0.sum(axis=0) / tot
    m2_new = m2 + (diff0 * (obs - mean_new)).sum(axis=0)
    std = np.clip(np.sqrt(np.maximum(m2_new / tot, 0.0)), std_min, std_max)
    return (tot, mean_new, m2_new), mean_new, 1.0 / std


def adam_step(params, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-5, max_grad_norm=0.5, grad_scale=1.0, n_train=None):
    """torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the first n_train entries.  Returns (params, m, v, norm)."""
    p = np.array(params, dtype=np.float64); m = np.array(m, dtype=np.float64); v = np.array(v, dtype=np.float64)
    n_train = p.size if n_train is None else n_train
    gg = np.asarray(g, dtype=np.float64)[:n_train] * grad_scale
    norm = float(np.sqrt(np.sum(gg * gg)))
    if max_grad_norm > 0:
        gg = gg * min(1.0, max_grad_norm / (norm + 1e-6))
    m[:n_train] = beta1 * m[:n_train] + (1 - beta1) * gg
    v[:n_train] = beta2 * v[:n_train] + (1 - beta2) * gg * gg
    step_size = lr / (1 - beta1 ** step)
    denom = np.sqrt(v[:n_train]) / np.sqrt(1 - beta2 ** step) + eps
    p[:n_train] -= step_size * m[:n_train] / denom
    return p, m, v, norm


# --------------------------------------------------------------------------------------------------------------------
# Minibatch shuffle: keyed Feistel bijection + cycle walking (csrc/qs_ppo.cuh: feistel_index, quadsim.cu:
# qs_ppo_permutation).  uint32 arithmetic restated with NumPy; bit-exact parity is tested.
def _mix(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13; x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def feistel_permutation(n, seed, epoch):
    """out[i] for i in 0..n-1: SB3's per-epoch ``np.random.permutation`` (RolloutBuffer.get) as a keyed bijection."""
    bits = 1
    while bits < 31 and (1 << bits) < n:
        bits += 1
    h = max((bits + 1) // 2, 1)
    mask = (1 << h) - 1
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0 = (int(_mix((seed & 0xFFFFFFFF) ^ 0xA511E9B3)) + epoch * 0x9E3779B9) & 0xFFFFFFFF
    k1 = int(_mix(((seed >> 32) + 0x632BE5AB) & 0xFFFFFFFF)) ^ int(_mix((epoch + 0x85157AF5) & 0xFFFFFFFF))
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        v = x[todo]
        l, r = v >> h, v & mask
        for rnd in range(4):
            f = _mix((r * 0x9E3779B1 + k0 + rnd * 0x7F4A7C15) & 0xFFFFFFFF) ^ _mix((k1 + rnd) & 0xFFFFFFFF)
            l, r = r, l ^ (f & mask)
        v = (l << h) | r
        x[todo] = v
        todo[todo] = v >= n
    return x.astype(np.int64)
