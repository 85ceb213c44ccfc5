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

``bf16=True`` models the rounding points of the tcgen05 kernel (csrc/qs_ppo.cuh): every MMA operand -- normalised
observation, weights, relu(H1), relu(H2), the head gradient, the masked hidden gradients -- is rounded to bfloat16,
products and sums are exact-ish (float64 here, fp32 in TMEM), hidden biases enter as bf16 hi + lo pairs.
"""
from __future__ import annotations

import numpy as np

from .ppo_ref import A, H, LOG_SQRT_2PI, bf16_round, bias_hilo, unpack

PARAM_ORDER = ["aW1", "ab1", "aW2", "ab2", "aW3", "ab3", "cW1", "cb1", "cW2", "cb2", "cW3", "cb3", "log_std", "mean", "inv_std"]


def adv_normalise(adv):
    adv = np.asarray(adv, dtype=np.float64)
    n = adv.size
    std = np.sqrt(np.sum((adv - adv.mean()) ** 2) / (n - 1)) if n > 1 else 0.0
    return (adv - adv.mean()) / (std + 1e-8)


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


def grad(params, obs, act, old_logp, adv, ret, clip_range, vf_coef, ent_coef, normalize_adv=True, bf16=False):
    """-> (flat gradient in the packed layout, stats dict).  obs [n,12], act [n,4] raw samples, old_logp/adv/ret [n]."""
    obs_dim = np.asarray(obs).shape[1]
    pp = unpack(params, obs_dim, 0)
    n = len(obs)
    q = bf16_round if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    qb = bias_hilo if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    if bf16:
        x = q((np.asarray(obs, dtype=np.float32) - pp["mean"].astype(np.float32)) * pp["inv_std"].astype(np.float32))
    else:
        x = (np.asarray(obs, dtype=np.float64) - pp["mean"]) * pp["inv_std"]
    Ahat = adv_normalise(adv) if normalize_adv else np.asarray(adv, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    old_logp = np.asarray(old_logp, dtype=np.float64)
    ret = np.asarray(ret, dtype=np.float64)
    ls = pp["log_std"]
    inv_sig = np.exp(-ls)
    st = {}
    extra = {}

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

    W3a = np.zeros((H, 16)); W3a[:, :A] = pp["aW3"]
    W3c = np.zeros((H, 16)); W3c[:, 0] = pp["cW3"]
    _, ga = _mlp_fwd_bwd(x, pp["aW1"], pp["ab1"], pp["aW2"], pp["ab2"], W3a, actor_dout, q, qb)
    _, gc = _mlp_fwd_bwd(x, pp["cW1"], pp["cb1"], pp["cW2"], pp["cb2"], W3c, critic_dout, q, qb)
    s = 1.0 / n
    parts = {
        "aW1": ga["W1"] * s, "ab1": ga["b1"] * s, "aW2": ga["W2"] * s, "ab2": ga["b2"] * s,
        "aW3": ga["W3"][:, :A] * s, "ab3": ga["b3"][:A] * s,
        "cW1": gc["W1"] * s, "cb1": gc["b1"] * s, "cW2": gc["W2"] * s, "cb2": gc["b2"] * s,
        "cW3": gc["W3"][:, 0] * s, "cb3": gc["b3"][:1] * s,
        "log_std": extra["dls"] * s - ent_coef,
        "mean": np.zeros(obs_dim), "inv_std": np.zeros(obs_dim),
    }
    flat = np.concatenate([np.asarray(parts[k], dtype=np.float64).reshape(-1) for k in PARAM_ORDER])
    return flat, st


def split(flat, obs_dim=12):
    """Packed vector -> dict of named views (for per-tensor error reports)."""
    return unpack(flat, obs_dim, 0)


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
