"""ORACLE (test infrastructure, not product code): policy forward, sampling and GAE in NumPy float64.

Restates the third-party pieces that sit around the env in the reference's two trainers
(neither brax nor stable-baselines3 is vendored or installable here; semantics are from their
published source, call sites cited):

* SB3 ``MlpPolicy`` with ``net_arch=dict(pi=[128,128], vf=[128,128])``, ReLU (train.py:61-64):
  separate actor / critic MLPs, ``DiagGaussianDistribution`` with a state-independent log_std,
  actions clipped to [-1, 1] only when sent to the env, log-prob of the unclipped sample.
* Brax ``make_ppo_networks(policy_hidden_layer_sizes=(128,128), value_hidden_layer_sizes=(128,128),
  activation=relu)`` with ``NormalTanhDistribution`` (train_brax_ppo.py:605-612): the actor emits
  loc | raw_scale, scale = softplus(raw_scale) + 0.001, action = tanh(loc + scale * eps).
* GAE: SB3 ``RolloutBuffer.compute_returns_and_advantage`` (train.py:57-58 gamma/lambda) and brax
  ``compute_gae`` (train_brax_ppo.py:442,451).

Parameter packing = the layout documented at qs_policy_param_count in include/quadsim_abi.h.
"""
from __future__ import annotations

import numpy as np

from . import philox

H, A = 128, 4
LOG_SQRT_2PI = 0.9189385332046727


def unpack(params, obs_dim, dist):
    p = np.asarray(params, dtype=np.float64)
    Ao = 2 * A if dist == 1 else A
    o = 0
    out = {}

    def take(name, *shape):
        nonlocal o
        n = int(np.prod(shape))
        out[name] = p[o:o + n].reshape(shape)
        o += n
    take("aW1", obs_dim, H); take("ab1", H); take("aW2", H, H); take("ab2", H); take("aW3", H, Ao); take("ab3", Ao)
    take("cW1", obs_dim, H); take("cb1", H); take("cW2", H, H); take("cb2", H); take("cW3", H); take("cb3", 1)
    if dist == 0:
        take("log_std", A)
    take("mean", obs_dim); take("inv_std", obs_dim)
    assert o == p.size, (o, p.size)
    return out


def param_count(obs_dim, dist):
    Ao = 2 * A if dist == 1 else A
    return 2 * (obs_dim * H + H + H * H + H) + H * Ao + Ao + H + 1 + (A if dist == 0 else 0) + 2 * obs_dim


def bf16_round(a):
    """Round-to-nearest-even to bfloat16 precision (returned as float64)."""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def bias_hilo(b):
    """Bias as the tcgen05 path feeds it to the MMA: a bf16 high part plus a bf16 low part (two K slots)."""
    hi = bf16_round(b)
    return hi + bf16_round(np.asarray(b, dtype=np.float64) - hi)


def forward(pp, obs, bf16=False):
    """bf16=True models the tcgen05 path: every MMA operand (activations and weights of all three layers)
    rounded to bfloat16, products and sums in (at least) float32; the layer-1 and layer-2 biases ride inside
    the MMAs as bf16 hi + lo pairs (bias_hilo), the head biases are added in float32."""
    q = bf16_round if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    qb = bias_hilo if bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    x = (np.asarray(obs, dtype=np.float32) - pp["mean"].astype(np.float32)) * pp["inv_std"].astype(np.float32) if bf16 \
        else (np.asarray(obs, dtype=np.float64) - pp["mean"]) * pp["inv_std"]
    x = q(x)
    h = np.maximum(x @ q(pp["aW1"]) + qb(pp["ab1"]), 0.0)
    h = np.maximum(q(h) @ q(pp["aW2"]) + qb(pp["ab2"]), 0.0)
    head = q(h) @ q(pp["aW3"]) + pp["ab3"]
    c = np.maximum(x @ q(pp["cW1"]) + qb(pp["cb1"]), 0.0)
    c = np.maximum(q(c) @ q(pp["cW2"]) + qb(pp["cb2"]), 0.0)
    value = q(c) @ q(pp["cW3"]) + pp["cb3"][0]
    return head, value


def softplus(x):
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def policy_noise(seed, env_ids, t):
    """The four standard normals the kernel draws for (env, step): Philox stream 1, Box-Muller pairs."""
    raw = philox.draw_blocks(seed, env_ids, np.uint32(t), 1, philox.STREAM_POLICY)
    e0, e1 = philox.normal_pair(raw[:, 0], raw[:, 1])
    e2, e3 = philox.normal_pair(raw[:, 2], raw[:, 3])
    return np.stack([e0, e1, e2, e3], axis=1)


def sample(pp, head, eps, dist, deterministic=False):
    """Returns (raw_action, env_action, log_prob)."""
    z = np.zeros_like(eps) if deterministic else eps
    if dist == 0:
        ls = pp["log_std"]
        raw = head + np.exp(ls) * z
        logp = np.sum(-0.5 * z * z - ls - LOG_SQRT_2PI, axis=1)
        return raw, np.clip(raw, -1.0, 1.0), logp
    loc, scale = head[:, :A], softplus(head[:, A:]) + 0.001
    raw = loc + scale * z
    ldj = 2.0 * (np.log(2.0) - raw - softplus(-2.0 * raw))
    logp = np.sum(-0.5 * z * z - np.log(scale) - LOG_SQRT_2PI - ldj, axis=1)
    return raw, np.tanh(raw), logp


def gae_sb3(reward, value, done, trunc, last_value, gamma, lam):
    """SB3 RolloutBuffer.compute_returns_and_advantage; an episode boundary = done or truncated."""
    T, B = reward.shape
    adv = np.zeros((T, B)); a = np.zeros(B); next_v = np.asarray(last_value, dtype=np.float64)
    for t in range(T - 1, -1, -1):
        fin = np.maximum(done[t], trunc[t]) if trunc is not None else done[t]
        nnt = 1.0 - fin
        delta = reward[t] + gamma * next_v * nnt - value[t]
        a = delta + gamma * lam * nnt * a
        adv[t] = a
        next_v = value[t]
    return adv, adv + value


def gae_brax(reward, value, done, trunc, last_value, gamma, lam):
    """brax.training.agents.ppo.losses.compute_gae with termination = done * (1 - truncation)."""
    T, B = reward.shape
    term = done * (1.0 - trunc)
    mask = 1.0 - trunc
    v_next = np.concatenate([value[1:], np.asarray(last_value, dtype=np.float64)[None]], axis=0)
    deltas = (reward + gamma * (1 - term) * v_next - value) * mask
    acc = np.zeros(B); vs_minus_v = np.zeros((T, B))
    for t in range(T - 1, -1, -1):
        acc = deltas[t] + gamma * (1 - term[t]) * mask[t] * lam * acc
        vs_minus_v[t] = acc
    vs = vs_minus_v + value
    vs_next = np.concatenate([vs[1:], np.asarray(last_value, dtype=np.float64)[None]], axis=0)
    adv = (reward + gamma * (1 - term) * vs_next - value) * mask
    return adv, vs
