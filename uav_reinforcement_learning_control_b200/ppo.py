"""PPO learner around the fused rollout / GAE kernels (data-parallel over GPUs).

Replaces the caller side of the hot path: SB3 ``PPO.learn`` (train.py:50-68,133-137: MlpPolicy
[128,128] ReLU for pi and vf, n_steps 1024, clip 0.19, gamma 0.9906, lambda 0.9079) and Brax
``ppo_train.train`` (train_brax_ppo.py:589-620).  Rollout collection and GAE are the sm_100a kernels
(``qs_rollout_policy``, ``qs_gae``); the clipped-surrogate update itself is ordinary torch autograd on
the GPU (library GEMMs -- it is not on the rollout hot path), with one flat NCCL all-reduce of the
gradient per minibatch (parallel.flat_allreduce_mean_).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from . import config as Q
from .parallel import DistContext, flat_allreduce_mean_, reduce_stats

__all__ = ["PPOConfig", "ActorCritic", "PPOTrainer", "load_sb3_policy_zip", "sb3_state_dict_to_packed"]

H, A = 128, 4


@dataclass
class PPOConfig:
    # train.py:53-60 defaults
    n_steps: int = 128
    gamma: float = 0.99063
    gae_lambda: float = 0.90794
    clip_range: float = 0.1915
    ent_coef: float = 9.1e-5
    vf_coef: float = 0.5
    learning_rate: float = 1.5478e-4
    n_epochs: int = 4
    num_minibatches: int = 8
    max_grad_norm: float = 0.5
    normalize_advantage: bool = True
    timeout_bootstrap: bool = True


class ActorCritic:
    """Separate 2x128 ReLU actor / critic + state-independent log_std, stored as torch Parameters in the
    [in][out] orientation of the packed kernel layout (include/quadsim_abi.h, qs_policy_param_count)."""

    def __init__(self, obs_dim: int, device, seed: int = 0, log_std_init: float = 0.0):
        import torch
        g = torch.Generator(device="cpu"); g.manual_seed(seed)

        def lin(i, o, gain=1.0):
            lim = gain * math.sqrt(3.0 / i)
            w = ((torch.rand(i, o, generator=g) * 2 - 1) * lim).to(device).requires_grad_(True)
            b = torch.zeros(o, device=device, requires_grad=True)
            return [w, b]
        self.obs_dim = obs_dim
        self.actor = lin(obs_dim, H) + lin(H, H) + lin(H, A, 0.01)
        self.critic = lin(obs_dim, H) + lin(H, H) + lin(H, 1, 1.0)
        self.log_std = torch.full((A,), float(log_std_init), device=device, requires_grad=True)
        self.obs_mean = torch.zeros(obs_dim, device=device)
        self.obs_inv_std = torch.ones(obs_dim, device=device)

    def parameters(self):
        return self.actor + self.critic + [self.log_std]

    def pack(self):
        """Flat float32 vector in the kernel's layout (dist 0)."""
        import torch
        with torch.no_grad():
            a, c = self.actor, self.critic
            parts = [a[0].reshape(-1), a[1], a[2].reshape(-1), a[3], a[4].reshape(-1), a[5],
                     c[0].reshape(-1), c[1], c[2].reshape(-1), c[3], c[4].reshape(-1), c[5],
                     self.log_std, self.obs_mean, self.obs_inv_std]
            return torch.cat([p.reshape(-1).float() for p in parts]).contiguous()

    def evaluate(self, obs, raw_action):
        """-> (log_prob, value, entropy) of stored unclipped Gaussian samples (SB3 evaluate_actions)."""
        import torch
        x = (obs - self.obs_mean) * self.obs_inv_std
        a, c = self.actor, self.critic
        h = torch.relu(x @ a[0] + a[1]); h = torch.relu(h @ a[2] + a[3]); mean = h @ a[4] + a[5]
        v = torch.relu(x @ c[0] + c[1]); v = torch.relu(v @ c[2] + c[3]); value = (v @ c[4] + c[5]).squeeze(-1)
        z = (raw_action - mean) * torch.exp(-self.log_std)
        logp = (-0.5 * z * z - self.log_std - 0.9189385332046727).sum(-1)
        ent = (0.5 + 0.9189385332046727 + self.log_std).sum()
        return logp, value, ent


class PPOTrainer:
    """Data-parallel PPO: every rank owns a shard of envs (engine with env_id_offset) and the full policy."""

    def __init__(self, engine, cfg: PPOConfig | None = None, ctx: DistContext | None = None, seed: int = 0):
        import torch
        self.torch = torch
        self.engine = engine
        self.cfg = cfg or PPOConfig()
        self.ctx = ctx or DistContext()
        self.policy = ActorCritic(engine.obs_dim, engine.device, seed=seed)     # same seed on every rank
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=self.cfg.learning_rate, eps=1e-5)
        self.state = engine.new_state()
        engine.reset(self.state)
        self.t = 0
        self.buf = None
        self.adv = self.ret = None

    def collect(self):
        c, eng = self.cfg, self.engine
        boot = c.gamma if c.timeout_bootstrap else 0.0
        self.buf = eng.rollout_policy(self.state, self.policy.pack(), T=c.n_steps, t0=self.t, dist=0,
                                      bootstrap_gamma=boot, buffers=self.buf)
        self.t += c.n_steps
        b = self.buf
        self.adv, self.ret = eng.gae(b["reward"], b["value"], b["done"], b["trunc"], b["last_value"], c.gamma,
                                     c.gae_lambda, brax_form=False, adv=self.adv, ret=self.ret)
        return b

    def update(self):
        torch, c, b = self.torch, self.cfg, self.buf
        T, B = b["reward"].shape
        N = T * B
        obs = b["obs"].reshape(N, -1); act = b["act"].reshape(N, 4)
        old_logp = b["logp"].reshape(N); adv = self.adv.reshape(N); ret = self.ret.reshape(N)
        mb = N // c.num_minibatches
        stats = {"pg_loss": 0.0, "v_loss": 0.0, "n": 0}
        for _ in range(c.n_epochs):
            perm = torch.randperm(N, device=obs.device)
            for k in range(c.num_minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                a_mb = adv[idx]
                if c.normalize_advantage:
                    a_mb = (a_mb - a_mb.mean()) / (a_mb.std() + 1e-8)
                logp, value, ent = self.policy.evaluate(obs[idx], act[idx])
                ratio = torch.exp(logp - old_logp[idx])
                pg = torch.max(-a_mb * ratio, -a_mb * torch.clamp(ratio, 1 - c.clip_range, 1 + c.clip_range)).mean()
                vl = torch.nn.functional.mse_loss(value, ret[idx])
                loss = pg + c.vf_coef * vl - c.ent_coef * ent
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                # the ONLY collective on the training path: one flat all-reduce of ~37.5k floats
                flat_allreduce_mean_([p.grad for p in self.policy.parameters()], self.ctx.world, self.ctx.group)
                torch.nn.utils.clip_grad_norm_(self.policy.parameters(), c.max_grad_norm)
                self.opt.step()
                stats["pg_loss"] += float(pg.detach()); stats["v_loss"] += float(vl.detach()); stats["n"] += 1
        return stats

    def episode_stats(self):
        """Global (all ranks) mean reward per step and episodes finished in the last rollout."""
        b = self.buf
        fin = ((b["done"] != 0) | (b["trunc"] != 0)).sum().item()
        vals = {"reward_sum": b["reward"].sum().item(), "steps": b["reward"].numel(), "episodes": fin}
        out = reduce_stats(vals, self.ctx.world, self.ctx.group, device=self.engine.device)
        out["mean_reward"] = out["reward_sum"] / max(out["steps"], 1.0)
        return out

    def train(self, iterations: int):
        log = []
        for _ in range(iterations):
            self.collect()
            s = self.update()
            s.update(self.episode_stats())
            log.append(s)
        return log


# ------------------------------------------------------------------------------------------------------------------
# SB3 policy import ("SB3 .zip policy import for evaluation parity").
# ``model.save(final_path)`` (train.py:141) writes a zip whose ``policy.pth`` is the torch ``state_dict`` of SB3's
# ``ActorCriticPolicy`` with ``net_arch=dict(pi=[128,128], vf=[128,128])``, ReLU (train.py:61-64).  Its tensors are
# plain ``torch.Tensor``s, so no stable-baselines3 install is needed to read them.  torch Linear stores [out][in]; the
# kernels' packed vector (include/quadsim_abi.h: qs_policy_param_count) stores [in][out].
_SB3_KEYS = {
    "aW1": "mlp_extractor.policy_net.0.weight", "ab1": "mlp_extractor.policy_net.0.bias",
    "aW2": "mlp_extractor.policy_net.2.weight", "ab2": "mlp_extractor.policy_net.2.bias",
    "aW3": "action_net.weight", "ab3": "action_net.bias",
    "cW1": "mlp_extractor.value_net.0.weight", "cb1": "mlp_extractor.value_net.0.bias",
    "cW2": "mlp_extractor.value_net.2.weight", "cb2": "mlp_extractor.value_net.2.bias",
    "cW3": "value_net.weight", "cb3": "value_net.bias",
}


def sb3_state_dict_to_packed(sd, obs_dim: int = 12, obs_mean=None, obs_var=None, eps: float = 1e-8):
    """SB3 ActorCriticPolicy state_dict -> float32 packed parameter vector for ``Engine.rollout_policy(dist=0)``.

    ``obs_mean`` / ``obs_var`` are VecNormalize statistics if the run used them (the reference's train.py does not);
    they become the kernels' (mean, inv_std) observation normaliser."""
    import torch
    missing = [k for k in list(_SB3_KEYS.values()) + ["log_std"] if k not in sd]
    if missing:
        raise KeyError(f"not an SB3 MlpPolicy [128,128]/[128,128] state_dict, missing: {missing}")
    f = lambda k: sd[k].detach().to(torch.float32).cpu()
    shapes = {"aW1": (H, obs_dim), "aW2": (H, H), "aW3": (A, H), "cW1": (H, obs_dim), "cW2": (H, H), "cW3": (1, H)}
    parts = []
    for net in ("a", "c"):
        for layer in ("1", "2", "3"):
            w = f(_SB3_KEYS[f"{net}W{layer}"]); b = f(_SB3_KEYS[f"{net}b{layer}"])
            if tuple(w.shape) != shapes[f"{net}W{layer}"]:
                raise ValueError(f"{_SB3_KEYS[f'{net}W{layer}']}: shape {tuple(w.shape)}, expected {shapes[f'{net}W{layer}']}")
            parts += [w.t().contiguous().reshape(-1), b.reshape(-1)]
    parts.append(f("log_std").reshape(-1))
    mean = torch.zeros(obs_dim) if obs_mean is None else torch.as_tensor(obs_mean, dtype=torch.float32).reshape(-1)
    inv = torch.ones(obs_dim) if obs_var is None else 1.0 / torch.sqrt(torch.as_tensor(obs_var, dtype=torch.float32).reshape(-1) + eps)
    parts += [mean, inv]
    return torch.cat(parts).contiguous()


def load_sb3_policy_zip(path: str, obs_dim: int = 12, device=None):
    """``PPO.save`` zip (train.py:141) -> packed parameter vector on ``device``."""
    import io
    import zipfile
    import torch
    with zipfile.ZipFile(path) as z:
        if "policy.pth" not in z.namelist():
            raise ValueError(f"{path}: no policy.pth inside (not a stable-baselines3 model zip)")
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    p = sb3_state_dict_to_packed(sd, obs_dim)
    return p if device is None else p.to(device)
