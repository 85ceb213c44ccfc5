"""PPO learner around the fused rollout / GAE kernels (data-parallel over GPUs).

Replaces the caller side of the hot path: SB3 ``PPO.learn`` (train.py:50-68,133-137: MlpPolicy
[128,128] ReLU for pi and vf, n_steps 1024, clip 0.19, gamma 0.9906, lambda 0.9079) and Brax
``ppo_train.train`` (train_brax_ppo.py:589-620).  Rollout collection and GAE are the sm_100a kernels
(``qs_rollout_policy``, ``qs_gae``).  The clipped-surrogate update is on device too (``FusedUpdater``:
``qs_ppo_grad`` -- forward + analytic backward of both networks on tcgen05 with the weight gradients accumulating in
TMEM -- and ``qs_ppo_adam``, csrc/qs_ppo.cuh) for the SB3 policy; ``PPOTrainer(fused=False)`` keeps the torch-autograd
update (library GEMMs) that the fused one is tested against.  Either way there is one flat NCCL all-reduce of the
gradient per minibatch and nothing else.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from . import config as Q
from .parallel import DistContext, flat_allreduce_mean_, reduce_stats

__all__ = ["PPOConfig", "ActorCritic", "FusedUpdater", "PPOTrainer", "init_packed_params", "load_sb3_policy_zip", "sb3_state_dict_to_packed",
           "packed_to_sb3_state_dict", "save_sb3_policy_weights", "save_sb3_policy_zip"]

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
    adam_eps: float = 1e-5                 # SB3's PPO default; optax.adam (brax): 1e-8

    @staticmethod
    def sb3_reference(**kw):
        """The SB3 learner exactly as train.py:50-68 configures it (n_steps 1024, 20 epochs, batch_size 128 on 16 envs =
        128 minibatches); the class defaults above keep the reference's loss hyper-parameters but use a rollout / minibatch
        geometry sized for 10^4..10^6 envs per GPU."""
        from dataclasses import replace
        return replace(PPOConfig(n_steps=1024, n_epochs=20, num_minibatches=128), **kw)

    @staticmethod
    def brax_reference(**kw):
        """brax ppo_train.train as train_brax_ppo.py:431-455,589-620 calls it: unroll_length 10, 16 minibatches x 4 updates
        per batch, lr 3e-4, entropy_cost 1e-3, discounting 0.99, gae_lambda 0.95, brax's default clipping_epsilon 0.3,
        value loss 0.5 * 0.5 * err^2, advantages normalised with the population std, optax.adam (eps 1e-8), no gradient
        clipping, running observation normaliser."""
        from dataclasses import replace
        return replace(PPOConfig(n_steps=10, gamma=0.99, gae_lambda=0.95, clip_range=0.3, ent_coef=1e-3, vf_coef=0.25,
                                 learning_rate=3e-4, n_epochs=4, num_minibatches=16, max_grad_norm=0.0,
                                 timeout_bootstrap=False, adam_eps=1e-8), **kw)


def init_packed_params(obs_dim: int, dist: int, seed: int = 0, log_std_init: float = 0.0):
    """Random-init 2x128 ReLU actor-critic in the packed layout of qs_policy_param_count (lecun-uniform weights, zero
    biases, small actor head; identity observation normaliser) -- float32 CPU tensor."""
    import torch
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    Ao = 2 * A if dist == 1 else A
    parts = []
    for out, gain in ((Ao, 0.01), (1, 1.0)):
        for (i, o, gn) in ((obs_dim, H, 1.0), (H, H, 1.0), (H, out, gain)):
            lim = gn * math.sqrt(3.0 / i)
            parts += [((torch.rand(i, o, generator=g) * 2 - 1) * lim).reshape(-1), torch.zeros(o)]
    if dist == 0:
        parts.append(torch.full((A,), float(log_std_init)))
    parts += [torch.zeros(obs_dim), torch.ones(obs_dim)]
    return torch.cat(parts).contiguous()


class ActorCritic:
    """Separate 2x128 ReLU actor / critic + state-independent log_std, stored as torch Parameters in the
    [in][out] orientation of the packed kernel layout (include/quadsim_abi.h, qs_policy_param_count)."""

    def __init__(self, obs_dim: int, device, seed: int = 0, log_std_init: float = 0.0, init: str = "orthogonal"):
        """init = "orthogonal": SB3's ActorCriticPolicy default (ortho_init=True: gain sqrt(2) on the hidden layers, 0.01 on
        the action head, 1 on the value head, zero biases); "uniform": lecun-uniform."""
        import torch
        g = torch.Generator(device="cpu"); g.manual_seed(seed)

        def lin(i, o, gain_ortho, gain_uniform=1.0):
            if init == "orthogonal":
                w = torch.empty(o, i)                       # torch.nn.Linear orientation [out][in] ...
                torch.nn.init.orthogonal_(w, gain=gain_ortho, generator=g)
                w = w.t().contiguous()                      # ... stored [in][out] like the packed kernel layout
            else:
                w = (torch.rand(i, o, generator=g) * 2 - 1) * (gain_uniform * math.sqrt(3.0 / i))
            return [w.to(device).requires_grad_(True), torch.zeros(o, device=device, requires_grad=True)]
        self.obs_dim = obs_dim
        self.actor = lin(obs_dim, H, math.sqrt(2)) + lin(H, H, math.sqrt(2)) + lin(H, A, 0.01, 0.01)
        self.critic = lin(obs_dim, H, math.sqrt(2)) + lin(H, H, math.sqrt(2)) + lin(H, 1, 1.0)
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

    def load_packed(self, vec):
        """Inverse of ``pack`` (the obs normaliser included)."""
        import torch
        o = 0
        with torch.no_grad():
            for group in (self.actor, self.critic, [self.log_std], [self.obs_mean], [self.obs_inv_std]):
                for t in group:
                    t.copy_(vec[o:o + t.numel()].reshape(t.shape).to(t.dtype))
                    o += t.numel()
        assert o == vec.numel(), (o, vec.numel())

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


class FusedUpdater:
    """The PPO minibatch update as sm_100a kernels (include/quadsim_abi.h: qs_ppo_grad / qs_ppo_adam) on the PACKED
    parameter vector -- the same tensor the rollout kernels read, so there is no pack / unpack between rollout and
    update.  Holds the Adam moments, the per-CTA partial-gradient workspace and the flat gradient (+ 8 statistics)."""

    N_STATS = 8

    def __init__(self, device, obs_dim: int = 12, dist: int = 0):
        import ctypes as C
        import torch
        from .engine import QuadSimError, load_library
        if not torch.cuda.is_available():
            raise QuadSimError("no CUDA device: the fused PPO update has no CPU fallback")
        self.torch, self.C = torch, C
        self.lib = load_library()
        self.device = torch.device(device)
        d = Q.QsPolicyDesc()
        d.obs_dim, d.hidden, d.act_dim, d.dist = int(obs_dim), H, A, int(dist)
        self.desc = d
        self.obs_dim, self.dist = int(obs_dim), int(dist)
        with torch.cuda.device(self.device):
            nbytes = int(self.lib.qs_ppo_workspace_bytes(C.byref(d)))
        if nbytes <= 0:
            raise QuadSimError(f"qs_ppo_workspace_bytes: {self.lib.qs_last_error_string().decode()}")
        self.P = int(self.lib.qs_policy_param_count(C.byref(d)))
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.grad_buf = torch.zeros(self.P + self.N_STATS, dtype=torch.float32, device=self.device)
        self.m = torch.zeros(self.P, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(self.P, dtype=torch.float32, device=self.device)
        self.norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.step = 0
        self.comm = None          # multi-GPU: peer-memory exchange (enable_peer)
        self.epoch = 0
        # running observation normaliser (brax normalize_observations=True): count | mean[D] | M2[D], float64 on the device
        self.obs_running = torch.zeros(1 + 2 * self.obs_dim, dtype=torch.float64, device=self.device)
        self._obs_ws = None

    def _check(self, rc, what):
        if rc != 0:
            from .engine import QuadSimError
            raise QuadSimError(f"{what}: libquadsim error {rc}: {self.lib.qs_last_error_string().decode()}")

    def enable_peer(self, world: int, rank: int, group=None) -> bool:
        """Multi-GPU, one process per GPU: exchange the gradients through CUDA-IPC-mapped peer buffers and one fused
        wait + sum + clip + Adam kernel per rank (qs_ppo_adam_peer) instead of reduce -> NCCL all-reduce -> Adam.  The
        64-byte IPC handles travel through one torch.distributed all_gather at set-up; nothing else uses NCCL.

        Set-up is split into phases that contain only LOCAL work, each followed by a MIN all-reduce of the status, so every
        rank issues the same sequence of collectives whatever fails where.  Returns True when every rank mapped every peer;
        otherwise every rank has torn its part down (collectively: unmap, barrier, free) and the caller stays on NCCL."""
        import torch.distributed as dist
        torch, C = self.torch, self.C

        def unanimous(ok):
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1

        def report(e):
            print(f"[rank {rank}] peer-memory gradient exchange unavailable ({e}); using NCCL", flush=True)

        self.world, self.rank, self._group = int(world), int(rank), group
        # phase 1 (local): create this rank's buffer and export its handle
        buf, ok = (C.c_ubyte * 64)(), True
        try:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                self._check(self.lib.qs_ppo_comm_create(C.byref(self.desc), int(world), int(rank), C.byref(h)), "qs_ppo_comm_create")
                self.comm = h                      # from here on close() frees it
                self._check(self.lib.qs_ppo_comm_export(h, buf), "qs_ppo_comm_export")
        except Exception as e:                     # noqa: BLE001 -- reported, then decided collectively
            ok = False
            report(e)
        if not unanimous(ok):
            self.close(collective=False)           # nobody has mapped anything yet
            return False
        # phase 2 (collective, unconditional): exchange the handles
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=self.device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        # phase 3 (local): map the peers
        try:
            with torch.cuda.device(self.device):
                for p in range(world):
                    if p != rank:
                        raw = (C.c_ubyte * 64).from_buffer_copy(bytes(handles[p].cpu().tolist()))
                        self._check(self.lib.qs_ppo_comm_import(self.comm, p, raw), "qs_ppo_comm_import")
        except Exception as e:                     # noqa: BLE001
            ok = False
            report(e)
        if not unanimous(ok):
            self.close(collective=True)            # some ranks map this rank's buffer: unmap, barrier, then free
            return False
        dist.barrier(group=group)
        return True

    def close(self, collective: bool = True):
        """Tear the peer exchange down.  collective=True (every rank calls it): unmap the peers, barrier, then free -- an
        exported buffer must not be freed while a peer still maps it.  Safe to call on a rank whose set-up failed."""
        if collective and getattr(self, "_group", "unset") != "unset" and getattr(self, "world", 1) > 1:
            import torch.distributed as dist
            self.torch.cuda.synchronize(self.device)
            dist.barrier(group=self._group)                  # nobody is still reading a slot
            if self.comm is not None:
                self.lib.qs_ppo_comm_close_peers(self.comm)
            dist.barrier(group=self._group)                  # nobody still maps this rank's buffer
        if self.comm is not None:
            self.lib.qs_ppo_comm_destroy(self.comm)
        self.comm = None

    def _stream(self):
        return self.C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def update_obs_stats(self, params, obs, std_min=1e-6, std_max=1e6, world: int = 1, group=None):
        """Merge the batch ``obs`` [N, D] into the running statistics and refresh obs_mean / obs_inv_std INSIDE the packed
        parameter vector ``params`` (qs_obs_stats_update; brax running_statistics.update, train_brax_ppo.py:611).
        world > 1 (data-parallel ranks): every rank reduces its own shard to (n, mean, M2) with the kernel, the 1 + 2 D
        doubles are all-gathered and merged in rank order on every rank, so the normaliser -- part of the policy -- stays
        bitwise identical across ranks (brax pmean's its statistics for the same reason)."""
        torch, C = self.torch, self.C
        D = self.obs_dim
        obs = obs.reshape(-1, D)
        if obs.dtype != torch.float32 or not obs.is_cuda or not obs.is_contiguous():
            raise ValueError("obs: expected a contiguous float32 CUDA tensor [N, obs_dim]")
        if self._obs_ws is None:
            with torch.cuda.device(self.device):
                self._obs_ws = torch.zeros(int(self.lib.qs_obs_stats_workspace_bytes(D)), dtype=torch.uint8, device=self.device)
        P = params.numel()

        def run(running, mean_ptr, inv_ptr):
            with torch.cuda.device(self.device):
                self._check(self.lib.qs_obs_stats_update(C.c_void_p(obs.data_ptr()), int(obs.shape[0]), D,
                                                         C.c_void_p(running.data_ptr()), C.c_void_p(mean_ptr), C.c_void_p(inv_ptr),
                                                         float(std_min), float(std_max), C.c_void_p(self._obs_ws.data_ptr()),
                                                         self._stream()), "qs_obs_stats_update")
        if world <= 1:
            run(self.obs_running, params.data_ptr() + 4 * (P - 2 * D), params.data_ptr() + 4 * (P - D))
            return
        import torch.distributed as dist
        local = torch.zeros_like(self.obs_running)                 # merge into an empty state = this shard's (n, mean, M2)
        scratch = torch.empty(2 * D, dtype=torch.float32, device=self.device)
        run(local, scratch.data_ptr(), scratch.data_ptr() + 4 * D)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        R = self.obs_running
        for b in parts:                                            # Chan merge, rank order, float64, identical on every rank
            na, nb = R[0].clone(), b[0]
            tot = na + nb
            delta = b[1:1 + D] - R[1:1 + D]
            R[1 + D:] = R[1 + D:] + b[1 + D:] + delta * delta * na * nb / tot
            R[1:1 + D] = R[1:1 + D] + delta * nb / tot
            R[0] = tot
        sd = torch.sqrt(torch.clamp(R[1 + D:] / R[0], min=0.0)).clamp(std_min, std_max)
        params[P - 2 * D:P - D] = R[1:1 + D].float()
        params[P - D:] = (1.0 / sd).float()

    ROW_FLOATS = 32      # one 128-byte line per sample (include/quadsim_abi.h: qs_ppo_pack)

    def pack(self, obs, act, old_logp, adv, ret, out=None):
        """Rollout / GAE arrays -> packed sample rows [N, 32] (obs | act | old_logp | adv | ret | 0..), once per rollout.
        The minibatch gather of ``grad(packed=...)`` then touches one full cache line per sample (qs_ppo_pack)."""
        torch = self.torch
        N = old_logp.numel()
        for name, t, shp in (("obs", obs, (N, self.obs_dim)), ("act", act, (N, 4)), ("old_logp", old_logp, (N,)),
                             ("adv", adv, (N,)), ("ret", ret, (N,))):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shp:
                raise ValueError(f"{name}: expected contiguous float32 CUDA tensor {shp}, got {t.dtype} {tuple(t.shape)}")
        if out is None or tuple(out.shape) != (N, self.ROW_FLOATS):
            out = torch.empty((N, self.ROW_FLOATS), dtype=torch.float32, device=self.device)
        p = lambda t: self.C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            self._check(self.lib.qs_ppo_pack(self.C.byref(self.desc), p(obs), p(act), p(old_logp), p(adv), p(ret), int(N), p(out),
                                             self._stream()), "qs_ppo_pack")
        return out

    def grad(self, params, obs=None, act=None, old_logp=None, adv=None, ret=None, idx=None, clip_range=0.2, vf_coef=0.5,
             ent_coef=0.0, normalize_adv=True, sample_seed=0, packed=None):
        """Gradient of the PPO loss (SB3 form for dist 0, brax form for dist 1) over rows ``idx`` (int32, or None: all rows) of
        the flattened rollout buffers -- either the five arrays, or ``packed`` rows from ``pack`` (+ ``adv`` for the minibatch
        statistics).  normalize_adv: False / True (unbiased std) / 2 (population std, brax).
        Returns the [P + 8] buffer: packed gradient | loss statistics (sums, see quadsim_abi.h)."""
        torch = self.torch
        if packed is not None:
            N = packed.shape[0]
            checks = (("params", params, (self.P,)), ("packed", packed, (N, self.ROW_FLOATS))) + \
                     ((("adv", adv, (N,)),) if normalize_adv else ())
        else:
            N = old_logp.numel()
            checks = (("params", params, (self.P,)), ("obs", obs, (N, self.desc.obs_dim)), ("act", act, (N, 4)),
                      ("old_logp", old_logp, (N,)), ("adv", adv, (N,)), ("ret", ret, (N,)))
        for name, t, shp in checks:
            if t is None or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shp:
                raise ValueError(f"{name}: expected contiguous float32 CUDA tensor {shp}, got "
                                 f"{None if t is None else (t.dtype, tuple(t.shape))}")
        if idx is not None and (idx.dtype != torch.int32 or not idx.is_cuda or not idx.is_contiguous() or idx.dim() != 1):
            raise ValueError("idx: expected a contiguous 1-D int32 CUDA tensor")
        n = N if idx is None else idx.numel()
        p = lambda t: None if t is None else self.C.c_void_p(t.data_ptr())
        # peer mode: the gradient goes straight into this rank's exported slot of the NEXT update
        out = p(self.grad_buf) if self.comm is None else self.C.c_void_p(self.lib.qs_ppo_comm_slot(self.comm, self.epoch + 1))
        self.desc.sample_seed = int(sample_seed) & 0x7FFFFFFF
        with torch.cuda.device(self.device):
            if packed is not None:
                self._check(self.lib.qs_ppo_grad_packed(self.C.byref(self.desc), p(params), p(packed), p(adv), p(idx), int(n),
                                                        float(clip_range), float(vf_coef), float(ent_coef), int(normalize_adv),
                                                        p(self.workspace), out, self._stream()), "qs_ppo_grad_packed")
            else:
                self._check(self.lib.qs_ppo_grad(self.C.byref(self.desc), p(params), p(obs), p(act), p(old_logp), p(adv), p(ret),
                                                 p(idx), int(n), float(clip_range), float(vf_coef), float(ent_coef),
                                                 int(normalize_adv), p(self.workspace), out, self._stream()),
                            "qs_ppo_grad")
        return self.grad_buf if self.comm is None else None

    def update_epoch(self, params, packed, adv, perm, num_minibatches, lr, clip_range=0.2, vf_coef=0.5, ent_coef=0.0,
                     normalize_adv=True, max_grad_norm=0.5, beta1=0.9, beta2=0.999, eps=1e-5, stats_acc=None, sample_seed0=0):
        """One epoch of minibatch updates over ``perm`` (int32 permutation of the N packed rows), launched back to back from
        native code (qs_ppo_update_epoch): per minibatch {advantage statistics, gradient, optimiser step} -- through the
        peer-memory exchange when ``enable_peer`` succeeded -- with the loss statistics accumulated into ``stats_acc``
        (8 floats) inside the optimiser kernel.  Bitwise identical to calling ``grad`` + ``adam`` / ``adam_peer`` per
        minibatch; it exists because the reference's own geometry (train_brax_ppo.py: 10 240-sample minibatches) spends
        more time in the Python / ctypes round trips of those calls than on the GPU."""
        torch = self.torch
        N = packed.shape[0]
        for name, t, shp, dt in (("params", params, (self.P,), torch.float32), ("packed", packed, (N, self.ROW_FLOATS), torch.float32),
                                 ("perm", perm, (N,), torch.int32)) + \
                                ((("adv", adv, (N,), torch.float32),) if normalize_adv else ()) + \
                                ((("stats_acc", stats_acc, (self.N_STATS,), torch.float32),) if stats_acc is not None else ()):
            if t is None or t.dtype != dt or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shp:
                raise ValueError(f"{name}: expected contiguous {dt} CUDA tensor {shp}, got "
                                 f"{None if t is None else (t.dtype, tuple(t.shape))}")
        p = lambda t: None if t is None else self.C.c_void_p(t.data_ptr())
        k = int(num_minibatches)
        with torch.cuda.device(self.device):
            self._check(self.lib.qs_ppo_update_epoch(
                self.C.byref(self.desc), p(params), p(packed), p(adv), p(perm), int(N), k, float(clip_range), float(vf_coef),
                float(ent_coef), int(normalize_adv), p(self.m), p(self.v), int(self.step + 1), float(lr), float(beta1), float(beta2),
                float(eps), float(max_grad_norm), self.comm, int(self.epoch + 1) if self.comm is not None else 0, p(self.workspace),
                p(self.grad_buf), p(stats_acc), p(self.norm), int(sample_seed0) & 0xFFFFFFFFFFFFFFFF, self._stream()),
                "qs_ppo_update_epoch")
        self.step += k
        if self.comm is not None:
            self.epoch += k
        return self.norm

    def adam_peer(self, params, lr, max_grad_norm=0.5, beta1=0.9, beta2=0.999, eps=1e-5, stats_acc=None):
        """Peer mode: wait for every rank's slot of this update, sum them over NVLink, clip, Adam -- one kernel."""
        self.step += 1
        self.epoch += 1
        p = lambda t: None if t is None else self.C.c_void_p(t.data_ptr())
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qs_ppo_adam_peer(self.C.byref(self.desc), self.comm, int(self.epoch), p(params), p(self.m),
                                                  p(self.v), int(self.step), float(lr), float(beta1), float(beta2), float(eps),
                                                  float(max_grad_norm), p(self.norm), p(stats_acc), self._stream()),
                        "qs_ppo_adam_peer")
        return self.norm

    def permutation(self, n: int, seed: int, epoch: int, out=None):
        """int32 [n] pseudo-random permutation of 0..n-1 keyed by (seed, epoch) (qs_ppo_permutation): the per-epoch shuffle
        of SB3's RolloutBuffer.get computed element-wise on the device instead of a sort-based randperm."""
        torch = self.torch
        if out is None or out.numel() != n:
            out = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.qs_ppo_permutation(int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF,
                                                    self.C.c_void_p(out.data_ptr()), self._stream()), "qs_ppo_permutation")
        return out

    def adam(self, params, lr, max_grad_norm=0.5, grad_scale=1.0, beta1=0.9, beta2=0.999, eps=1e-5, grad=None):
        """clip_grad_norm_ + Adam step on ``params`` in place, from ``self.grad_buf`` (after any all-reduce)."""
        self.step += 1
        g = self.grad_buf if grad is None else grad
        p = lambda t: self.C.c_void_p(t.data_ptr())
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qs_ppo_adam(self.C.byref(self.desc), p(params), p(g), p(self.m), p(self.v), int(self.step),
                                             float(lr), float(beta1), float(beta2), float(eps), float(max_grad_norm),
                                             float(grad_scale), p(self.norm), self._stream()), "qs_ppo_adam")
        return self.norm


class PPOTrainer:
    """Data-parallel PPO: every rank owns a shard of envs (engine with env_id_offset) and the full policy.

    fused=True (default for the 12-D SB3 policy): the update runs as qs_ppo_grad / qs_ppo_adam on the packed parameter
    vector ``self.params``; fused=False: torch autograd on ``self.policy``."""

    def __init__(self, engine, cfg: PPOConfig | None = None, ctx: DistContext | None = None, seed: int = 0,
                 fused: bool | None = None, tensor_cores: bool | None = None, peer: bool = False, packed_rows: bool = True,
                 native_epochs: bool = True):
        import torch
        self.torch = torch
        self.packed_rows = bool(packed_rows)
        self.native_epochs = bool(native_epochs)       # False: one grad + adam call per minibatch from Python (A/B, tests)
        self.engine = engine
        self.cfg = cfg or PPOConfig()
        self.ctx = ctx or DistContext()
        # Brax env variants (21-D raw observation) train brax's tanh-normal policy with brax's loss; gym variants the SB3 one
        self.dist = 1 if engine.cfg.mode in (Q.MODE_MJX_BRAX, Q.MODE_HOVER_BRAX) else 0
        self.brax = self.dist == 1
        self.fused = (True if fused is None else bool(fused))
        if self.brax and not self.fused:
            raise ValueError("the Brax-policy update exists as fused kernels only (fused=True)")
        self.policy = None if self.brax else ActorCritic(engine.obs_dim, engine.device, seed=seed)     # same seed on every rank
        self.tensor_cores = self.fused if tensor_cores is None else bool(tensor_cores)
        if self.brax:
            self.params = init_packed_params(engine.obs_dim, 1, seed).to(engine.device)
        else:
            self.params = self.policy.pack() if self.fused else None             # fused: THE master copy of the weights
        self.updater = FusedUpdater(engine.device, engine.obs_dim, self.dist) if self.fused else None
        if peer and self.fused and self.ctx.world > 1:
            # Set-up (CUDA IPC) is the only part that can fail for environmental reasons; enable_peer decides
            # unanimously (every rank uses the peer path or every rank stays on the NCCL all-reduce).
            self.updater.enable_peer(self.ctx.world, self.ctx.rank, self.ctx.group)
        self.shuffle_seed = (int(seed) << 20) ^ (0x5EED + 7919 * self.ctx.rank)     # every rank shuffles its own rows
        self._epochs_done = 0
        self.opt = None if self.brax else torch.optim.Adam(self.policy.parameters(), lr=self.cfg.learning_rate, eps=self.cfg.adam_eps)
        self.state = engine.new_state()
        # Brax AutoResetWrapper restores the episode-0 state: the engine needs that state for every env
        self.first_state = (torch.zeros(21, engine.num_envs, device=engine.device)
                            if engine.cfg.auto_reset == Q.RESET_RESTORE_FIRST else None)
        engine.reset(self.state, first_state=self.first_state)
        self._updates = 0
        self.t = 0
        self.buf = None
        self.adv = self.ret = None

    def _prefetch_permutations(self):
        """The n_epochs shuffles of the coming update depend on (seed, epoch counter) only: compute them on a side stream
        while the rollout runs (pure integer ALU work next to a latency- / memory-bound kernel) instead of on the update's
        critical path (0.11 ms per epoch at 2^23 rows: the Feistel cycle walk diverges)."""
        torch, c, up = self.torch, self.cfg, self.updater
        N = c.n_steps * self.engine.num_envs
        if getattr(self, "_perms", None) is None or self._perms[0].numel() != N or len(self._perms) != c.n_epochs:
            self._perms = [torch.empty(N, dtype=torch.int32, device=self.engine.device) for _ in range(c.n_epochs)]
            self._side = torch.cuda.Stream(device=self.engine.device)
        main = torch.cuda.current_stream(self.engine.device)
        self._side.wait_stream(main)                      # the previous update has finished reading the buffers
        with torch.cuda.stream(self._side):
            for e in range(c.n_epochs):
                up.permutation(N, self.shuffle_seed, self._epochs_done + e, out=self._perms[e])
            self._perms_ready = torch.cuda.Event()
            self._perms_ready.record(self._side)
        self._perms_epoch0 = self._epochs_done

    def collect(self):
        c, eng = self.cfg, self.engine
        boot = c.gamma if (c.timeout_bootstrap and not self.brax) else 0.0
        if self.fused:
            self._prefetch_permutations()
        self.buf = eng.rollout_policy(self.state, self.params if self.fused else self.policy.pack(), T=c.n_steps,
                                      t0=self.t, dist=self.dist, bootstrap_gamma=boot, tensor_cores=self.tensor_cores,
                                      buffers=self.buf, first_state=self.first_state)
        self.t += c.n_steps
        b = self.buf
        self.adv, self.ret = eng.gae(b["reward"], b["value"], b["done"], b["trunc"], b["last_value"], c.gamma,
                                     c.gae_lambda, brax_form=self.brax, adv=self.adv, ret=self.ret)
        if self.brax:
            # brax merges the batch's observations into the running normaliser before the SGD steps; the refreshed
            # obs_mean / obs_inv_std live inside self.params, i.e. the next rollout and this update both see them
            self.updater.update_obs_stats(self.params, b["obs"], world=self.ctx.world, group=self.ctx.group)
        return b

    def update(self):
        return self.update_fused() if self.fused else self.update_autograd()

    def update_fused(self):
        """n_epochs x num_minibatches of {qs_ppo_grad -> one flat all-reduce -> qs_ppo_adam}; no host sync inside."""
        torch, c, b, up = self.torch, self.cfg, self.buf, self.updater
        T, B = b["reward"].shape
        N = T * B
        obs = b["obs"].reshape(N, -1); act = b["act"].reshape(N, 4)
        old_logp = b["logp"].reshape(N); adv = self.adv.reshape(N); ret = self.ret.reshape(N)
        mb = N // c.num_minibatches
        acc = torch.zeros(up.N_STATS, dtype=torch.float32, device=obs.device)
        world = self.ctx.world
        # one 128-byte row per sample, written once per rollout: every minibatch gather then reads full cache lines
        self._packed = packed = up.pack(obs, act, old_logp, adv, ret, out=getattr(self, "_packed", None)) if self.packed_rows else None
        norm_mode = (2 if self.brax else 1) if c.normalize_advantage else 0
        # single GPU, or gradients exchanged through peer memory: the whole epoch is one native call; only the NCCL
        # fallback needs the host between the gradient and the optimiser step
        native = self.native_epochs and packed is not None and (world == 1 or up.comm is not None)
        prefetched = getattr(self, "_perms", None) is not None and getattr(self, "_perms_epoch0", -1) == self._epochs_done \
            and self._perms[0].numel() == N and len(self._perms) == c.n_epochs
        if prefetched:
            torch.cuda.current_stream(obs.device).wait_event(self._perms_ready)
        for e in range(c.n_epochs):
            if prefetched:
                perm = self._perms[e]                      # computed next to the rollout (collect)
            else:
                self._perm = perm = up.permutation(N, self.shuffle_seed, self._epochs_done, out=getattr(self, "_perm", None))
            self._epochs_done += 1
            if native:
                up.update_epoch(self.params, packed, adv, perm, c.num_minibatches, c.learning_rate, clip_range=c.clip_range,
                                vf_coef=c.vf_coef, ent_coef=c.ent_coef, normalize_adv=norm_mode, max_grad_norm=c.max_grad_norm,
                                eps=c.adam_eps, stats_acc=acc, sample_seed0=self.shuffle_seed * 2654435761 + self._updates + 1)
                self._updates += c.num_minibatches
                continue
            for k in range(c.num_minibatches):
                self._updates += 1
                # (the last minibatch takes the N % num_minibatches remainder rows, as SB3's RolloutBuffer.get does)
                g = up.grad(self.params, obs, act, old_logp, adv, ret, packed=packed,
                            idx=perm[k * mb:((k + 1) * mb if k + 1 < c.num_minibatches else N)],
                            clip_range=c.clip_range, vf_coef=c.vf_coef, ent_coef=c.ent_coef,
                            normalize_adv=norm_mode,
                            sample_seed=(self.shuffle_seed * 2654435761 + self._updates) & 0x7FFFFFFF)
                if up.comm is not None:
                    # gradients meet in NVLink peer memory inside the optimiser kernel: no collective call at all
                    up.adam_peer(self.params, c.learning_rate, max_grad_norm=c.max_grad_norm, eps=c.adam_eps, stats_acc=acc)
                    continue
                if world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(g, group=self.ctx.group)       # the ONLY collective: P + 8 floats, sum
                acc += g[up.P:]
                up.adam(self.params, c.learning_rate, max_grad_norm=c.max_grad_norm, grad_scale=1.0 / world, eps=c.adam_eps)
        s = acc.tolist()                                             # one D2H read per update
        if up.comm is not None:                                      # (the stream is idle now) did every peer arrive in time?
            up._check(up.lib.qs_ppo_comm_error(up.comm), "qs_ppo_adam_peer")
        nb = c.n_epochs * c.num_minibatches
        return {"pg_loss": s[0] / max(s[4], 1.0) * nb, "v_loss": s[1] / max(s[4], 1.0) * nb,
                "clip_frac": s[2] / max(s[4], 1.0), "approx_kl": s[3] / max(s[4], 1.0), "n": nb}

    def update_autograd(self):
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
                idx = perm[k * mb:((k + 1) * mb if k + 1 < c.num_minibatches else N)]
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

    def packed_params(self):
        return self.params if self.fused else self.policy.pack()

    def set_log_std(self, value: float):
        if self.brax:
            raise ValueError("the tanh-normal policy has no state-independent log_std")
        self.policy.log_std.data.fill_(float(value))
        if self.fused:
            D = self.engine.obs_dim
            self.params[self.params.numel() - 2 * D - A:self.params.numel() - 2 * D] = float(value)

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


def packed_to_sb3_state_dict(packed, obs_dim: int = 12):
    """Inverse of ``sb3_state_dict_to_packed``: the kernels' packed vector (e.g. ``PPOTrainer.packed_params()`` after
    training) -> the ``state_dict`` of SB3's ``ActorCriticPolicy`` with ``net_arch=dict(pi=[128,128], vf=[128,128])``
    (train.py:61-64), so that a policy trained here can go back through ``model.policy.load_state_dict`` into the
    reference's evaluate.py.  The observation normaliser must be the identity (the reference uses no VecNormalize)."""
    import torch
    v = packed.detach().to(torch.float32).cpu()
    sizes = [("aW1", (obs_dim, H)), ("ab1", (H,)), ("aW2", (H, H)), ("ab2", (H,)), ("aW3", (H, A)), ("ab3", (A,)),
             ("cW1", (obs_dim, H)), ("cb1", (H,)), ("cW2", (H, H)), ("cb2", (H,)), ("cW3", (H, 1)), ("cb3", (1,)),
             ("log_std", (A,)), ("mean", (obs_dim,)), ("inv_std", (obs_dim,))]
    t, o = {}, 0
    for name, shp in sizes:
        n = 1
        for d in shp:
            n *= d
        t[name] = v[o:o + n].reshape(shp)
        o += n
    if o != v.numel():
        raise ValueError(f"packed vector has {v.numel()} floats, expected {o} for obs_dim {obs_dim}")
    if bool((t["mean"] != 0).any()) or bool((t["inv_std"] != 1).any()):
        raise ValueError("the packed policy carries a non-identity observation normaliser, which an SB3 MlpPolicy cannot hold")
    sd = {"log_std": t["log_std"].clone()}
    for net in ("a", "c"):
        for layer in ("1", "2", "3"):
            sd[_SB3_KEYS[f"{net}W{layer}"]] = t[f"{net}W{layer}"].t().contiguous()      # torch Linear stores [out][in]
            sd[_SB3_KEYS[f"{net}b{layer}"]] = t[f"{net}b{layer}"].clone()
    return sd


def save_sb3_policy_weights(path: str, packed, obs_dim: int = 12):
    """Write ``policy.pth`` (the state_dict above) into a zip.  WEIGHTS ONLY: the archive has the member SB3's ``PPO.save``
    calls ``policy.pth`` but none of its pickled ``data`` (observation / action spaces, hyper-parameters), so ``PPO.load``
    cannot open it (that would need gymnasium + stable-baselines3, absent here).  On the reference side build the model
    as train.py:50-68 does and load the weights into it:

        sd = torch.load(io.BytesIO(zipfile.ZipFile(path).read("policy.pth")), weights_only=True)
        model.policy.load_state_dict(sd)                      # evaluate.py:309,459,628 then use `model` as usual

    ``load_sb3_policy_zip`` reads both this file and a real ``PPO.save`` archive."""
    import io
    import zipfile
    import torch
    buf = io.BytesIO()
    torch.save(packed_to_sb3_state_dict(packed, obs_dim), buf)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("README.txt", "weights only (state_dict of SB3 ActorCriticPolicy pi=[128,128] vf=[128,128]); load with "
                                 "model.policy.load_state_dict, not PPO.load")


save_sb3_policy_zip = save_sb3_policy_weights      # former name
