"""Experiment: qs_step writing obs / reward / done straight into pinned host memory (zero-copy over PCIe) versus
qs_step_host's staged copies.  Run on the GPU box."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200.engine import Engine

n = 1 << 20
eng = Engine(Q.EnvConfig.north_star(seed=0), n, device=0)
st = eng.new_state(); eng.reset(st)
pin = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory()
h_act = (torch.rand(n, 4) * 2 - 1).pin_memory()
h_obs, h_rew, h_done = pin(n, 12), pin(n), pin(n)
stream = torch.cuda.current_stream()
p = lambda t: C.c_void_p(t.data_ptr())
lib = eng.lib


def zero_copy(act_on_host=True):
    d_act = h_act if act_on_host else h_act.cuda()
    rc = lib.qs_step(eng.handle, p(st), p(d_act), p(h_obs), p(h_rew), p(h_done), None, None, None, None,
                     C.c_void_p(stream.cuda_stream))
    assert rc == 0, lib.qs_last_error_string()
    torch.cuda.synchronize()


for name, fn in (("step_host (staged copies)", lambda: eng.step_host(st, h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())),
                 ("zero-copy outputs + zero-copy actions", lambda: zero_copy(True)),
                 ("zero-copy outputs, actions H2D by torch", lambda: zero_copy(False))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print(f"{name:45s} {dt * 1e3:.3f} ms/step  {n / dt:.3e} env-steps/s  (D2H-equivalent {56 * n / dt / 1e9:.1f} GB/s)", flush=True)
ref = h_obs.clone()
eng.step_host(st, h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())
print("finite", bool(torch.isfinite(h_obs).all()), bool(torch.isfinite(ref).all()))
