"""Same-box A/B of the step kernel across library builds (build/variants/libquadsim_*.so): only qs_create / qs_reset /
qs_step are bound, so builds with a different export list can be compared.  CUDA events, 2^20 hover envs, 300 steps."""
import ctypes as C
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import model as M

n = 1 << 20
cfg = Q.EnvConfig.north_star(seed=0)
params = Q.pack_params(M.derive_constants(M.load_mjcf(M.default_model_path())), cfg)
dev = torch.device("cuda", 0)
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libquadsim_*.so")))
for rnd in range(3):
    for path in libs:
        lib = C.CDLL(path)
        lib.qs_create.argtypes = [C.POINTER(Q.QsParams), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.qs_reset.argtypes = [C.c_void_p] * 6
        lib.qs_step.argtypes = [C.c_void_p] * 11
        lib.qs_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        assert lib.qs_create(C.byref(params), n, 0, None, None, C.byref(h)) == 0
        st = torch.zeros(Q.NPLANES, n, device=dev); st[3] = 1.0
        obs = torch.empty(n, 12, device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, device=dev)
        act = torch.rand(n, 4, device=dev) * 2 - 1
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        assert lib.qs_reset(h, p(st), None, p(obs), None, s) == 0
        step = lambda: lib.qs_step(h, p(st), p(act), p(obs), p(rew), p(done), None, None, None, None, s)
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300):
            step()
        e1.record(); torch.cuda.synchronize()
        print(f"round {rnd} {os.path.basename(path):32s} {e0.elapsed_time(e1) / 300 * 1e3:7.2f} us/step", flush=True)
        lib.qs_destroy(h)
