"""Per-phase cycle breakdown of the tcgen05 rollout kernel (run on the GPU box).

Build the instrumented variant first (on the dev box, nvcc cross-compiles):
    python tools/tc_profile.py --build
then under gpurun:
    python tools/tc_profile.py
The kernel prints clock64 cycles per step and phase for the first and last thread of CTA 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "build", "tc", "libquadsim_tcprof.so")

if "--build" in sys.argv:
    from uav_reinforcement_learning_control_b200 import build as B
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    print(B.build_library(out=LIB, defines=["QS_TC_PROFILE=1"]))
    sys.exit(0)

os.environ["QS_LIB_PATH"] = LIB
import torch
from bench import make_policy_params
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200.engine import Engine

dev = torch.device("cuda", 0)
for nb, T in ((8192, 256), (1 << 18, 32)):
    eng = Engine(Q.EnvConfig.north_star(seed=1), nb, device=0)
    st = eng.new_state()
    eng.reset(st)
    params = make_policy_params(eng, torch, dev, seed=0)
    buf = eng.rollout_policy(st, params, T=T, t0=0, dist=0, tensor_cores=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.rollout_policy(st, params, T=T, t0=T, dist=0, buffers=buf, tensor_cores=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"== {nb} envs x {T} steps: {ms:.3f} ms, {nb * T / ms / 1e3:.3e} env-steps/s, {ms / T * 1e3:.2f} us/step", flush=True)
    del eng
