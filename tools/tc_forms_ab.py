"""A/B of the tcgen05 rollout's CTA forms on one box (run under gpurun): QS_TC_FORM = 2 (two plain tiles per CTA) vs 3 (two
compact tiles, each with a partner warpgroup) at large batches, 1 (one tile + partners) vs default at 8192 envs.
CUDA events around `reps` back-to-back launches after a warm-up launch; prints env-steps/s per form."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_policy_params
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200.engine import Engine

dev = torch.device("cuda", 0)
cases = [(8192, 256, ("1", "3", "2")), (1 << 18, 32, ("2", "3")), (1 << 20, 32, ("2", "3"))]
for nb, T, forms in cases:
    eng = Engine(Q.EnvConfig.north_star(seed=1), nb, device=0)
    params = make_policy_params(eng, torch, dev, seed=0)
    if os.environ.get("QS_AB_CALM_POLICY"):      # zero network, sigma = exp(-4): long episodes, few resets per step
        params[:-28] = 0.0
        params[-28:-24] = -4.0
    for rnd in range(2):
        for form in forms:
            os.environ["QS_TC_FORM"] = form
            st = eng.new_state()
            eng.reset(st)
            buf = eng.rollout_policy(st, params, T=T, t0=0, dist=0, tensor_cores=True)
            torch.cuda.synchronize()
            reps = 4
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                eng.rollout_policy(st, params, T=T, t0=T * (r + 1), dist=0, buffers=buf, tensor_cores=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fin = float(buf["done"].sum() + buf["trunc"].sum()) / (nb * T)
            print(f"{nb} envs x {T} steps, form {form}: {ms:.3f} ms, {nb * T / ms / 1e3:.3e} M env-steps/s, episodes finished per env-step {fin:.4f}", flush=True)
    del eng
