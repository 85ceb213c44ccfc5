"""Same-box A/B of library builds for the PPO gradient kernel (run on the GPU box): for each
build/variants/libquadsim_*.so run the fused-gradient tests and time `qs_ppo_grad_packed` + `qs_ppo_adam` on
2^20-sample minibatches (tools/ppo_update_bench.py's loop).  One summary line per variant."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMER = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
from uav_reinforcement_learning_control_b200.ppo import ActorCritic, FusedUpdater
N = 8192 * 1024; mb = N // 8; dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
obs = torch.rand(N, 12, device=dev, generator=g) * 2 - 1
act = torch.randn(N, 4, device=dev, generator=g) * 0.4
old_logp = torch.randn(N, device=dev, generator=g) * 0.1 - 1.0
adv = torch.randn(N, device=dev, generator=g); ret = torch.randn(N, device=dev, generator=g)
params = ActorCritic(12, dev, seed=0, log_std_init=-1.0).pack()
up = FusedUpdater(dev)
perm = torch.randperm(N, device=dev, generator=g).to(torch.int32)
packed = up.pack(obs, act, old_logp, adv, ret)
def fused(k):
    up.grad(params, obs, act, old_logp, adv, ret, idx=perm[k * mb:(k + 1) * mb], clip_range=0.19, vf_coef=0.5, ent_coef=1e-4, packed=packed)
    up.adam(params, 1.5e-4)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
res = []
for rep in range(3):
    for k in range(3): fused(k)
    torch.cuda.synchronize(); ev[0].record()
    for r in range(4):
        for k in range(8): fused(k)
    ev[1].record(); torch.cuda.synchronize()
    res.append(ev[0].elapsed_time(ev[1]) / 32)
# the trainer's form: one native call per epoch of 8 minibatches (qs_ppo_update_epoch)
acc = torch.zeros(up.N_STATS, device=dev)
res2 = []
if hasattr(up, "update_epoch"):
    for rep in range(3):
        up.update_epoch(params, packed, adv, perm, 8, 1.5e-4, clip_range=0.19, vf_coef=0.5, ent_coef=1e-4, stats_acc=acc)
        torch.cuda.synchronize(); ev[0].record()
        for r in range(4):
            up.update_epoch(params, packed, adv, perm, 8, 1.5e-4, clip_range=0.19, vf_coef=0.5, ent_coef=1e-4, stats_acc=acc)
        ev[1].record(); torch.cuda.synchronize()
        res2.append(ev[0].elapsed_time(ev[1]) / 32)
print(json.dumps({"ms_per_minibatch": res, "epoch_call_ms_per_minibatch": res2, "finite": bool(torch.isfinite(params).all())}))
''' % ROOT

tests = ["tests/test_ppo_update.py", "-q", "-x", "-m", "gpu", "-k",
         "fused_gradient_matches_oracle or bitwise or fused_update_follows or buffer_bounds"]
only = sys.argv[1:]
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libquadsim_*.so"))):
    name = os.path.basename(lib)[len("libquadsim_"):-3]
    if only and name not in only:
        continue
    env = dict(os.environ, QS_LIB_PATH=lib)
    t = subprocess.run([sys.executable, "-m", "pytest"] + tests, env=env, capture_output=True, text=True, cwd=ROOT)
    tl = (t.stdout.strip().splitlines() or ["?"])[-1]
    if os.environ.get("PPO_AB_PROFILE"):
        print(t.stdout[-3000:])
    b = subprocess.run([sys.executable, "-c", TIMER], env=env, capture_output=True, text=True, cwd=ROOT)
    bl = [l for l in b.stdout.splitlines() if l.startswith("{")]
    prof = [l for l in b.stdout.splitlines() if l.startswith("ppoprof")]
    print(f"{name:12s} tests: {tl} | bench: {bl[-1] if bl else 'FAILED ' + b.stderr[-600:]}", flush=True)
    for l in prof[-8:]:
        print("    ", l)
