"""BASELINE.json configs[3] end to end on one B200: train a hover policy with the fused PPO pipeline (tcgen05 rollout ->
GAE -> qs_ppo_grad / qs_ppo_adam), then fly it deterministically over the circle / figure-8 / square waypoint tables of
utils/trajectories.py with evaluate.py's advance rule (reach radius 0.25 m, lap = all waypoints) fused in the step, on
262 144 envs.  Prints one JSON line: waypoints reached, laps, crashes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import trajectories as TJ
from uav_reinforcement_learning_control_b200.engine import Engine
from uav_reinforcement_learning_control_b200.ppo import PPOConfig, PPOTrainer

iters = int(os.environ.get("DEMO_ITERS", 300))
torch.manual_seed(0)
eng = Engine(Q.EnvConfig.north_star(seed=0), 65536, device=0)
tr = PPOTrainer(eng, PPOConfig(n_steps=64, learning_rate=3e-4, ent_coef=0.0), seed=0)
tr.set_log_std(-1.0)
torch.cuda.synchronize(); t0 = time.time()
log = tr.train(iters)
torch.cuda.synchronize(); t_train = time.time() - t0
params = tr.packed_params().clone()

nb, T, reps = 1 << 18, 128, 16                       # 2048 steps = 20.5 s of flight per env
names = ("eight", "circle", "square")
cfg = Q.EnvConfig.waypoint_eval(TJ.default_tables(0.5), auto_reset=Q.RESET_NONE, seed=4)
ev = Engine(cfg, nb, device=0)
st = ev.new_state(); ev.reset(st)
buf = None
dead = torch.zeros(nb, dtype=torch.bool, device="cuda")
torch.cuda.synchronize(); t0 = time.time()
for r in range(reps):
    buf = ev.rollout_policy(st, params, T=T, t0=r * T, dist=0, deterministic=True, tensor_cores=True, buffers=buf)
    dead |= (buf["done"] != 0).any(dim=0)
torch.cuda.synchronize(); t_eval = time.time() - t0
reached = st[29].view(torch.int32); laps = st[30].view(torch.int32)
shape = torch.arange(nb, device="cuda") % 3
out = {"train_iters": iters, "train_seconds": t_train, "train_env_steps": 65536 * 64 * iters,
       "episodes_per_rollout_first_last": [log[0]["episodes"], log[-1]["episodes"]],
       "eval_envs": nb, "eval_steps": T * reps, "eval_seconds": t_eval, "eval_env_steps_per_s": nb * T * reps / t_eval,
       "left_bounds_fraction": float(dead.float().mean()), "per_shape": {}}
for k, nm in enumerate(names):
    m = (shape == k) & ~dead
    out["per_shape"][nm] = {"waypoints": len(TJ.default_tables(0.5)[k]),
                            "mean_waypoints_reached": float(reached[m].float().mean()) if bool(m.any()) else None,
                            "mean_laps": float(laps[m].float().mean()) if bool(m.any()) else None,
                            "max_laps": int(laps[m].max()) if bool(m.any()) else None}
print(json.dumps(out))
