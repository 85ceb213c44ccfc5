"""Small run of every kernel for compute-sanitizer (memcheck / racecheck), sized to finish quickly under the tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from uav_reinforcement_learning_control_b200 import config as Q
from uav_reinforcement_learning_control_b200 import trajectories as TJ
from uav_reinforcement_learning_control_b200.engine import Engine

sys.path.insert(0, ROOT)
from bench import make_policy_params


def run(cfg, n, tc=False, first=False, rollout=True):
    eng = Engine(cfg, n, device=0)
    st = eng.new_state()
    fs = torch.zeros(21, n, device="cuda") if first else None
    eng.reset(st, first_state=fs)
    act = torch.rand(n, 4, device="cuda") * 2 - 1
    trunc = torch.zeros(n, device="cuda"); met = torch.zeros(4, n, device="cuda")
    term = torch.zeros(n, cfg.obs_dim, device="cuda")
    for _ in range(3):
        eng.step(st, act, truncated=trunc, metrics=met, terminal_obs=term, first_state=fs)
    if fs is None:
        for _ in range(3):
            eng.step(st, act)                       # lean instantiation where the configuration allows it
    if cfg.mode == Q.MODE_TRAJ_GYM:
        eng.traj_info(st)
    eng.observe(st, act)
    stats = torch.zeros(4, n, device="cuda")
    eng.rollout_random(st, 5, t0=0, stats=stats, first_state=fs)
    if rollout:
        dist = 1 if cfg.mode == Q.MODE_MJX_BRAX else 0
        p = make_policy_params(eng, torch, torch.device("cuda"), seed=0, dist=dist)
        b = eng.rollout_policy(st, p, T=3, dist=dist, first_state=fs, bootstrap_gamma=0.9 if dist == 0 else 0.0)
        eng.gae(b["reward"], b["value"], b["done"], b["trunc"], b["last_value"], 0.99, 0.95, brax_form=bool(dist))
        if tc:
            eng.rollout_policy(st, p, T=3, dist=0, tensor_cores=True, bootstrap_gamma=0.9)
    torch.cuda.synchronize()
    eng.close()


run(Q.EnvConfig.north_star(max_episode_steps=4), 300, tc=True)
if not os.environ.get("QS_SANITIZE_SMALL"):
    run(Q.EnvConfig.north_star(max_episode_steps=4), 40000, tc=True)      # 2-tile tcgen05 variant
run(Q.EnvConfig.hover_gym(rate_wrapper=True, auto_reset=Q.RESET_RESAMPLE, max_episode_steps=3), 130)
run(Q.EnvConfig.mjx_brax(episode_length=3, auto_reset=Q.RESET_RESTORE_FIRST), 130, first=True)
run(Q.EnvConfig.traj_gym(auto_reset=Q.RESET_RESAMPLE, max_episode_steps=3), 200)
run(Q.EnvConfig.hover_brax(), 70, rollout=False)
run(Q.EnvConfig.mjx_playground(), 70, rollout=False)
run(Q.EnvConfig.waypoint_eval(TJ.default_tables(0.5), auto_reset=Q.RESET_RESAMPLE), 100, rollout=True)
print("sanitize smoke done")
