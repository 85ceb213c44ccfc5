"""Horizon-resolved error of the float32 sm_100a step kernel against the float64 oracle over full 512-step
episodes (BASELINE.json north_star: "a horizon-resolved error curve is reported over full 512-step episodes").

Both integrators start from the same Philox-reset states and receive the SAME action sequence: a cascaded PD hover
controller evaluated on the ORACLE's state (random actions terminate an episode within a few steps, which would leave
no horizon to resolve).  HoverEnv semantics with battery sag, no auto-reset.  Writes profiles/horizon_error_r01.json.
Runs on the GPU box:  python tools/horizon_error.py [--envs 256] [--steps 512]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pd_hover_action(qpos, qvel, target, mass=0.22274432, inertia=(4.93e-4, 5.00e-4, 6.40e-4)):
    """Normalised [thrust, tau_x, tau_y, tau_z] action that flies to `target` and holds it (float64 in, float32 out)."""
    w, x, y, z = qpos[:, 3], qpos[:, 4], qpos[:, 5], qpos[:, 6]
    roll = np.arctan2(2 * (y * z + w * x), 1 - 2 * (x * x + y * y))
    pitch = np.arcsin(np.clip(2 * (w * y - x * z), -1, 1))
    yaw = np.arctan2(2 * (x * y + w * z), 1 - 2 * (y * y + z * z))
    e = target - qpos[:, 0:3]
    ax = np.clip(1.2 * e[:, 0] - 1.6 * qvel[:, 0], -3, 3); ay = np.clip(1.2 * e[:, 1] - 1.6 * qvel[:, 1], -3, 3)
    az = np.clip(4.0 * e[:, 2] - 3.0 * qvel[:, 2], -4, 6)
    pitch_des = np.clip((ax * np.cos(yaw) + ay * np.sin(yaw)) / 9.81, -0.35, 0.35)
    roll_des = np.clip((ax * np.sin(yaw) - ay * np.cos(yaw)) / 9.81, -0.35, 0.35)
    thrust = mass * (9.81 + az) / np.maximum(np.cos(roll) * np.cos(pitch), 0.5)
    tau = np.stack([inertia[0] * (180.0 * (roll_des - roll) - 26.0 * qvel[:, 3]),
                    inertia[1] * (180.0 * (pitch_des - pitch) - 26.0 * qvel[:, 4]),
                    inertia[2] * (80.0 * (0.0 - yaw) - 30.0 * qvel[:, 5])], axis=1)
    a = np.concatenate([(2.0 * thrust / 52.0 - 1.0)[:, None], tau / 0.5], axis=1)
    return np.clip(a, -1, 1).astype(np.float32)


def run(n_envs=256, steps=512, seed=0, device=0):
    import torch
    from oracle.envs import OracleEnv
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200 import model as M
    from uav_reinforcement_learning_control_b200.engine import Engine
    cfg = Q.EnvConfig.hover_gym(seed=seed)                         # battery on, auto_reset off, 512-step episodes
    eng = Engine(cfg, n_envs, device=device)
    st = eng.new_state()
    eng.reset(st)
    torch.cuda.synchronize()
    s = OracleEnv.from_planes(st.cpu().numpy())
    orc = OracleEnv(M.load_mjcf(M.default_model_path()), cfg)
    curve = {k: [] for k in ("pos_max", "pos_med", "quat_max", "vel_max", "omega_max", "obs_max", "reward_max")}
    flags_equal, alive = True, np.ones(n_envs, bool)
    for t in range(steps):
        act = pd_hover_action(s["qpos"], s["qvel"], s["target"].astype(np.float64))
        o = orc.step(s, act)
        trunc = torch.zeros(n_envs, device=eng.device)
        obs, rew, done = eng.step(st, torch.from_numpy(act).to(eng.device), truncated=trunc)
        g = st.cpu().numpy()
        d_done = done.cpu().numpy(); d_trunc = trunc.cpu().numpy()
        flags_equal &= bool(np.array_equal(d_done, o["done"]) and np.array_equal(d_trunc, o["truncated"]))
        alive &= (o["done"] == 0)
        a = alive
        if not a.any():
            break
        dp = np.linalg.norm(g[0:3].T[a] - s["qpos"][a, 0:3], axis=1)
        dq = np.minimum(np.abs(g[3:7].T[a] - s["qpos"][a, 3:7]).max(1), np.abs(g[3:7].T[a] + s["qpos"][a, 3:7]).max(1))
        curve["pos_max"].append(float(dp.max())); curve["pos_med"].append(float(np.median(dp)))
        curve["quat_max"].append(float(dq.max()))
        curve["vel_max"].append(float(np.abs(g[11:14].T[a] - s["qvel"][a, 0:3]).max()))
        curve["omega_max"].append(float(np.abs(g[14:17].T[a] - s["qvel"][a, 3:6]).max()))
        curve["obs_max"].append(float(np.abs(obs.cpu().numpy()[a] - o["obs"][a]).max()))
        curve["reward_max"].append(float(np.abs(rew.cpu().numpy()[a] - o["reward"][a]).max()))
    final_err = float(np.linalg.norm(s["qpos"][alive, 0:3] - s["target"][alive], axis=1).mean()) if alive.any() else None
    return {"n_envs": n_envs, "steps": len(curve["pos_max"]), "alive_at_end": int(alive.sum()), "flags_bit_exact": flags_equal,
            "mean_distance_to_target_at_end_m": final_err, "curve": curve}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "horizon_error_r01.json"))
    a = ap.parse_args()
    res = run(a.envs, a.steps)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f)
    c = res["curve"]
    for t in (0, 9, 63, 127, 255, 511):
        if t < res["steps"]:
            print(f"t={t + 1:4d}  pos_max {c['pos_max'][t]:.3e}  quat {c['quat_max'][t]:.3e}  vel {c['vel_max'][t]:.3e}  "
                  f"omega {c['omega_max'][t]:.3e}  obs {c['obs_max'][t]:.3e}")
    print({k: v for k, v in res.items() if k != "curve"})
