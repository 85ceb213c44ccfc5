"""Generates tests/golden/reference_utils.json by IMPORTING the reference's own utilities.

Run in the dev container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Only the parts of the reference that import without MuJoCo/JAX can be executed: utils/trajectories.py,
utils/normalization.py, utils/state.py (scipy Euler convention) and utils/drone_config.py.  A 12-line
stub of ``gymnasium.spaces.Box`` is put on sys.path because normalization.py / state.py import it for
type hints and ``.low/.high`` only.  Nothing from the reference is copied into the repo: the JSON holds
inputs and the reference's OUTPUTS.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_utils.json")


def main():
    gym = types.ModuleType("gymnasium"); spaces = types.ModuleType("gymnasium.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype); self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape; self.dtype = dtype
    spaces.Box = Box; gym.spaces = spaces
    sys.modules["gymnasium"] = gym; sys.modules["gymnasium.spaces"] = spaces
    sys.path.insert(0, REF)
    from utils import drone_config
    from utils.normalization import denormalize, normalize
    from utils.state import QuadState
    from utils.trajectories import TRAJECTORY_GENERATORS

    rng = np.random.default_rng(20261018)
    out = {"source": "Karl-Liu-ch/uav_reinforcement_learning_control utils/*, imported from /root/reference"}
    out["drone_config"] = {k: getattr(drone_config, k) for k in
                           ("MAX_MOTOR_THRUST", "ARM_LENGTH", "YAW_TORQUE_COEFF", "MASS", "G", "DT", "IXX", "IYY", "IZZ",
                            "MAX_TOTAL_THRUST", "MAX_TORQUE", "HOVER_THRUST_PER_MOTOR")}
    out["waypoints"] = {}
    for name, fn in TRAJECTORY_GENERATORS.items():
        for spacing in (0.5, 0.3):
            out["waypoints"][f"{name}@{spacing}"] = [w.tolist() for w in fn(spacing=spacing)]
    out["waypoints"]["circle@0.5,r=0.7,c=(0.2,-0.1,1.3)"] = [w.tolist() for w in TRAJECTORY_GENERATORS["circle"](
        spacing=0.5, radius=0.7, center=np.array([0.2, -0.1, 1.3]))]
    # observation normalisation (hover_env.py:36-39 bounds) and action denormalisation (:60-65)
    obs_b = Box(np.array([-4, -4, -2, -np.pi, -np.pi, -np.pi, -10, -10, -10, -6 * np.pi, -6 * np.pi, -6 * np.pi], dtype=np.float32),
                np.array([4, 4, 2, np.pi, np.pi, np.pi, 10, 10, 10, 6 * np.pi, 6 * np.pi, 6 * np.pi], dtype=np.float32))
    act_b = Box(np.array([0.0, -0.5, -0.5, -0.5], dtype=np.float32), np.array([52.0, 0.5, 0.5, 0.5], dtype=np.float32))
    x = (rng.uniform(-1.2, 1.2, (64, 12)) * obs_b.high).astype(np.float32)
    x[:2] = [obs_b.low, obs_b.high]
    out["normalize"] = {"x": x.tolist(), "y": [normalize(r, obs_b).astype(np.float32).tolist() for r in x]}
    a = rng.uniform(-1.3, 1.3, (64, 4)).astype(np.float32)
    a[:3] = [[0, 0, 0, 0], [-1, -1, -1, -1], [1, 1, 1, 1]]
    out["denormalize"] = {"a": a.tolist(), "u": [denormalize(r, act_b).astype(np.float64).tolist() for r in a]}
    # quaternion <-> Euler (utils/state.py:28-65)
    q = rng.normal(size=(64, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = [1, 0, 0, 0]
    qpos = np.concatenate([rng.uniform(-2, 2, (64, 3)), q], axis=1); qvel = rng.uniform(-5, 5, (64, 6))
    st = []
    for i in range(64):
        s = QuadState(); s.set_from_mujoco(qpos[i], qvel[i]); st.append(s.vec().tolist())
    out["set_from_mujoco"] = {"qpos": qpos.tolist(), "qvel": qvel.tolist(), "state12": st}
    e = rng.uniform(-0.3, 0.3, (64, 3)); e[0] = 0; e[1] = [3.0, 1.5, -3.0]
    back = []
    for i in range(64):
        s = QuadState(); s.state[3:6] = e[i]; qp, qv = s.get_mujoco_state(); back.append(qp[3:7].tolist())
    out["get_mujoco_state"] = {"rpy": e.astype(np.float32).tolist(), "quat_wxyz": back}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
