"""Scripted policies expressed as weights of the 2x128 ReLU actor-critic the rollout kernels run.

``pd_waypoint_policy`` is a cascaded PD position / attitude controller written into the packed parameter vector of
``qs_policy_param_count`` (include/quadsim_abi.h), so that the waypoint-tracking workload of BASELINE.json configs[3]
(evaluate.py:440-557: circle / figure-8 / square tables, reach radius 0.25 m) can be flown by the very kernels that are
being measured -- fp32 FMA path or tcgen05 path -- without a trained checkpoint.  It replaces the *policy* the
reference evaluates there (``PPO.load(...)``, evaluate.py:459) by a deterministic stand-in; the reference's own PID
controllers (controllers/) are out of scope.

The controller acts on the 12-D normalised observation of HoverEnv (hover_env.py:126-136): obs = x / bound with
x = [target - pos (world), roll, pitch, yaw, v (world), omega (body)] and bounds [4,4,2, pi,pi,pi, 10,10,10, 6pi,6pi,6pi].
A ReLU network represents a linear map u exactly as relu(u) - relu(-u) and a saturation as
clip(u, -c, c) = relu(u + c) - relu(u - c) - c, which is all a cascaded PD loop needs:

  layer 1:  desired accelerations a_x, a_y (saturated), a_z (saturated); pass-through of rpy and omega
  layer 2:  thrust = m (g + a_z);  roll_des = -a_y / g, pitch_des = a_x / g (yaw held at 0, small-angle);
            tau = I (kp_att (att_des - att) - kd_att omega), each as a +/- pair
  head:     action = [2 thrust / 52 - 1, tau / 0.5]  (denormalised by hover_env.py:169)
"""
from __future__ import annotations

import math

import numpy as np

H, A = 128, 4

# hover_env.py:36-39 observation bounds
_OBS_BOUND = np.array([4, 4, 2, math.pi, math.pi, math.pi, 10, 10, 10, 6 * math.pi, 6 * math.pi, 6 * math.pi], dtype=np.float64)


def pd_waypoint_policy(kp_xy=6.0, kd_xy=4.0, kp_z=25.0, kd_z=9.0, a_max_xy=3.5, a_max_z=6.0,
                       kp_att=180.0, kd_att=26.0, kp_yaw=40.0, kd_yaw=12.0, log_std=-5.0,
                       mass=0.22274432, inertia=(4.97e-4, 5.04e-4, 6.40e-4), g=9.81,
                       max_total_thrust=52.0, max_torque=0.5):
    """float32 packed parameter vector (obs_dim 12, dist 0: SB3 Gaussian head with state-independent log_std)."""
    D = 12
    W1 = np.zeros((D, H)); b1 = np.zeros(H); W2 = np.zeros((H, H)); b2 = np.zeros(H); W3 = np.zeros((H, A)); b3 = np.zeros(A)
    B = _OBS_BOUND

    def lin(weights):
        """row of W1 for the physical linear form sum_k w_k x_k, given {obs index: weight on the physical quantity}"""
        col = np.zeros(D)
        for k, w in weights.items():
            col[k] = w * B[k]                      # obs_k = x_k / B_k
        return col

    # ---- layer 1 -----------------------------------------------------------------------------------------------
    u = 0
    sat = {}                                       # name -> (unit of relu(a + c), unit of relu(a - c), c)
    for name, wts, c in (("ax", {0: kp_xy, 6: -kd_xy}, a_max_xy), ("ay", {1: kp_xy, 7: -kd_xy}, a_max_xy),
                         ("az", {2: kp_z, 8: -kd_z}, a_max_z)):
        W1[:, u] = lin(wts); b1[u] = c
        W1[:, u + 1] = lin(wts); b1[u + 1] = -c
        sat[name] = (u, u + 1, c)
        u += 2
    passthru = {}                                  # name -> (unit of relu(x), unit of relu(-x))
    for name, k in (("roll", 3), ("pitch", 4), ("yaw", 5), ("wx", 9), ("wy", 10), ("wz", 11)):
        W1[:, u] = lin({k: 1.0}); W1[:, u + 1] = lin({k: -1.0})
        passthru[name] = (u, u + 1)
        u += 2

    def add_sat(col, name, w):                     # col += w * clip(a_name): returns the constant to put in the bias
        p, m, c = sat[name]
        col[p] += w; col[m] -= w
        return -w * c

    def add_pt(col, name, w):
        p, m = passthru[name]
        col[p] += w; col[m] -= w

    # ---- layer 2: the four commands, each as a +/- pair so the head can be linear ---------------------------------
    Ixx, Iyy, Izz = inertia
    cmds = []
    # thrust (normalised): 2 m (g + a_z) / T_max - 1
    col = np.zeros(H); bias = 2.0 * mass * g / max_total_thrust - 1.0
    bias += add_sat(col, "az", 2.0 * mass / max_total_thrust)
    cmds.append((col, bias))
    # roll torque: Ixx (kp_att (-a_y / g - roll) - kd_att wx) / tau_max
    col = np.zeros(H); bias = add_sat(col, "ay", -Ixx * kp_att / g / max_torque)
    add_pt(col, "roll", -Ixx * kp_att / max_torque); add_pt(col, "wx", -Ixx * kd_att / max_torque)
    cmds.append((col, bias))
    # pitch torque: Iyy (kp_att (a_x / g - pitch) - kd_att wy) / tau_max
    col = np.zeros(H); bias = add_sat(col, "ax", Iyy * kp_att / g / max_torque)
    add_pt(col, "pitch", -Iyy * kp_att / max_torque); add_pt(col, "wy", -Iyy * kd_att / max_torque)
    cmds.append((col, bias))
    # yaw torque: Izz (-kp_yaw yaw - kd_yaw wz) / tau_max
    col = np.zeros(H); bias = 0.0
    add_pt(col, "yaw", -Izz * kp_yaw / max_torque); add_pt(col, "wz", -Izz * kd_yaw / max_torque)
    cmds.append((col, bias))
    for j, (col, bias) in enumerate(cmds):
        W2[:, 2 * j] = col; b2[2 * j] = bias
        W2[:, 2 * j + 1] = -col; b2[2 * j + 1] = -bias
        W3[2 * j, j] = 1.0; W3[2 * j + 1, j] = -1.0

    critic = [np.zeros(D * H), np.zeros(H), np.zeros(H * H), np.zeros(H), np.zeros(H), np.zeros(1)]
    parts = [W1.reshape(-1), b1, W2.reshape(-1), b2, W3.reshape(-1), b3] + critic + \
            [np.full(A, float(log_std)), np.zeros(D), np.ones(D)]
    return np.concatenate(parts).astype(np.float32)
