"""quadsim-b200: B200-native batched quadrotor env step / rollout / GAE behind the reference's env APIs.

Public surface:
    config.EnvConfig          the five reference env variants + north-star default
    engine.Engine             ctypes binding of libquadsim.so (C ABI, include/quadsim_abi.h)
    brax_env.JaxMJXQuadBraxEnv / QuadHoverBraxEnv   Brax-protocol adapter (State layout)
    gym_vec.HoverVecEnv       Gymnasium VectorEnv / SB3 VecEnv facade
    ppo.PPOTrainer            data-parallel PPO around the fused rollout / GAE kernels
Importing the package never touches the GPU and never loads the oracle.
"""
from . import config, model, trajectories  # noqa: F401

__version__ = "0.1.0"
