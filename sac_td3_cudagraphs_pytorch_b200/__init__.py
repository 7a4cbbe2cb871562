"""b2rl — a B200-native SAC/TD3 learner update.

Host side: ``agents.Agent`` (the reference's agent API), ``replay.ReplayBuffer`` (the torchrl
call surface the reference uses) and ``engine.LearnerEngine`` (whole learner iterations as CUDA
graphs). Device side: ``libb2rl.so`` (csrc/*.cu, C ABI in include/b2rl.h), loaded with ctypes.
"""
from .hps import Hps, load_hps, sac_hps, td3_hps  # noqa: F401

__all__ = ["Hps", "load_hps", "sac_hps", "td3_hps", "Agent", "ReplayBuffer", "LearnerEngine"]


def __getattr__(name):  # heavy pieces are imported lazily so that `import package` works without CUDA
    if name == "Agent":
        from .agents.agent import Agent
        return Agent
    if name == "ReplayBuffer":
        from .replay import ReplayBuffer
        return ReplayBuffer
    if name == "LearnerEngine":
        from .engine import LearnerEngine
        return LearnerEngine
    raise AttributeError(name)
