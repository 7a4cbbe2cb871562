from .agent import Agent  # noqa: F401
from .nets import Actor, Critic, TanhGaussActor  # noqa: F401
