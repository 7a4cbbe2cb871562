"""Stand-in for tensordict.nn (generator-only)."""
from . import TensorDict


class TensorDictModule:
    """fn(*[td[k] for k in in_keys]) -> dict; selected out_keys are written back (agent.py:74-95)."""

    def __init__(self, module, in_keys, out_keys):
        self.module, self.in_keys, self.out_keys = module, list(in_keys), list(out_keys)

    def __call__(self, td):
        res = self.module(*[td[k] for k in self.in_keys])
        for k in self.out_keys:
            td[k] = res[k]
        return td


class CudaGraphModule:  # imported by orchestrator.py only; never used by the generator
    def __init__(self, fn, in_keys=None, out_keys=None):
        self.fn = fn

    def __call__(self, *a, **k):
        return self.fn(*a, **k)
