"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference's SAC/TD3 update.

Nothing in the product package (``sac_td3_cudagraphs_pytorch_b200``) may import
this package. It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it.

Parity status: PINNED by outputs of the reference itself. The reference ships
no tests or golden vectors, so ``tests/golden/make_golden.py``
imports the reference's unmodified ``agents/agent.py`` + ``agents/nets.py`` from
``/root/reference`` (with tiny stand-ins for the absent ``tensordict`` /
``torchrl`` / ``omegaconf`` containers) and records its outputs; the committed
fixtures in ``tests/golden/*.npz`` are what the oracle is checked against.
"""
from .sac_td3_oracle import (  # noqa: F401
    OracleAgent,
    OracleHps,
    PARAM_NAMES,
    init_mlp_params,
    make_synthetic_transitions,
    philox4x32_10,
    philox_normal_pairs,
    philox_randint,
    sac_defaults,
    td3_defaults,
)
