"""The oracle against the golden fixtures recorded from the reference itself.

``tests/golden/*.npz`` hold the outputs of the reference's unmodified agents/agent.py
(see tests/golden/make_golden.py). On the authoring machine the oracle reproduced them
bit for bit (``meta.oracle_vs_reference_max_rel_dev == 0``); on another CPU, BLAS kernel
selection may move fp32 results in the last bits, and a few Adam steps amplify that by
about the reference's own fp32-vs-fp64 gap, which each fixture records. Tolerance here:
max(2e-5, 4 x that recorded gap), relative to each tensor's max.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.golden import portable as P
from tests.golden.cases import CASES, case_inputs, inputs_digest
from tests.golden.make_golden import GROUPS, run_oracle

GOLD = Path(__file__).parent / "golden"


def load(name):
    z = np.load(GOLD / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.mark.parametrize("name", list(CASES))
def test_fixture_was_bit_exact_when_recorded(name):
    _, meta = load(name)
    assert meta["oracle_vs_reference_max_rel_dev"] == 0.0
    assert "reference agents/agent.py" in meta["source"]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_outputs(name):
    torch.set_num_threads(1)
    z, meta = load(name)
    rec = run_oracle(case_inputs(name))
    tol = max(2e-5, 4 * meta["reference_fp32_vs_fp64_oracle"])
    want_logs = z["logs"]
    m = ~np.isnan(want_logs)
    assert (np.isnan(rec["logs"]) == np.isnan(want_logs)).all()
    rel = np.abs(rec["logs"][m] - want_logs[m]) / np.maximum(np.abs(want_logs[m]), 1e-30)
    assert rel.max() <= tol, f"loss trajectory off by {rel.max():.3e}"
    checked = 0
    for g in GROUPS:
        for n, t in rec.get(g, {}).items():
            ok, e = P.summary_close(P.summarize(t), z[f"{g}/{n}"], tol)
            assert ok, f"{name}: {g}/{n} off by {e:.3e} (tol {tol:.1e})"
            checked += 1
    assert checked >= 30


@pytest.mark.parametrize("name", list(CASES))
def test_fixture_inputs_are_reproducible(name):
    """Inputs are regenerated from seeds; their digest must equal the one recorded with the outputs."""
    _, meta = load(name)
    assert inputs_digest(case_inputs(name)) == meta["inputs_sha256"]
