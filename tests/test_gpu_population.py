"""Population (BASELINE.json config 4): N stacked learners == N separate learners, bitwise, and the result
does not depend on how the agent ids are sharded (the multi-GPU story: a partition, no collective)."""
import numpy as np
import pytest
import torch

from oracle import make_synthetic_transitions

pytestmark = pytest.mark.gpu


def _hps(algo):
    from sac_td3_cudagraphs_pytorch_b200 import sac_hps, td3_hps
    return (sac_hps if algo == "sac" else td3_hps)(batch_size=32)


def _data(agent_id, n=600, ob=11, ac=3):
    return make_synthetic_transitions(n, ob, ac, [-1.0] * ac, [1.0] * ac, seed=77 + agent_id)


def _population(ids, algo, use_graphs=True):
    from sac_td3_cudagraphs_pytorch_b200.population import Population
    pop = Population(ids, 11, 3, [-1.0] * 3, [1.0] * 3, _hps(algo), "cuda", seed=42, rb_capacity=1000,
                     use_graphs=use_graphs)
    for g, aid in enumerate(ids):
        pop.fill_replay(_data(aid), agent=g)
    return pop


@pytest.mark.parametrize("algo", ["sac", "td3"])
def test_population_equals_separate_agents(algo):
    from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
    from sac_td3_cudagraphs_pytorch_b200.engine import LearnerEngine
    from sac_td3_cudagraphs_pytorch_b200.population import init_agent_params
    from sac_td3_cudagraphs_pytorch_b200.replay import ReplayBuffer
    ids, n_it = [5, 6, 7], 7
    pop = _population(ids, algo)
    for i in range(n_it):
        pop.iteration()
    torch.cuda.synchronize()
    hps = _hps(algo)
    lo, hi = torch.full((3,), -1.0), torch.full((3,), 1.0)
    for g, aid in enumerate(ids):
        rb = ReplayBuffer(1000, "cuda", seed=42, agent_id=aid)
        rb.extend({k: v.cuda() for k, v in _data(aid).items()})
        ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, lo.numpy(), hi.numpy(), torch.device("cuda"), hps, rb=rb,
                   seed=42, agent_id=aid)
        ag.load_params(*init_agent_params(aid, 42, 11, 3, algo == "td3", True, lo, hi))
        eng = LearnerEngine(ag, use_graphs=False)
        for i in range(n_it):
            eng.iteration(i)
        torch.cuda.synchronize()
        assert torch.equal(pop.arena.flat[g], ag.arena.flat[0]), f"agent {aid}: stacked != separate"
        assert torch.equal(pop.counters[g, :3], ag.counters[:3])
        assert torch.equal(pop.alpha_state[g, :4], ag._alpha_state[:4])
        assert torch.equal(pop.out[g, :4], ag.out[:4])
    # agents are independent: different data and seeds give different parameters
    assert not torch.equal(pop.arena.flat[0], pop.arena.flat[1])
    assert torch.isfinite(pop.out).all()


def test_population_is_invariant_to_sharding():
    whole = _population([0, 1, 2, 3], "sac")
    parts = [_population([0, 1], "sac"), _population([2, 3], "sac", use_graphs=False)]
    for i in range(6):
        whole.iteration()
        for p in parts:
            p.iteration()
    torch.cuda.synchronize()
    got = torch.cat([p.arena.flat for p in parts])
    assert torch.equal(whole.arena.flat, got)
    assert torch.equal(whole.idx, torch.cat([p.idx for p in parts]))  # same index draws per global agent id
