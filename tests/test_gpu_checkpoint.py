"""Checkpoint round trip: ``Agent.save`` / ``load_from_disk`` keep the reference's on-disk layout
(agents/agent.py:333-371: a torch.save'd dict with keys hps / timesteps_so_far / actor / qnet1 / qnet2 / actor_optimizer /
q_optimizer, state_dicts with nets.py's parameter names, Adam-shaped optimizer state) and restore the learner exactly:
save -> fresh Agent -> load_from_disk -> bitwise-equal arena, counters and temperature state, and training continues
bit-identically."""
import pytest
import torch

from tests.golden.cases import CASES, case_inputs
from tests.helpers import batch_of, make_agent

pytestmark = pytest.mark.gpu

REF_TOP_KEYS = {"hps", "timesteps_so_far", "actor", "qnet1", "qnet2", "actor_optimizer", "q_optimizer"}  # agent.py:343-352
NET_KEYS = [f"fc_stack.fc_block_{i}.{m}.{p}" for i in (1, 2) for m in ("fc", "ln") for p in ("weight", "bias")] + \
           ["head.weight", "head.bias"]  # agents/nets.py:66-84


def _train(ag, inp, its):
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    for i in its:
        ag.update_qnets(dev(batch_of(inp, i)), eps=inp["eps_q"][i].cuda())
        ag.qnet_updates_so_far += 1
        if i % 3 == 0:
            for j in range(2):
                ag.update_actor(dev(batch_of(inp, i)), eps=inp["eps_pi"][i][j].cuda(), eps_alpha=inp["eps_alpha"][i][j].cuda())
        ag.update_targ_nets()
    torch.cuda.synchronize()


@pytest.mark.parametrize("name", ["sac_hopper", "td3_hopper"])
def test_checkpoint_round_trip(name, tmp_path):
    inp = case_inputs(CASES[name])
    ag = make_agent(inp)
    _train(ag, inp, range(0, 4))
    ag.timesteps_so_far = 1234
    path = ag.save(tmp_path, sfx="best")
    assert path.name == "ckpt_best.pth"  # agents/agent.py:337-341

    ck = torch.load(path, weights_only=False)
    assert REF_TOP_KEYS <= set(ck), sorted(ck)
    for k in ("qnet1", "qnet2"):
        assert list(ck[k]) == NET_KEYS
    assert [k for k in ck["actor"] if k.startswith(("fc_stack", "head"))] == NET_KEYS
    assert {"action_scale", "action_bias"} <= set(ck["actor"])  # nets.py buffers
    # shapes are torch's ([out, in] weights) and the optimizer state loads into a stock torch.optim.Adam
    O, A = inp["ob"], inp["ac"]
    assert ck["qnet1"]["fc_stack.fc_block_1.fc.weight"].shape == (256, O + A)
    assert ck["qnet1"]["head.weight"].shape == (1, 256)
    stock_params = [torch.nn.Parameter(torch.zeros_like(v)) for v in ag.actor_params.values()]
    stock = torch.optim.Adam(stock_params, lr=1.0)
    stock.load_state_dict(ck["actor_optimizer"])
    assert stock.param_groups[0]["lr"] == pytest.approx(float(inp["hps"]["actor_lr"]))
    # the critics on disk are the LIVE ones (documented deviation: the reference stores the stale pre-stack modules)
    assert torch.equal(ck["qnet1"]["head.weight"].cuda(), ag.qnet_params["head.weight"][0])

    fresh = make_agent(case_inputs(CASES[name]))
    with torch.no_grad():
        fresh.arena.flat.mul_(0.5)  # different weights, wrong on purpose
    fresh.load_from_disk(path)
    torch.cuda.synchronize()
    assert fresh.timesteps_so_far == 1234
    for r in range(4):  # online, target, exp_avg, exp_avg_sq (gradients are scratch)
        assert torch.equal(fresh.arena.flat[0, r], ag.arena.flat[0, r]), f"region {r}"
    assert fresh.counters[:3].tolist() == ag.counters[:3].tolist()
    if name.startswith("sac"):
        for slot in (0, 2, 3):  # log_alpha, exp_avg, exp_avg_sq (slot 1 is the gradient: scratch)
            assert torch.equal(fresh._alpha_state[slot], ag._alpha_state[slot])
    # resume: both continue bit-identically (host-side cadence counters are the trainer's to restore)
    fresh.qnet_updates_so_far, fresh.actor_updates_so_far = ag.qnet_updates_so_far, ag.actor_updates_so_far
    _train(ag, inp, range(4, 6))
    _train(fresh, inp, range(4, 6))
    assert torch.equal(fresh.arena.flat[0, :4], ag.arena.flat[0, :4])


def test_load_is_not_a_network_call():
    inp = case_inputs(CASES["td3_hopper"])
    with pytest.raises(NotImplementedError):
        make_agent(inp).load("entity/project/run")
