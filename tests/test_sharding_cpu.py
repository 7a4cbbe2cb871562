"""CPU tests of the multi-GPU host logic: agent-id sharding and shard-independent initialisation,
including a world_size-2 gloo run (one process per rank, as bench.py is launched)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sac_td3_cudagraphs_pytorch_b200.population import init_agent_params, shard


@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 8), (10, 4), (3, 8), (7, 2)])
def test_shard_is_a_partition(n, world):
    parts = [shard(n, world, r) for r in range(world)]
    ids = [i for p in parts for i in p]
    assert ids == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1


def _digest(agent_id):
    lo, hi = torch.full((3,), -1.0), torch.full((3,), 1.0)
    a, q1, q2 = init_agent_params(agent_id, 42, 11, 3, False, True, lo, hi)
    return torch.stack([sum(v.double().sum() for v in d.values()) for d in (a, q1, q2)])


def test_init_depends_only_on_global_agent_id():
    torch.manual_seed(1)
    d5 = _digest(5)
    torch.manual_seed(2)
    _ = _digest(9)
    assert torch.equal(_digest(5), d5)            # independent of global RNG state and of call order
    assert not torch.equal(_digest(6), d5)
    before = torch.get_rng_state()
    _digest(3)
    assert torch.equal(torch.get_rng_state(), before)  # leaves the caller's RNG untouched


def _worker(rank, world, port, n_agents, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard(n_agents, world, rank)
    local = torch.stack([_digest(i) for i in mine]) if len(mine) else torch.zeros(0, 3, dtype=torch.float64)
    sizes = [len(shard(n_agents, world, r)) for r in range(world)]
    pad = torch.zeros(max(sizes), 3, dtype=torch.float64)
    pad[:len(mine)] = local
    bufs = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)  # report-time gather of per-agent scalars: the only traffic a population needs
    if rank == 0:
        torch.save(torch.cat([b[:s] for b, s in zip(bufs, sizes)]), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_population_matches_single_process(tmp_path):
    n_agents = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, n_agents, out), nprocs=2, join=True)
    got = torch.load(out)
    want = torch.stack([_digest(i) for i in range(n_agents)])
    assert torch.equal(got, want)


# ---------------------------------------------------------------- data-parallel gradient bucket (gloo)
def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sac_td3_cudagraphs_pytorch_b200 import _lib as L
    from sac_td3_cudagraphs_pytorch_b200.arena import Arena, make_layout
    from sac_td3_cudagraphs_pytorch_b200.dp import GradComm, reduce_grad_span
    lay = make_layout(11, 3, False, True)
    ar = Arena(lay, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    ar.flat[0, L.REGION_G].copy_(torch.randn(lay.region, generator=g))
    comm = GradComm()
    assert (comm.world, comm.rank) == (world, rank)
    reduce_grad_span(ar, lay, comm, "critic")
    reduce_grad_span(ar, lay, comm, "actor")
    if rank == 0:
        torch.save(ar.flat[0, L.REGION_G].clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_all_reduce(tmp_path):
    """One all-reduce per optimizer step over the contiguous gradient span of the trained nets (W = 2, gloo). The w2n
    shadows' gradient slots are NOT part of the contract any more: the Adam launch steps w2t from its reduced gradient and
    writes the result to both layouts (csrc/adam.cu "shadow pairs"; checked on the GPU by tests/dp_worker.py: replicas
    bit-identical, shadows == primaries transposed), so a slot that lies inside the span is summed along unused and the
    last net's is left alone."""
    from sac_td3_cudagraphs_pytorch_b200.arena import make_layout
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "g.pt")
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    lay = make_layout(11, 3, False, True)
    locals_ = [torch.randn(lay.region, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)]
    want = locals_[0] + locals_[1]
    for net in (*lay.critic, lay.actor):
        assert torch.equal(got[net.begin:net.core_end], want[net.begin:net.core_end])      # summed over ranks
    last = lay.actor  # (its shadow slot lies beyond the reduced span: untouched local values)
    assert torch.equal(got[last.off["w2n"]:last.off["w2n"] + 65536], locals_[0][last.off["w2n"]:last.off["w2n"] + 65536])
