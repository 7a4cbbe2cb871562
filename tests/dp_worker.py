"""torchrun worker for tests/test_gpu_dp.py: W data-parallel ranks vs one rank on the concatenated batch."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200 import _lib as L, sac_hps, td3_hps  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.dp import DataParallelLearner, GradComm  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.replay import pack_rows  # noqa: E402
from tests.helpers import rel_dev  # noqa: E402


def make_agent(hps, agent_id, dev):
    torch.manual_seed(0)  # identical initial parameters on every rank
    return Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32),
                 dev, hps, seed=3, agent_id=agent_id)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # B2RL_DP_BACKEND=gloo: all ranks share GPU 0 and reduce through gloo (CUDA tensors staged through the host) — the same
    # kernels, grad_scale = 1/W and deferred alpha step as under NCCL, runnable on a one-GPU box (NCCL refuses two ranks
    # on one device)
    backend = os.environ.get("B2RL_DP_BACKEND", "nccl")
    local = local % torch.cuda.device_count() if backend == "gloo" else local
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    B, n_it = 64, 4
    # row-group kernels (fp32, tight bound) and the tensor-core wide path (3xTF32: fp32-class gradients, but a ReLU mask
    # may flip where a pre-activation is within 1e-6 of zero, and Adam's first steps amplify: looser stated bound)
    for algo, wide, tol in (("sac", None, 5e-5), ("td3", None, 5e-5), ("sac", "3xtf32", 1e-3), ("td3", "3xtf32", 1e-3)):
        mk = sac_hps if algo == "sac" else td3_hps
        ag = make_agent(mk(batch_size=B), rank, dev)
        dp = DataParallelLearner(ag, None, B, GradComm(), wide=wide)
        ref = make_agent(mk(batch_size=B * world), 0, dev) if rank == 0 else None
        delay = 2
        for i in range(n_it):
            td = make_synthetic_transitions(B * world, 11, 3, [-1.0] * 3, [1.0] * 3, seed=500 + i)
            rows_all = pack_rows({k: v.to(dev) for k, v in td.items()}, ag.fmt)
            g = torch.Generator().manual_seed(900 + i)
            eq = torch.randn(B * world, 3, generator=g).to(dev)
            ep = [torch.randn(B * world, 3, generator=g).to(dev) for _ in range(delay)]
            ea = [torch.randn(B * world, 3, generator=g).to(dev) for _ in range(delay)]
            sl = slice(rank * B, (rank + 1) * B)
            dp.iteration(i, rows=rows_all[sl].contiguous(), eps_q=eq[sl].contiguous(),
                         eps_pi=[e[sl].contiguous() for e in ep], eps_alpha=[e[sl].contiguous() for e in ea])
            if ref is not None:
                ref.update_qnets(rows_all, eps=eq)
                ref.qnet_updates_so_far += 1
                if i % (delay + 1) == 0:
                    for j in range(delay):
                        ref.update_actor(rows_all, eps=ep[j], eps_alpha=ea[j])
                ref.update_targ_nets()
        torch.cuda.synchronize()
        # replicas are bit-identical
        mine = ag.arena.flat[0, :4].double().sum().reshape(1)
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert all(torch.equal(g, got[0]) for g in got), f"{algo}: replicas diverged {got}"
        if rank == 0:
            worst = 0.0
            for r in (L.REGION_P, L.REGION_T):
                for net in (*ag.layout.critic, ag.layout.actor):
                    for name, t in ag.arena.named(net, r).items():
                        worst = max(worst, rel_dev(t, ref.arena.named(net, r)[name]))
            if algo == "sac":
                worst = max(worst, rel_dev(ag.log_alpha, ref.log_alpha))
            print(f"DP_RESULT {algo} wide={wide} world={world} worst_rel_dev_vs_single_rank={worst:.3e}", flush=True)
            assert worst <= tol, worst
            assert ag.counters[:3].tolist() == ref.counters[:3].tolist()
    if backend == "nccl":
        graph_replay_equals_eager(rank, world, dev)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DP_OK", flush=True)


def graph_replay_equals_eager(rank, world, dev):
    """W ranks, device-side sampling: the iteration graphs (NCCL all-reduces captured inside) replay the same launches as
    the eager loop — parameters, optimizer state and counters bitwise equal after 9 iterations, replicas identical."""
    from sac_td3_cudagraphs_pytorch_b200.replay import ReplayBuffer
    for algo, wide in (("sac", "3xtf32"), ("td3", None)):
        mk = sac_hps if algo == "sac" else td3_hps
        agents, dps = [], []
        for graphs in (True, False):
            ag = make_agent(mk(batch_size=256), rank, dev)
            rb = ReplayBuffer(2000, dev, seed=9, agent_id=rank)
            td = make_synthetic_transitions(1500, 11, 3, [-1.0] * 3, [1.0] * 3, seed=40 + rank)
            rb.extend({k: v.to(dev) for k, v in td.items()})
            agents.append(ag)
            dps.append(DataParallelLearner(ag, rb, 256, GradComm(), wide=wide, graphs=graphs))
        assert dps[0].graphs and not dps[1].graphs
        for i in range(9):
            for dp in dps:
                dp.iteration(i)
        torch.cuda.synchronize()
        assert len(dps[0]._graphs) >= 2
        assert torch.equal(agents[0].arena.flat[:, :4], agents[1].arena.flat[:, :4]), f"{algo}: graph != eager"
        assert torch.equal(agents[0].counters[:4], agents[1].counters[:4])
        assert torch.equal(agents[0]._alpha_state[[0, 2, 3]], agents[1]._alpha_state[[0, 2, 3]])
        mine = agents[0].arena.flat[0, :4].double().sum().reshape(1)
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert all(torch.equal(g, got[0]) for g in got), f"{algo}: replicas diverged {got}"
        if rank == 0:
            print(f"DP_GRAPH_OK {algo} wide={wide} world={world} graphs={len(dps[0]._graphs)}", flush=True)


if __name__ == "__main__":
    main()
