"""world_size-2 gloo test of the data-parallel host logic (graph sharding + flat-gradient all-reduce)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meshdqn_b200.parallel import allreduce_mean_, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 8192):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import lively_state_dict
    from meshdqn_b200.data import Batch, Data
    from oracle import gnn_ref
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(net))
    g = torch.Generator().manual_seed(5)
    graphs = [Data(x=torch.randn(30, 17, generator=g), edge_index=torch.randint(0, 30, (2, 40), generator=g)) for _ in range(8)]
    w = torch.randn(8, 181, generator=g)
    lo, hi = shard_range(8, rank, world)
    q = net(Batch.from_data_list(graphs[lo:hi]))
    # global-mean loss: each rank contributes (local sum) / (local batch); equal shards -> mean of means
    ((q * w[lo:hi]).sum() / (hi - lo)).backward()
    used = [p for p in net.parameters() if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in used])
    allreduce_mean_(flat, world)
    t = max_over_ranks(float(rank + 1))
    if rank == 0:
        torch.save(dict(flat=flat, tmax=t), out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    from conftest import lively_state_dict
    from meshdqn_b200.data import Batch, Data
    from oracle import gnn_ref
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(net))
    g = torch.Generator().manual_seed(5)
    graphs = [Data(x=torch.randn(30, 17, generator=g), edge_index=torch.randint(0, 30, (2, 40), generator=g)) for _ in range(8)]
    w = torch.randn(8, 181, generator=g)
    ((net(Batch.from_data_list(graphs)) * w).sum() / 8).backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
    assert res["tmax"] == 2.0
    # fp32, different summation order (per-shard vs whole batch): compare against the gradient scale
    assert (res["flat"] - full).abs().max() <= 1e-4 * full.abs().max()


def _cand_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import lively_state_dict
    from meshdqn_b200.data import Batch, Data
    from meshdqn_b200.parallel import evaluate_candidates
    from oracle import gnn_ref
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(net))
    g = torch.Generator().manual_seed(9)
    graphs = [Data(x=torch.randn(24, 17, generator=g), edge_index=torch.randint(0, 24, (2, 30), generator=g)) for _ in range(7)]

    def q_eval(gs):
        with torch.no_grad():
            q = net(Batch.from_data_list(gs))
        return q.argmax(1), q.max(1).values
    act, q = evaluate_candidates(q_eval, graphs, rank, world)
    torch.save(dict(act=act, q=q), out + f".{rank}")
    dist.destroy_process_group()


def test_candidate_evaluation_shards_by_graph_and_gathers(tmp_path):
    """7 candidates over 2 ranks (ragged shards 4 + 3): both ranks end with the full (action, q) table, equal to the
    single-process evaluation."""
    out = str(tmp_path / "cand")
    port = 29800 + os.getpid() % 150
    mp.spawn(_cand_worker, args=(2, port, out), nprocs=2, join=True)
    from conftest import lively_state_dict
    from meshdqn_b200.data import Batch, Data
    from oracle import gnn_ref
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(net))
    g = torch.Generator().manual_seed(9)
    graphs = [Data(x=torch.randn(24, 17, generator=g), edge_index=torch.randint(0, 24, (2, 30), generator=g)) for _ in range(7)]
    with torch.no_grad():
        q = net(Batch.from_data_list(graphs))
    for r in range(2):
        res = torch.load(out + f".{r}")
        assert torch.equal(res["act"], q.argmax(1))
        assert torch.allclose(res["q"], q.max(1).values, rtol=1e-5, atol=0)   # batch-of-4 vs batch-of-7 BLAS blocking
