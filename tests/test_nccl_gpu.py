"""Two-rank NCCL correctness of the data-parallel replay step on the CUDA path.

SCALE runs prove speed; this proves that R ranks x B/R transitions leave the weights and the loss of one rank x B
(/root/reference/airfoil_dqn.py:303-336: Huber `mean` over the global minibatch, one optimizer step on the summed
gradient).  Spawns two processes on two GPUs; skipped on a single-GPU box.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from conftest import lively_state_dict
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    from meshdqn_b200.data import Data
    from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
    from oracle import gnn_ref
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = torch.Generator().manual_seed(21)
    mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 369), generator=g))
    trans = [(mk(), int(torch.randint(0, 181, (1,), generator=g)), None if i % 5 == 0 else mk(), float(torch.randn(1, generator=g)))
             for i in range(32)]

    def nets():
        out = []
        for seed in (7, 8):
            torch.manual_seed(1370)
            ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
            ref.set_num_nodes(17)
            ref.load_state_dict(lively_state_dict(ref, seed=seed))
            n = NodeRemovalNet(181, 128, 0.1)
            n.set_num_nodes(17)
            n.load_state_dict(ref.state_dict())
            out.append(n.to(dev))
        return out
    res = {}
    for graphs in (False, True):
        n2 = nets()
        tr = ReplayTrainer(n2[0], n2[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2, graphs=graphs)
        assert tr.world == world
        per = len(trans) // world
        rb = ReplayBatch.from_transitions(trans[rank * per:(rank + 1) * per]).pin_memory(slim=True).to(dev).mark_static()
        losses = []
        for _ in range(6):
            loss = tr.step(rb).clone()
            dist.all_reduce(loss)
            losses.append(float(loss) / world)
        torch.cuda.synchronize()
        res[graphs] = (losses, n2[0]._flat.clone(), n2[1]._flat.clone())
        tr._graphs.clear()
    if rank == 0:
        n1 = nets()
        tr1 = ReplayTrainer(n1[0], n1[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2)
        tr1.world = 1
        rb = ReplayBatch.from_transitions(trans).to(dev)
        l1 = [float(tr1.step(rb)) for _ in range(6)]
        torch.cuda.synchronize()
        torch.save({"l1": l1, "w1": [n1[0]._flat.cpu(), n1[1]._flat.cpu()],
                    "l2": res[False][0], "w2": [res[False][1].cpu(), res[False][2].cpu()],
                    "l2g": res[True][0], "w2g": [res[True][1].cpu(), res[True][2].cpu()]}, out_path)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)        # no NCCL teardown under captured graphs


def test_two_ranks_equal_one_rank(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    port = 29600 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("rank did not finish")
        assert p.exitcode == 0
    z = torch.load(out)
    for a, b in zip(z["l1"], z["l2"]):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (z["l1"], z["l2"])
    assert z["l2"] == z["l2g"]                                        # graph replay under NCCL changes nothing
    for i in range(2):
        assert torch.equal(z["w2"][i], z["w2g"][i])
        d = (z["w1"][i] - z["w2"][i]).abs().max().item()
        assert d < 5e-4, d    # Adam at lr 1e-3: a different fp32 summation order of a near-zero gradient can flip a step's sign
