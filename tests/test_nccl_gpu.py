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
    log = open(out_path + f".rank{rank}.log", "w")
    for name, fused_ar, graphs in (("nccl", False, False), ("fused", True, False), ("fused_graph", True, True)):
        n2 = nets()
        tr = ReplayTrainer(n2[0], n2[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2, graphs=graphs,
                           fused_allreduce=fused_ar)
        assert tr.world == world
        if fused_ar:
            assert tr.fused_allreduce, "symmetric-memory mapping failed: the fused all-reduce + Adam kernel did not run"
        per = len(trans) // world
        rb = ReplayBatch.from_transitions(trans[rank * per:(rank + 1) * per]).pin_memory(slim=True).to(dev).mark_static()
        losses = []
        for k in range(6):
            loss = tr.step(rb).clone()
            tr.flush()
            torch.cuda.synchronize()
            print(name, "step", k, "done", file=log, flush=True)
            dist.all_reduce(loss)
            losses.append(float(loss) / world)
        tr.flush()
        torch.cuda.synchronize()
        res[name] = (losses, n2[0]._flat.clone(), n2[1]._flat.clone())
        tr._graphs.clear()
        dist.barrier()
    if rank == 0:
        n1 = nets()
        # single-rank reference: no collective set-up (rank 1 is not taking part), world forced to 1
        tr1 = ReplayTrainer(n1[0], n1[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2, fused_allreduce=False)
        tr1.world = 1
        rb = ReplayBatch.from_transitions(trans).to(dev)
        l1 = [float(tr1.step(rb)) for _ in range(6)]
        torch.cuda.synchronize()
        torch.save({"l1": l1, "w1": [n1[0]._flat.cpu(), n1[1]._flat.cpu()],
                    **{k: (v[0], [v[1].cpu(), v[2].cpu()]) for k, v in res.items()}}, out_path)
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
        p.join(timeout=100)
    for p in procs:
        if p.is_alive():
            p.kill()
            logs = "".join(open(out + f".rank{r}.log").read() for r in range(2) if os.path.exists(out + f".rank{r}.log"))
            pytest.fail("rank did not finish; progress:\n" + logs)
        assert p.exitcode == 0
    z = torch.load(out)
    for name in ("nccl", "fused", "fused_graph"):
        losses, w = z[name]
        for a, b in zip(z["l1"], losses):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (name, z["l1"], losses)
        for i in range(2):
            d = (z["w1"][i] - w[i]).abs().max().item()
            assert d < 5e-4, (name, d)   # Adam at lr 1e-3: another fp32 summation order of a near-zero gradient can flip a step's sign
    # two ranks: a + b in either order is the same float, so the peer-memory kernel equals NCCL + Adam bit for bit,
    # and graph replay changes nothing
    assert z["fused"][0] == z["nccl"][0] == z["fused_graph"][0]
    for i in range(2):
        assert torch.equal(z["fused"][1][i], z["nccl"][1][i]) and torch.equal(z["fused"][1][i], z["fused_graph"][1][i])


def _actor_worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import contextlib
    import io
    import torch.distributed as dist
    from conftest import make_config, oracle_fields
    from meshdqn_b200 import dqn
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    mk = lambda: Env2DAirfoil(cfg, mesh=(coords, cells), device=dev)
    torch.manual_seed(1370)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n.to(dev))
    w0 = nets[0]._flat.clone() if getattr(nets[0], "_flat", None) is not None else None
    with contextlib.redirect_stdout(io.StringIO()):
        h = dqn.train_replicas(mk, nets[0], nets[1], n_envs=2, rounds=5, batch_size=4, eps_decay=8.0, target_update=3,
                               memory_capacity=64, device=dev, save_prefix=out_path + "_", lr=1e-3)
    torch.cuda.synchronize()
    sd = {k: v.cpu() for k, v in nets[0].state_dict().items()}
    torch.save({"sd": sd, "actions": h.actions, "losses": h.losses, "epss": h.epss}, out_path + f".rank{rank}.pt")
    dist.barrier()
    os._exit(0)


def test_actor_loop_two_ranks_keep_identical_nets(tmp_path):
    """dqn.train_replicas under a 2-rank NCCL group: each rank explores with its own replicas and replay memory, the
    gradients are averaged every optimisation step, so both ranks end with bit-identical policy nets."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "actor")
    port = 29800 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_actor_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=150)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("actor-loop rank did not finish")
        assert p.exitcode == 0
    a, b = torch.load(out + ".rank0.pt"), torch.load(out + ".rank1.pt")
    assert all(torch.equal(a["sd"][k], b["sd"][k]) for k in a["sd"])
    assert len(a["losses"]) == len(b["losses"]) == 10 - 2 and a["actions"] != b["actions"]   # different exploration per rank
    assert os.path.exists(out + "_policy_net_1.pt") and os.path.exists(out + "_rank1_eps.npy")
