"""GPU parity of the staged tensor-core Q-path (csrc/gnn_staged.cuh, gnn_tail.cuh) through the C-ABI.

conv1 / conv2 run on tcgen05 as 3xTF32 (fp32 operands split hi + lo, fp32 accumulation in TMEM), the tail in fp32 FMA.
Every tolerance is stated against a FLOAT64 evaluation of the oracle network, next to the fp32 oracle's own distance
to that evaluation, so the bound says how much worse than "any fp32 implementation" the kernels are:

  * reference-initialised weights: Q within 1e-5 relative of the fp32 oracle (BASELINE.json's bar) -- measured ~1.6e-6;
  * large-logit weights (|logit| ~ 15): fp32 rounding of the logits alone moves the probabilities by ~1e-4 relative, so
    the bar is |Q - Q64| <= 8 x max|Q32 - Q64| (3xTF32 carries ~2^-21 per product against fp32's 2^-24);
  * chosen actions equal to the oracle's in every case; gradients within 1e-4 of the gradient scale.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import lively_state_dict
from meshdqn_b200.data import Batch, Data
from oracle import gnn_ref

pytestmark = pytest.mark.gpu


def rand_graph(g, n=180, e=369, f=17):
    return Data(x=torch.randn(n, f, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))


def make(dev, lively, path="staged"):
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    if lively:
        ref.load_state_dict(lively_state_dict(ref))
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev)
    net.qpath = path
    return net, ref


def q_float64(ref, b):
    """The oracle network evaluated in float64 (its `_unpack` casts to float32, so the layers are walked here)."""
    r64 = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    r64.set_num_nodes(17)
    r64.load_state_dict(ref.state_dict())
    r64 = r64.double()
    x, ei = b.x.double(), b.edge_index
    batch = b.batch if getattr(b, "batch", None) is not None else torch.zeros(x.shape[0], dtype=torch.long)
    ng = int(batch.max()) + 1
    acc = None
    for conv, pool in ((r64.conv1, r64.pool1), (r64.conv2, r64.pool2), (r64.conv4, r64.pool4), (r64.conv5, r64.pool5)):
        x, ei, batch, _, _ = pool(F.relu(conv(x, ei)), ei, batch, ng)
        r = torch.cat([gnn_ref.global_max_pool(x, batch, ng), gnn_ref.global_mean_pool(x, batch, ng)], dim=1)
        acc = r if acc is None else acc + r
    y = F.relu(r64.lin2(F.relu(r64.lin1(acc))))
    return F.softmax(r64.lin3(y), dim=1)


def ragged_graphs(g):
    gs = [rand_graph(g, n=int(torch.randint(1, 181, (1,), generator=g)), e=int(torch.randint(0, 500, (1,), generator=g)))
          for _ in range(40)]
    gs.append(Data(x=torch.randn(1, 17, generator=g), edge_index=torch.zeros(2, 0, dtype=torch.long)))          # one node, no edges
    gs.append(Data(x=torch.randn(9, 17, generator=g), edge_index=torch.tensor([[0, 0, 0, 3], [1, 1, 1, 3]])))   # duplicates + self loop
    gs.append(Data(x=torch.zeros(180, 17), edge_index=torch.randint(0, 180, (2, 369), generator=g)))            # all scores tie
    gs.append(rand_graph(g, n=256, e=700))                                                                      # two row tiles, the cap
    gs.append(rand_graph(g, n=129, e=300))
    return gs


def rel(a, b):
    return ((a - b).abs() / b.abs().clamp_min(1e-30)).max().item()


def test_staged_path_is_the_default_and_supported(cuda_device):
    net, _ = make(cuda_device, False, path="auto")
    g = torch.Generator().manual_seed(1)
    d = rand_graph(g).to(cuda_device)
    net(d)
    assert net._use_staged(180, 369)
    assert getattr(net, "_stg_w", None) is not None          # the hi / lo tiles were built: the tensor-core path ran
    assert not net._use_staged(300, 369)                     # > 256 nodes: fused / layered kernels


@pytest.mark.parametrize("lively", [False, True])
def test_forward_against_float64(cuda_device, lively):
    net, ref = make(cuda_device, lively)
    g = torch.Generator().manual_seed(4)
    for graphs in ([rand_graph(g)], [rand_graph(g) for _ in range(64)], ragged_graphs(g)):
        b = Batch.from_data_list(graphs) if len(graphs) > 1 else graphs[0]
        with torch.no_grad():
            q32 = ref(b)
            q64 = q_float64(ref, b)
            q = net(b.to(cuda_device)).cpu()
            am, q2 = net.select_action(b.to(cuda_device))
        assert torch.equal(q.argmax(1), q32.argmax(1))
        assert torch.equal(am.cpu().long(), q.argmax(1)) and torch.equal(q2.cpu(), q)
        fp32_own = (q32.double() - q64).abs().max().item()
        err = (q.double() - q64).abs().max().item()
        print(f"lively={lively} B={len(graphs)}: |Q - Q64| {err:.3e}, fp32 oracle's own {fp32_own:.3e}, rel vs oracle32 {rel(q, q32):.3e}")
        if lively:
            assert err <= 8 * fp32_own + 1e-7
        else:
            assert rel(q, q32) < 1e-5


def test_embedding_and_fused_agreement(cuda_device):
    st, ref = make(cuda_device, True)
    fu, _ = make(cuda_device, True, path="fused")
    g = torch.Generator().manual_seed(5)
    b = Batch.from_data_list(ragged_graphs(g))
    with torch.no_grad():
        e_ref = ref(b, embedding=True)
        e_st = st(b.to(cuda_device), embedding=True).cpu()
        e_fu = fu(b.to(cuda_device), embedding=True).cpu()
    scale = e_ref.abs().max()
    assert (e_st - e_ref).abs().max() < 1e-5 * scale
    assert (e_st - e_fu).abs().max() < 1e-5 * scale


@pytest.mark.parametrize("lively", [False, True])
def test_backward_against_oracle(cuda_device, lively):
    net, ref = make(cuda_device, lively)
    g = torch.Generator().manual_seed(6)
    graphs = ragged_graphs(g) if not lively else [rand_graph(g) for _ in range(24)]
    b = Batch.from_data_list(graphs)
    w = torch.randn(len(graphs), 181, generator=g)
    (ref(b) * w).sum().backward()
    (net(b.to(cuda_device)) * w.to(cuda_device)).sum().backward()
    for (k, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        if pr.grad is None:
            assert p.grad is None, k
            continue
        sc = pr.grad.abs().max().item()
        assert (p.grad.cpu() - pr.grad).abs().max().item() <= 1e-4 * sc + 1e-9, k


def test_weight_updates_refresh_the_tiles(cuda_device):
    """derived weight copies must follow load_state_dict / optimizer steps / the Adam kernel."""
    net, ref = make(cuda_device, False)
    g = torch.Generator().manual_seed(7)
    b = Batch.from_data_list([rand_graph(g) for _ in range(8)])
    with torch.no_grad():
        net(b.to(cuda_device))
    ref.load_state_dict(lively_state_dict(ref, seed=11))
    net.load_state_dict(ref.state_dict())
    with torch.no_grad():
        q = net(b.to(cuda_device)).cpu()
        q32, q64 = ref(b), q_float64(ref, b)
    assert torch.equal(q.argmax(1), q32.argmax(1))
    assert (q.double() - q64).abs().max() <= 8 * (q32.double() - q64).abs().max() + 1e-7
    # an in-place torch optimizer step
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    (net(b.to(cuda_device))[:, 3].sum()).backward()
    opt.step()
    ref.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    with torch.no_grad():
        q = net(b.to(cuda_device)).cpu()
        q32, q64 = ref(b), q_float64(ref, b)
    assert (q.double() - q64).abs().max() <= 8 * (q32.double() - q64).abs().max() + 1e-7


def test_replay_steps_staged_equals_fused(cuda_device):
    """Four replay steps (both select branches, terminal transitions, overlap on / off): staged and fused trainers
    stay together to 3xTF32 rounding."""
    from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
    g = torch.Generator().manual_seed(8)
    trans = []
    for i in range(48):
        s = rand_graph(g)
        nx = None if i % 7 == 0 else rand_graph(g)
        trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
    rb = ReplayBatch.from_transitions(trans).to(cuda_device)
    out = {}
    for name, path, overlap in (("staged", "staged", True), ("staged_serial", "staged", False), ("fused", "fused", False),
                                ("staged_early", "staged", True)):
        nets = [make(cuda_device, True, path)[0] for _ in range(2)]
        tr = ReplayTrainer(nets[0], nets[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2)
        tr.overlap = overlap
        tr.early_tail = name == "staged_early"      # the tail launched before Q_other exists, waiting for it on the device
        losses = [float(tr.step(rb)) for _ in range(4)]
        out[name] = (losses, nets[0]._flat.clone(), nets[1]._flat.clone())
    assert out["staged"][0] == out["staged_serial"][0] == out["staged_early"][0]   # neither the side stream nor the early tail changes a bit
    assert torch.equal(out["staged"][1], out["staged_early"][1]) and torch.equal(out["staged"][2], out["staged_early"][2])
    assert torch.equal(out["staged"][1], out["staged_serial"][1]) and torch.equal(out["staged"][2], out["staged_serial"][2])
    for a, b_ in zip(out["staged"][0], out["fused"][0]):
        assert abs(a - b_) <= 1e-5 * max(1.0, abs(b_))
    for i in (1, 2):
        d = (out["staged"][i] - out["fused"][i]).abs().max().item()
        assert d < 5e-4, d          # lr 1e-3 Adam steps: a sign-level change of a tiny gradient moves a weight by 2e-3 at most


def test_staged_gradients_are_deterministic(cuda_device):
    net, _ = make(cuda_device, True)
    g = torch.Generator().manual_seed(9)
    b = Batch.from_data_list([rand_graph(g) for _ in range(32)]).to(cuda_device)
    w = torch.randn(32, 181, generator=g).to(cuda_device)
    grads = []
    for _ in range(2):
        net.zero_grad()
        (net(b) * w).sum().backward()
        grads.append(torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]).clone())
    assert torch.equal(grads[0], grads[1])


def test_graph_replayed_steps_equal_eager_steps(cuda_device):
    """ReplayTrainer(graphs=True): captured steps (both select branches, two alternating static minibatch arenas as the
    e2e loop uses them, 32-bit edges) leave exactly the weights and losses of launch-by-launch steps."""
    from meshdqn_b200.replay import DevicePrefetcher, ReplayBatch, ReplayTrainer
    g = torch.Generator().manual_seed(10)
    trans = []
    for i in range(40):
        s = rand_graph(g)
        nx = None if i % 6 == 0 else rand_graph(g)
        trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
    host = ReplayBatch.from_transitions(trans).pin_memory(slim=True)
    full = ReplayBatch.from_transitions(trans).pin_memory()
    assert host.h2d_bytes() < 0.8 * full.h2d_bytes()
    out = {}
    for mode in ("eager", "graph"):
        nets = [make(cuda_device, True)[0] for _ in range(2)]
        tr = ReplayTrainer(nets[0], nets[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=3, graphs=(mode == "graph"))
        pf = DevicePrefetcher(cuda_device, static=True)
        losses = []
        pf.submit(host)
        for k in range(11):
            rb = pf.take()
            pf.submit(host)
            losses.append(tr.step(rb).clone())
            pf.release()
        torch.cuda.synchronize()
        out[mode] = ([float(x) for x in losses], nets[0]._flat.clone(), nets[1]._flat.clone(), tr.select, tr.num_grads)
    assert out["graph"][3:] == out["eager"][3:]
    assert out["graph"][0] == out["eager"][0]
    assert torch.equal(out["graph"][1], out["eager"][1]) and torch.equal(out["graph"][2], out["eager"][2])
