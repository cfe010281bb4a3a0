"""GPU harness for the staged tensor-core Q-path: stage-by-stage comparison with the CPU oracle (float32 and float64
evaluations), the fused kernel beside it, gradients, and per-path timings (CUDA events, L2 flushed).

    python tools/staged_check.py [--quick]

Test infrastructure: imports oracle/ as the checker only.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import lively_state_dict  # noqa: E402
from meshdqn_b200.airfoilgcnn import NodeRemovalNet  # noqa: E402
from meshdqn_b200.data import Batch, Data  # noqa: E402
from oracle import gnn_ref  # noqa: E402

dev = torch.device("cuda:0")


def rand_graph(g, n=180, e=369, f=17):
    return Data(x=torch.randn(n, f, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))


def make(lively):
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    if lively:
        ref.load_state_dict(lively_state_dict(ref))
    nets = []
    for path in ("staged", "fused"):
        net = NodeRemovalNet(181, 128, 0.1)
        net.set_num_nodes(17)
        net.load_state_dict(ref.state_dict())
        net = net.to(dev)
        net.qpath = path
        nets.append(net)
    ref64 = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref64.set_num_nodes(17)
    ref64.load_state_dict(ref.state_dict())
    ref64 = ref64.double()
    return nets[0], nets[1], ref, ref64


def oracle_levels(ref, b):
    """Per-level kept rows (x after pooling), readouts and perms of the oracle."""
    x, ei, batch, ng = gnn_ref._unpack(b)
    out = []
    for conv, pool in ((ref.conv1, ref.pool1), (ref.conv2, ref.pool2), (ref.conv4, ref.pool4), (ref.conv5, ref.pool5)):
        h = F.relu(conv(x, ei))
        x, ei, batch, perm, sc = pool(h, ei, batch, ng)
        r = torch.cat([gnn_ref.global_max_pool(x, batch, ng), gnn_ref.global_mean_pool(x, batch, ng)], dim=1)
        out.append(dict(x=x, ei=ei, batch=batch, perm=perm, score=sc, r=r))
    return out


def f64(ref64, b):
    b64 = Batch(x=b.x.double(), edge_index=b.edge_index)
    b64.batch = getattr(b, "batch", None)
    if hasattr(b, "num_graphs"):
        b64.num_graphs = b.num_graphs

    class D:  # _unpack casts x to float: bypass it
        pass
    x, ei, batch = b.x.double(), b.edge_index, (b.batch if getattr(b, "batch", None) is not None else torch.zeros(b.x.shape[0], dtype=torch.long))
    ng = int(batch.max()) + 1
    acc = None
    for conv, pool in ((ref64.conv1, ref64.pool1), (ref64.conv2, ref64.pool2), (ref64.conv4, ref64.pool4), (ref64.conv5, ref64.pool5)):
        h = F.relu(conv(x, ei))
        x, ei, batch, _, _ = pool(h, ei, batch, ng)
        r = torch.cat([gnn_ref.global_max_pool(x, batch, ng), gnn_ref.global_mean_pool(x, batch, ng)], dim=1)
        acc = r if acc is None else acc + r
    y = F.relu(ref64.lin1(acc))
    y = F.relu(ref64.lin2(y))
    return F.softmax(ref64.lin3(y), dim=1), acc


def stage_dump(net, b, B, max_n):
    """Level-1/2 tensors the staged forward left in its workspace (forward-only layout, gnn_staged_api.cuh:stg_carve)."""
    ws = [v for k, v in net._stg_wss.items() if not k[1]][0]
    R1 = int(np.ceil(np.float32(0.1) * np.float32(max_n)))
    R2 = int(np.ceil(np.float32(0.1) * np.float32(R1)))
    o = 0

    def take(n):
        nonlocal o
        v = ws[o:o + n]
        o += (n + 3) // 4 * 4
        return v
    x1 = take(B * R1 * 128).view(B, R1, 128).cpu()
    r0 = take(B * 256).view(B, 256).cpu()
    r1 = take(B * 256).view(B, 256).cpu()
    return x1, r0, r1, R1, R2


def check_forward(name, graphs, lively):
    st, fu, ref, ref64 = make(lively)
    b = Batch.from_data_list(graphs) if len(graphs) > 1 else graphs[0]
    with torch.no_grad():
        q_ref = ref(b)
        e_ref = ref(b, embedding=True)
        q64, e64 = f64(ref64, b)
        bd = b.to(dev)
        q_fu = fu(bd).cpu()
        q_st = st(bd).cpu()
        e_st = st(bd, embedding=True).cpu()
        am, _ = st.select_action(bd)
    B = q_ref.shape[0]
    lv = oracle_levels(ref, b)
    max_n = max(int(g.x.shape[0]) for g in graphs)
    x1, r0, r1, R1, R2 = stage_dump(st, b, B, max_n)
    # level-0 readout and kept rows against the oracle
    d_r0 = (r0 - lv[0]["r"]).abs().max().item()
    d_r1 = (r1 - lv[1]["r"]).abs().max().item()
    x1o = lv[0]["x"]
    cnt = torch.bincount(lv[0]["batch"], minlength=B)
    off = 0
    d_x1 = 0.0
    for g in range(B):
        k = int(cnt[g])
        d_x1 = max(d_x1, (x1[g, :k] - x1o[off:off + k]).abs().max().item())
        off += k

    def rel(a, bb):
        return ((a - bb).abs() / bb.abs().clamp_min(1e-30)).max().item()
    print(f"[{name}] B={B} lively={lively}")
    print(f"   stage0: |x1 - oracle| {d_x1:.3e}  |r0 - oracle| {d_r0:.3e}   stage1: |r1 - oracle| {d_r1:.3e}")
    print(f"   embedding: staged vs oracle32 {(e_st - e_ref).abs().max().item():.3e} (scale {e_ref.abs().max().item():.3e}); "
          f"oracle32 vs f64 {(e_ref.double() - e64).abs().max().item():.3e}; staged vs f64 {(e_st.double() - e64).abs().max().item():.3e}")
    print(f"   Q rel: staged/oracle32 {rel(q_st, q_ref):.3e}  fused/oracle32 {rel(q_fu, q_ref):.3e}  "
          f"oracle32/f64 {rel(q_ref.double(), q64):.3e}  staged/f64 {rel(q_st.double(), q64):.3e}  fused/f64 {rel(q_fu.double(), q64):.3e}")
    print(f"   actions: staged==oracle {int((q_st.argmax(1) == q_ref.argmax(1)).sum())}/{B}  fused==oracle "
          f"{int((q_fu.argmax(1) == q_ref.argmax(1)).sum())}/{B}  select_action==argmax {bool((am.cpu().long() == q_st.argmax(1)).all())}")
    return rel(q_st, q_ref)


def check_backward(graphs, lively):
    st, fu, ref, _ = make(lively)
    b = Batch.from_data_list(graphs)
    g = torch.Generator().manual_seed(11)
    w = torch.randn(len(graphs), 181, generator=g)
    ref.zero_grad()
    (ref(b) * w).sum().backward()
    res = {}
    for name, net in (("staged", st), ("fused", fu)):
        net.zero_grad()
        q = net(b.to(dev))
        (q * w.to(dev)).sum().backward()
        worst = 0.0
        detail = []
        for (k, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
            if pr.grad is None:
                assert p.grad is None, k
                continue
            sc = pr.grad.abs().max().item()
            d = (p.grad.cpu() - pr.grad).abs().max().item()
            detail.append((k, d / (sc + 1e-30), sc))
            worst = max(worst, d / (sc + 1e-30))
        res[name] = (worst, detail)
    print(f"[backward] B={len(graphs)} lively={lively}: worst |grad - oracle| / scale: staged {res['staged'][0]:.3e}  fused {res['fused'][0]:.3e}")
    for k, d, sc in res["staged"][1]:
        flag = "  <<<" if d > 1e-3 else ""
        print(f"      {k:24s} rel {d:.3e} scale {sc:.3e}{flag}")


def check_replay(lively):
    from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
    g = torch.Generator().manual_seed(5)
    trans = []
    for i in range(64):
        s = rand_graph(g)
        nx = None if i % 7 == 0 else rand_graph(g)
        trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
    rb = ReplayBatch.from_transitions(trans)
    out = {}
    for path in ("staged", "fused"):
        nets = []
        for seed in (1370, 1371):
            torch.manual_seed(seed)
            ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
            ref.set_num_nodes(17)
            if lively:
                ref.load_state_dict(lively_state_dict(ref, seed=seed))
            net = NodeRemovalNet(181, 128, 0.1)
            net.set_num_nodes(17)
            net.load_state_dict(ref.state_dict())
            net = net.to(dev)
            net.qpath = path
            nets.append(net)
        tr = ReplayTrainer(nets[0], nets[1], lr=1e-3, weight_decay=1e-6, gamma=1.0, target_update=2)
        losses = [float(tr.step(rb.to(dev))) for _ in range(4)]
        out[path] = (losses, nets[0]._flat.clone(), nets[1]._flat.clone())
    print(f"[replay] losses staged {out['staged'][0]}\n         losses fused  {out['fused'][0]}")
    for i in (1, 2):
        d = (out["staged"][i] - out["fused"][i]).abs().max().item()
        print(f"         |weights net{i}: staged - fused| after 4 steps {d:.3e}")


def timeit(fn, iters=30, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def timings():
    from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
    st, fu, ref, _ = make(True)
    g = torch.Generator().manual_seed(9)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B in (1, 256, 1024):
        b = Batch.from_data_list([rand_graph(g) for _ in range(B)]).to(dev)
        with torch.no_grad():
            t_st = timeit(lambda: st.select_action(b), flush=flush)
            t_fu = timeit(lambda: fu.select_action(b), flush=flush)
        print(f"[time] forward B={B}: staged {t_st:.1f} us   fused {t_fu:.1f} us")
    trans = []
    for i in range(256):
        s = rand_graph(g)
        nx = None if i % 9 == 0 else rand_graph(g)
        trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
    rb = ReplayBatch.from_transitions(trans).to(dev)
    for path in ("staged", "fused"):
        nets = []
        for seed in (1, 2):
            net = NodeRemovalNet(181, 128, 0.1)
            net.set_num_nodes(17)
            net = net.to(dev)
            net.qpath = path
            nets.append(net)
        tr = ReplayTrainer(nets[0], nets[1])
        t = timeit(lambda: tr.step(rb), flush=flush)
        tr.timers = {}
        for _ in range(10):
            flush.fill_(1)
            tr.step(rb)
        torch.cuda.synchronize()
        parts = {k: float(np.median([a.elapsed_time(bb) * 1e3 for a, bb in v])) for k, v in tr.timers.items()}
        print(f"[time] replay step B=256 {path}: {t:.1f} us   parts {parts}")


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    g = torch.Generator().manual_seed(3)
    t0 = time.time()
    check_forward("single", [rand_graph(g)], False)
    check_forward("single", [rand_graph(g)], True)
    check_forward("batch32", [rand_graph(g) for _ in range(32)], True)
    ragged = [rand_graph(g, n=int(torch.randint(1, 181, (1,), generator=g)), e=int(torch.randint(0, 500, (1,), generator=g)))
              for _ in range(40)]
    ragged.append(Data(x=torch.randn(1, 17, generator=g), edge_index=torch.zeros(2, 0, dtype=torch.long)))
    ragged.append(Data(x=torch.randn(9, 17, generator=g), edge_index=torch.tensor([[0, 0, 0, 3], [1, 1, 1, 3]])))
    ragged.append(Data(x=torch.zeros(180, 17), edge_index=torch.randint(0, 180, (2, 369), generator=g)))
    ragged.append(rand_graph(g, n=256, e=700))
    check_forward("ragged", ragged, False)
    check_forward("ragged", ragged, True)
    check_backward([rand_graph(g) for _ in range(16)], True)
    check_backward(ragged, False)
    check_replay(True)
    if not quick:
        timings()
    print(f"done in {time.time() - t0:.1f} s")
