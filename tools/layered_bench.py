"""Q-evaluation of ONE large state graph (BASELINE.json config 4: ~1M-triangle mesh, all vertices in the state):
whole-forward time per GEMM mode, and the message-passing / node-GEMM kernels alone against their rooflines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from meshdqn_b200 import _lib  # noqa: E402
from meshdqn_b200.airfoilgcnn import NodeRemovalNet  # noqa: E402
from meshdqn_b200.data import Data  # noqa: E402
from meshdqn_b200.synthetic import field_values, synthetic_airfoil_mesh  # noqa: E402


def state_graph(n_tri, seed=0, T=5):
    coords, cells, _ = synthetic_airfoil_mesh(n_tri, seed=seed, order="morton")
    u, p = field_values(coords, T, seed)
    x = np.concatenate([coords, u.transpose(1, 0, 2).reshape(len(coords), -1), p.T], axis=1).astype(np.float32)
    c = cells.astype(np.int64)
    ei = np.stack([np.stack([c[:, 0], c[:, 0], c[:, 1]], 1).ravel(), np.stack([c[:, 1], c[:, 2], c[:, 2]], 1).ravel()])
    return Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei))


def timed(fn, iters, flush_buf):
    ts = []
    for _ in range(iters):
        if flush_buf is not None:
            mode = os.environ.get("FLUSH", "write")
            if mode == "write":
                flush_buf.fill_(1)          # leaves L2 full of DIRTY lines: the timed kernel also pays their write-back
            elif mode == "read":
                flush_buf.view(torch.int64).sum()   # L2 full of clean lines

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ntri = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda:0")
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, bf16 = pk.get("hbm_gbs", 6650.0), pk.get("bf16_tflops", 1590.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t0 = time.time()
    d = state_graph(ntri).to(dev)
    N, E = int(d.x.shape[0]), int(d.edge_index.shape[1])
    print(f"state graph: {N} nodes, {E} directed edges, F=17 (mesh {ntri} triangles), setup {time.time()-t0:.1f}s", flush=True)
    from conftest import lively_state_dict
    from oracle import gnn_ref
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(ref))
    net = net.to(dev)
    out = {"nodes": N, "edges": E}
    with torch.no_grad():
        res = {}
        for gemm in ("fp32", "tf32x3"):
            net.layered_gemm = gemm
            am, q = net.select_action(d)
            torch.cuda.synchronize()
            res[gemm] = q.clone()
            for _ in range(2):
                net.select_action(d)
            ms_eager = timed(lambda: net.select_action(d), 5, flush)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                net.select_action(d)
            ms = timed(g.replay, 10, flush)
            print(f"  Q-eval [{gemm:6s}] eager {ms_eager:7.3f} ms  graph {ms:7.3f} ms  {N/ms/1e3:8.1f} M nodes/s  argmax {int(am[0])}", flush=True)
            out[f"q_eval_{gemm}_ms"] = ms
        print("  tf32x3 vs fp32 Q rel diff:", float(((res['tf32x3'] - res['fp32']).abs() / res['fp32'].abs().clamp_min(1e-30)).max()))
    # ---- kernels alone ----
    L, p = _lib.lib(), _lib.ptr
    st = _lib.stream_ptr
    src = d.edge_index[0].to(torch.int32).contiguous()
    dst = d.edge_index[1].to(torch.int32).contiguous()
    ecount = torch.tensor([E], dtype=torch.int32, device=dev)
    row_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(L.mdq_csr_build_scratch_words(E, N)), dtype=torch.int32, device=dev)
    csr = lambda: _lib.check(L.mdq_csr_build(p(src), p(dst), p(ecount), E, N, p(row_ptr), p(col), p(scratch), st()))
    csr()
    ms = timed(csr, 10, flush)
    print(f"  csr_build            {ms*1e3:8.1f} us")
    out["csr_build_us"] = ms * 1e3
    for F, nrows, label in ((17, N, "block 0 (17 features, all nodes)"), (128, N // 10, "block 1 shape (128 features, N/10 nodes)")):
        if F == 17:
            x, rp, cl = d.x, row_ptr, col
            lda = 40
        else:  # a sub-graph with the pooled level's size: first nrows nodes, edges among them
            m = (d.edge_index[0] < nrows) & (d.edge_index[1] < nrows)
            s2, d2 = src[m].contiguous(), dst[m].contiguous()
            e2 = int(s2.numel())
            ec2 = torch.tensor([e2], dtype=torch.int32, device=dev)
            rp = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
            cl = torch.empty(max(e2, 1), dtype=torch.int32, device=dev)
            sc2 = torch.empty(int(L.mdq_csr_build_scratch_words(e2, nrows)), dtype=torch.int32, device=dev)
            _lib.check(L.mdq_csr_build(p(s2), p(d2), p(ec2), e2, nrows, p(rp), p(cl), p(sc2), st()))
            x = torch.randn(nrows, 128, device=dev)
            lda = 256
        e_rows = int(rp[nrows].item())
        A = torch.empty(nrows, lda, device=dev)
        agg = lambda: _lib.check(L.mdq_sage_aggregate(p(x), F, 0, F, p(rp), p(cl), nrows, p(A), lda, st()))
        agg()
        ms = timed(agg, 10, flush)
        alg = 4 * (2 * nrows * F + e_rows + nrows + 1)             # x read once, aggregate written once, indices
        moved = 4 * (nrows * F + nrows * lda + e_rows + nrows + 1)  # what this kernel must move: it also writes the [x | pad] half
        print(f"  sage_aggregate {label}: {ms*1e3:8.1f} us  algorithmic {alg/1e6:.1f} MB -> {alg/ms/1e6:7.1f} GB/s = {alg/ms/1e6/hbm*100:5.1f}% of {hbm:.0f}"
              f"  (incl. the [x|pad] copy it also writes: {moved/ms/1e6:7.1f} GB/s = {moved/ms/1e6/hbm*100:5.1f}%)", flush=True)
        out[f"sage_F{F}"] = {"us": ms * 1e3, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6, "frac": alg / ms / 1e6 / hbm,
                             "moved_bytes": moved, "moved_frac": moved / ms / 1e6 / hbm}
        K = 2 * F
        W = torch.randn(K, 128, device=dev) / K ** 0.5
        kpad = (K + 7) // 8 * 8
        wp = torch.zeros(kpad, 128, device=dev)
        wp[:K] = W
        hi = (wp.view(torch.int32) & -8192).view(torch.float32)
        tile = lambda m: m.view(kpad // 4, 4, 16, 8).permute(0, 2, 3, 1).contiguous().view(-1)
        ws = torch.cat([tile(hi), tile(wp - hi)])
        bias, pool = torch.randn(128, device=dev), torch.randn(128, device=dev)
        score = torch.empty(nrows, device=dev)
        C = torch.empty(nrows, 128, device=dev)
        for mode, name in ((0, "fp32 FFMA"), (1, "tcgen05 3xTF32")):
            for store in (False, True):
                f = lambda: _lib.check(L.mdq_node_gemm(p(A), None, lda, K, nrows, 128, p(W), p(ws), p(bias), p(pool), None, 1, mode,
                                                       p(C) if store else None, p(score), st()))
                f()
                ms = timed(f, 10, flush)
                fl = 2.0 * nrows * K * 128
                byt = 4 * (nrows * lda + nrows + (nrows * 128 if store else 0))
                print(f"  node_gemm K={K:3d} M={nrows} [{name:14s}] {'store C+score' if store else 'score only   '}: {ms*1e3:8.1f} us  "
                      f"{fl/ms/1e9:7.1f} TFLOP/s (algorithmic 2MKN; tensor pipe does 3x)  bytes {byt/ms/1e6:7.1f} GB/s = {byt/ms/1e6/hbm*100:5.1f}% HBM", flush=True)
                out[f"gemm_F{F}_{mode}_{int(store)}"] = {"us": ms * 1e3, "tflops": fl / ms / 1e9, "GBps": byt / ms / 1e6}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"layered_bench_{ntri}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
