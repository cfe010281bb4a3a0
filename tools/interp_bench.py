"""Timing of the re-interpolation kernels on synthetic meshes (CUDA events, L2 flushed between launches)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200.Env2DAirfoil import SourceField  # noqa: E402
from meshdqn_b200.flow_solver import DeviceMesh  # noqa: E402
from meshdqn_b200.synthetic import synthetic_airfoil_mesh, synthetic_fields  # noqa: E402


def main():
    ntri = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
    orders = sys.argv[2].split(",") if len(sys.argv) > 2 else ["random", "morton"]
    dev = torch.device("cuda:0")
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    from scipy.spatial import Delaunay
    for order in orders:
        t0 = time.time()
        coords, cells, _ = synthetic_airfoil_mesh(ntri, seed=0, order=order)
        m0 = DeviceMesh(coords, cells, dev)
        U0, P0 = synthetic_fields(coords, m0.edges.cpu().numpy(), 5, 0)
        isb = m0.on_boundary.cpu().numpy().astype(bool)
        rng = np.random.RandomState(1)
        keep = np.ones(len(coords), bool)
        keep[rng.choice(np.nonzero(~isb)[0], max(1, len(coords) // 100), replace=False)] = False
        c2 = coords[keep]
        t2 = Delaunay(c2).simplices
        t2 = t2[isb[keep][t2].sum(1) != 3]
        m2 = DeviceMesh(c2, t2, dev)
        npt = m2.nv + m2.ne
        T = 5
        alg = 16 * m2.nv + 8 * m2.ne + 16 * m0.nv + 36 * m0.nc + 8 * T * (2 * (m0.nv + m0.ne) + m0.nv) + 8 * T * (2 * npt + m2.nv) + 4 * npt
        print(f"[{order}] mesh {m0.nc} cells, {npt} target points, algorithmic {alg/1e6:.1f} MB, setup {time.time()-t0:.1f}s", flush=True)
        res = {}
        leafs = [int(x) for x in os.environ.get("LEAF_CELLS", "256,128").split(",")]
        for name, kw in [("grid", dict(tiled=False))] + [(f"tiled{k}", dict(tiled=True, leaf_cells=k)) for k in leafs]:
            t0 = time.time()
            if kw["tiled"] and os.environ.get("MICRO_BINS"):
                kw = dict(kw, micro_bins_per_cell=float(os.environ["MICRO_BINS"]))
            src = SourceField(m0, U0, P0, **kw)
            tb = time.time() - t0
            out = src.interpolate(m2)
            torch.cuda.synchronize()
            res[name] = out
            for _ in range(3):
                src.interpolate(m2)
            def timed(fn):
                ts = []
                for _ in range(10):
                    mode = os.environ.get("FLUSH", "write")
                    if mode == "write":
                        flush_buf.fill_(1)
                    elif mode == "read":
                        flush_buf.view(torch.int64).sum()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                return float(np.median(ts))
            ms_eager = timed(lambda: src.interpolate(m2))
            # CUDA graph of the same call: no host launch latency between the kernels
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gout = src.interpolate(m2)
            ms = timed(g.replay)
            assert all(torch.equal(a, b) for a, b in zip(gout[:3], out[:3]))
            print(f"  {name:9s} eager {ms_eager*1e3:8.1f} us", flush=True)
            extra = f" index {src.tile_bytes/1e6:.1f} MB smem {src.tile_smem} leaves {src.tile_host.n_leaves}" if kw["tiled"] else ""
            print(f"  {name:9s} build {tb:.2f}s  {ms*1e3:8.1f} us  {alg/ms/1e6:8.1f} GB/s  {alg/ms/1e6/hbm*100:5.1f}% of {hbm:.0f}  "
                  f"{npt/ms/1e6:.2f} Gpts/s  miss {int(out[3])}{extra}", flush=True)
        for k in res:
            if k != "grid":
                print(f"  {k} == grid bit for bit:", all(torch.equal(a, b) for a, b in zip(res["grid"][:3], res[k][:3])), flush=True)


if __name__ == "__main__":
    main()
