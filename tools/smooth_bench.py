"""mdq_mesh_smooth(50) on the ys930 fixture: median CUDA-event time of the single-CTA ordered Gauss-Seidel sweep."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_mesh
from meshdqn_b200.flow_solver import DeviceMesh
dev = torch.device("cuda:0")
VARS = [int(v) for v in os.environ.get("SMOOTH_VARS", "0").split(",")]
for name, var in [(n, v) for n in ("ys930", "ah93w145") for v in VARS]:
    os.environ["MDQ_SMOOTH_VAR"] = str(var)
    print(f"---- variant {var}", flush=True)
    coords, cells = load_mesh(name)
    m = DeviceMesh(coords, cells, dev)
    c0 = m.coords.clone()
    val = (m.nbr_ptr[1:] - m.nbr_ptr[:-1]).cpu().numpy()
    ob = m.on_boundary.cpu().numpy().astype(bool)
    print(f"{name}: valence histogram of interior vertices {np.bincount(val[~ob]).tolist()}", flush=True)
    for it in (0, 10):
        ts = []
        for _ in range(5):
            m.coords.copy_(c0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.smooth(it); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{name}: smooth({it}) median {np.median(ts):.3f} ms", flush=True)
    ts = []
    for _ in range(12):
        m.coords.copy_(c0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.smooth(50); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    REF = globals().setdefault("REF", {})
    if name not in REF:
        REF[name] = m.coords.clone()
    print(f"{name}: variant {var} result bit-identical to the first variant: {bool(torch.equal(REF[name], m.coords))}", flush=True)
    print(f"{name}: smooth(50) median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms  (nv {m.nv})", flush=True)
    import ctypes
    from meshdqn_b200 import _lib
    buf = (ctypes.c_longlong * 8)()
    L = _lib.lib(); L.mdq_debug_smooth_trace(buf)
    t = list(buf)
    print(f"  trace: setup {t[1]-t[0]} cyc, sweep {t[2]-t[1]} cyc = {t[5]-t[4]} ns -> {(t[2]-t[1])/max(1,(t[5]-t[4])):.3f} GHz, D {t[6]}, cyc/round {(t[2]-t[1])/(50*max(1,t[6])):.0f}", flush=True)
