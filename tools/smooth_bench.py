"""mdq_mesh_smooth(50) on the ys930 fixture: median CUDA-event time of the single-CTA ordered Gauss-Seidel sweep."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_mesh
from meshdqn_b200.flow_solver import DeviceMesh
dev = torch.device("cuda:0")
for name in ("ys930", "ah93w145"):
    coords, cells = load_mesh(name)
    m = DeviceMesh(coords, cells, dev)
    c0 = m.coords.clone()
    ts = []
    for _ in range(12):
        m.coords.copy_(c0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.smooth(50); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name}: smooth(50) median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms  (nv {m.nv})", flush=True)
