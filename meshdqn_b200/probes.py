"""Drag / lift surface-integral probes on device.

Mirrors ``DragProbe`` / ``LiftProbe`` of /root/reference/probes.py:13-50:
``sample(u, p)`` = assemble( ((2 mu sym(grad u) - p I) n) . e_dir * ds(tag) ) over the facets tagged
``tags`` (1 = airfoil).  The integrand is linear on each facet (P2 velocity, P1 pressure), so the
midpoint rule FFC generates is exact.  ``u`` is a P2 nodal array [V+E, 2] (or [T, V+E, 2]), ``p`` a P1
array [V] (or [T, V]) on the flow solver's current mesh; both probes share one kernel launch
(mdq_drag_lift), cached per (u, p) pair so calling drag then lift costs one launch.
The unused PenetratedDragProbe / *ANN probes (probes.py:53-100) are dead code in the reference and
are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def drag_lift_device(fs, U, P):
    """[2, T] f64 device tensor: row 0 drag, row 1 lift, for snapshots U [T, V+E, 2], P [T, V]."""
    m = fs.mesh
    cached = fs.__dict__.get("_dl_cache")
    if cached is not None and cached[0] is m and cached[1] is U and cached[2] is P:
        return cached[3]            # evaluated by the interpolation launch's epilogue (mdq_interpolate_drag_lift)
    T = int(U.shape[0])
    if U.shape[1] != m.nv + m.ne or P.shape[1] != m.nv:
        raise ValueError(f"field sizes {tuple(U.shape)}, {tuple(P.shape)} do not match mesh (V={m.nv}, E={m.ne})")
    out = torch.empty((2, T), dtype=torch.float64, device=fs.device)
    L = _lib.lib()
    p = _lib.ptr
    with torch.cuda.device(fs.device):
        rc = L.mdq_drag_lift(p(m.coords), p(m.cells), p(m.cell_edges), m.nv, m.ne, p(fs.tags), p(m.edge_cell), T,
                             p(U), p(P), float(fs.viscosity), p(out), _lib.stream_ptr())
    _lib.check(rc, "mdq_drag_lift")
    return out


class _Probe:
    _row = 0

    def __init__(self, mu, flow_solver, tags, flow_dir=None):
        if list(tags) != [1]:
            raise NotImplementedError("only the airfoil tag [1] is used by the reference (flow_solver.py:186-187)")
        self.mu = mu
        self.fs = flow_solver
        self.tags = tags
        self.dim = 2

    def sample(self, u, p):
        U = torch.as_tensor(u, dtype=torch.float64, device=self.fs.device)
        P = torch.as_tensor(p, dtype=torch.float64, device=self.fs.device)
        single = U.dim() == 2
        if single:
            U, P = U[None], P[None]
        P = P.reshape(P.shape[0], -1)
        out = drag_lift_device(self.fs, U.contiguous(), P.contiguous())[self._row].cpu().numpy()
        return float(out[0]) if single else out


class DragProbe(_Probe):
    """probes.py:13-31 (flow_dir = (1, 0))."""
    _row = 0


class LiftProbe(_Probe):
    """probes.py:33-50 (flow_dir = (0, 1))."""
    _row = 1
