"""oracle/geom.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py; parity unpinned).

numpy topology + ctypes bindings to oracle/geom_oracle.c.  Restates the DOLFIN
mesh services the reference's step uses:

* ``Topology``          -- Mesh.init / MeshEditor.close ordering   (Env2DAirfoil.py:499-509)
* ``smooth``            -- Mesh.smooth(50)                          (flow_solver.py:67,237)
* ``facet_tags``        -- FlowSolver.mark_boundaries               (flow_solver.py:194-226)
* ``removable_mask``    -- coord not in bmesh.coordinates()         (flow_solver.py:75-78,247-250)
* ``polygon_distance``  -- shapely Polygon.distance(Point)          (Env2DAirfoil.py:232,240-241)
* ``locate`` / ``eval_fields`` -- Function.interpolate              (Env2DAirfoil.py:556-568)
* ``drag_lift``         -- DragProbe/LiftProbe.sample               (probes.py:23-31,43-50)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libgeom_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_locate.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Topology:
    """Edges, cell->edge map, boundary and adjacency of a triangle mesh (A.1).

    Cells are sorted ascending per cell (DOLFIN orders cell vertices on close()).
    Edges are the unique (a<b) pairs in lexicographic order; ``cell_edges[c,i]``
    is the edge opposite local vertex i (DOLFIN facet numbering).
    """

    def __init__(self, cells, nv):
        cells = np.sort(_i32(cells), axis=1)
        self.cells = cells
        self.nv = int(nv)
        a = np.concatenate([cells[:, 1], cells[:, 0], cells[:, 0]])
        b = np.concatenate([cells[:, 2], cells[:, 2], cells[:, 1]])
        key = a.astype(np.int64) * self.nv + b.astype(np.int64)
        ukey, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        self.edges = np.stack([ukey // self.nv, ukey % self.nv], axis=1).astype(np.int32)
        self.ne = len(ukey)
        nc = len(cells)
        self.cell_edges = np.ascontiguousarray(inv.reshape(3, nc).T.astype(np.int32))
        self.edge_ncells = cnt.astype(np.int32)
        bmask = np.zeros(self.nv, dtype=bool)
        be = self.edges[cnt == 1]
        bmask[be.ravel()] = True
        self.on_boundary = bmask
        self.boundary_vertices = np.nonzero(bmask)[0].astype(np.int32)
        # vertex -> neighbours (ascending) CSR
        src = np.concatenate([self.edges[:, 0], self.edges[:, 1]])
        dst = np.concatenate([self.edges[:, 1], self.edges[:, 0]])
        order = np.lexsort((dst, src))
        self.nbr_idx = dst[order].astype(np.int32)
        self.nbr_ptr = np.zeros(self.nv + 1, dtype=np.int32)
        np.cumsum(np.bincount(src, minlength=self.nv), out=self.nbr_ptr[1:])
        # vertex -> incident cells (ascending cell index) CSR
        vflat = cells.ravel()
        cidx = np.repeat(np.arange(nc, dtype=np.int32), 3)
        order = np.lexsort((cidx, vflat))
        self.vc_idx = cidx[order].astype(np.int32)
        self.vc_ptr = np.zeros(self.nv + 1, dtype=np.int32)
        np.cumsum(np.bincount(vflat, minlength=self.nv), out=self.vc_ptr[1:])

    def p2_points(self, coords):
        """P2 dof points: vertices then edge midpoints 0.5*a + 0.5*b (A.7)."""
        mid = 0.5 * coords[self.edges[:, 0]] + 0.5 * coords[self.edges[:, 1]]
        return np.concatenate([coords, mid], axis=0)

    def tagged_facets(self, tags, tag=1):
        """(cell, local opposite-vertex index) of exterior facets with ``tag``, ascending edge id."""
        want = np.zeros(self.ne, dtype=bool)
        want[np.nonzero(tags == tag)[0]] = True
        fc, fl = np.nonzero(want[self.cell_edges])
        eid = self.cell_edges[fc, fl]
        order = np.argsort(eid, kind="stable")
        return fc[order].astype(np.int32), fl[order].astype(np.int32)


def smooth(coords, topo: Topology, iters=50):
    x = _f64(coords).copy()
    ob = np.ascontiguousarray(topo.on_boundary, dtype=np.uint8)
    lib().orc_smooth(_p(x), ctypes.c_int(topo.nv), _p(topo.nbr_ptr), _p(topo.nbr_idx), _p(topo.vc_ptr),
                     _p(topo.vc_idx), _p(topo.cells), _p(ob), ctypes.c_int(iters))
    return x


def facet_tags(coords, topo: Topology):
    x = _f64(coords)
    tags = np.empty(topo.ne, dtype=np.int32)
    lib().orc_facet_tags(_p(x), _p(topo.edges), _p(topo.edge_ncells), ctypes.c_int(topo.ne), _p(tags))
    return tags


def removable_mask(coords, topo: Topology):
    x = _f64(coords)
    out = np.empty(topo.nv, dtype=np.uint8)
    bv = _i32(topo.boundary_vertices)
    lib().orc_removable(_p(x), ctypes.c_int(topo.nv), _p(bv), ctypes.c_int(len(bv)), _p(out))
    return out.astype(bool)


def polygon_distance(pts, ring):
    pts = _f64(pts)
    ring = _f64(ring)
    out = np.empty(len(pts), dtype=np.float64)
    lib().orc_polygon_distance(_p(pts), ctypes.c_int(len(pts)), _p(ring), ctypes.c_int(len(ring)), _p(out))
    return out


def locate(pts, coords, cells, tol=1e-12):
    pts = _f64(pts)
    x = _f64(coords)
    cells = _i32(cells)
    out = np.empty(len(pts), dtype=np.int32)
    d2 = np.empty(len(pts), dtype=np.float64)
    nmiss = lib().orc_locate(_p(pts), ctypes.c_int(len(pts)), _p(x), _p(cells), ctypes.c_int(len(cells)),
                             ctypes.c_double(tol), _p(out), _p(d2))
    return out, int(nmiss), d2


def eval_fields(pts, n_p1, cell_of, coords0, topo0: Topology, U, P):
    """U [T, V0+E0, 2], P [T, V0] -> u [T, len(pts), 2], p [T, n_p1]."""
    pts = _f64(pts)
    U = _f64(U)
    P = _f64(P)
    T = U.shape[0]
    out_u = np.empty((T, len(pts), 2), dtype=np.float64)
    out_p = np.empty((T, n_p1), dtype=np.float64)
    x = _f64(coords0)
    lib().orc_eval_fields(_p(pts), ctypes.c_int(len(pts)), ctypes.c_int(n_p1), _p(_i32(cell_of)), _p(x),
                          _p(topo0.cells), _p(topo0.cell_edges), ctypes.c_int(topo0.nv),
                          ctypes.c_int(U.shape[1]), ctypes.c_int(T), _p(U), _p(P), _p(out_u), _p(out_p))
    return out_u, out_p


def drag_lift(coords, topo: Topology, tags, U, P, mu):
    x = _f64(coords)
    U = _f64(U)
    P = _f64(P)
    T = U.shape[0]
    fc, fl = topo.tagged_facets(tags, 1)
    drag = np.empty(T, dtype=np.float64)
    lift = np.empty(T, dtype=np.float64)
    lib().orc_drag_lift(_p(x), _p(topo.cells), _p(topo.cell_edges), ctypes.c_int(topo.nv),
                        ctypes.c_int(U.shape[1]), ctypes.c_int(T), _p(U), _p(P), _p(fc), _p(fl),
                        ctypes.c_int(len(fc)), ctypes.c_double(mu), _p(drag), _p(lift))
    return drag, lift
