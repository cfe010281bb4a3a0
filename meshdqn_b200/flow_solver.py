"""Mesh services of the reference's ``flow_solver.FlowSolver`` on device (no Navier-Stokes solve).

Mirrors the part of /root/reference/flow_solver.py that the per-action hot path uses:

* ``FlowSolver(flow_params, geometry_params, solver_params)`` : load mesh (:58-62), ``smooth(50)``
  (:65-67), removable mask (:75-78), ``mark_boundaries`` (:194-226), probes (:186-187)
* ``.mesh`` (``DeviceMesh``: ``coordinates()``, ``cells()``, ``num_vertices()``), ``.removable``,
  ``.remesh(mesh)`` (:233-266,341-359, non-DEPLOY), ``.drag_probe`` / ``.lift_probe``, ``.num_vertices``

The one-time IPCS solve (``evolve``, :362-396) stays in FEniCS on the host and is out of scope:
``evolve()``/``deploy()`` raise.  All arrays live in HBM; host mirrors are fetched lazily.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .probes import DragProbe, LiftProbe
from .xdmf import read_xdmf_mesh


def _dev(device):
    if device is None:
        device = "cuda"
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("meshdqn_b200 has no CPU path: a CUDA device is required")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class DeviceMesh:
    """Triangle mesh resident in HBM with its topology tables (what ``dolfin.Mesh`` + ``mesh.init()`` give).

    Layout: ``coords`` f64 [V,2]; ``cells`` i32 [C,3] ascending ids per cell; ``edges`` i32 [E,2]
    lexicographic; ``cell_edges`` i32 [C,3] (edge opposite local vertex i); CSR vertex->neighbours
    (ascending) and vertex->cells (ascending); ``on_boundary`` u8 [V]; ``bverts`` ascending boundary ids.
    """

    def __init__(self, coords, cells, device=None):
        self.device = _dev(device)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cells = np.sort(np.ascontiguousarray(cells, dtype=np.int32), axis=1)
        self.nv, self.nc = int(coords.shape[0]), int(cells.shape[0])
        d = self.device
        self.coords = torch.from_numpy(coords).to(d)
        self.cells = torch.from_numpy(cells).to(d)
        nv, nc = self.nv, self.nc
        i32 = dict(dtype=torch.int32, device=d)
        self.nbr_ptr = torch.empty(nv + 1, **i32)
        self.nbr_idx = torch.empty(6 * nc, **i32)
        self.vc_ptr = torch.empty(nv + 1, **i32)
        self.vc_idx = torch.empty(3 * nc, **i32)
        self.edge_base = torch.empty(nv + 1, **i32)
        self.edges_buf = torch.empty((3 * nc, 2), **i32)
        self.cell_edges = torch.empty((nc, 3), **i32)
        self.edge_ncells_buf = torch.empty(3 * nc, **i32)
        self.edge_cell_buf = torch.zeros(3 * nc, **i32)
        self.on_boundary = torch.empty(nv, dtype=torch.uint8, device=d)
        self.bverts_buf = torch.empty(nv, **i32)
        self.counts = torch.zeros(4, **i32)
        scratch = torch.empty(2 * nv + 2, **i32)
        L = _lib.lib()
        p = _lib.ptr
        with torch.cuda.device(d):
            rc = L.mdq_mesh_topology(p(self.cells), nc, nv, p(self.nbr_ptr), p(self.nbr_idx), p(self.vc_ptr), p(self.vc_idx),
                                     p(self.edge_base), p(self.edges_buf), p(self.cell_edges), p(self.edge_ncells_buf),
                                     p(self.edge_cell_buf), p(self.on_boundary), p(self.bverts_buf), p(self.counts),
                                     p(scratch), _lib.stream_ptr())
        _lib.check(rc, "mdq_mesh_topology")
        cnt = self.counts.cpu().tolist()  # one small sync: edge / boundary counts size everything downstream
        if cnt[3]:
            raise RuntimeError("mesh vertex with more than 96 neighbours is not supported")
        self.ne, self.nb = int(cnt[0]), int(cnt[1])
        self.edges = self.edges_buf[: self.ne]
        self.edge_ncells = self.edge_ncells_buf[: self.ne]
        self.edge_cell = self.edge_cell_buf[: self.ne]
        self.bverts = self.bverts_buf[: self.nb]
        self._host = {}

    # dolfin.Mesh-like accessors (host copies, cached until the coordinates change)
    def coordinates(self):
        if "coords" not in self._host:
            self._host["coords"] = self.coords.cpu().numpy()
        return self._host["coords"]

    def cells_host(self):
        if "cells" not in self._host:
            self._host["cells"] = self.cells.cpu().numpy()
        return self._host["cells"]

    def num_vertices(self):
        return self.nv

    def num_cells(self):
        return self.nc

    def boundary_vertices(self):
        if "bverts" not in self._host:
            self._host["bverts"] = self.bverts.cpu().numpy()
        return self._host["bverts"]

    def smooth(self, iters=50):
        """``Mesh.smooth(iters)`` (flow_solver.py:67,237) in exact Gauss-Seidel vertex order."""
        self.smooth_status = torch.zeros(1, dtype=torch.int32, device=self.device)
        L = _lib.lib()
        p = _lib.ptr
        with torch.cuda.device(self.device):
            rc = L.mdq_mesh_smooth(p(self.coords), self.nv, self.nc, p(self.nbr_ptr), p(self.nbr_idx), p(self.vc_ptr), p(self.vc_idx),
                                   p(self.cells), p(self.on_boundary), int(iters), p(self.smooth_status), _lib.stream_ptr())
        _lib.check(rc, "mdq_mesh_smooth")
        self._host.pop("coords", None)


class FlowSolver:
    def __init__(self, flow_params, geometry_params, solver_params, mesh=None, device=None):
        self.device = _dev(device)
        self.viscosity = float(flow_params["mu"])
        self.density = float(flow_params.get("rho", 1.0))
        self.smooth = bool(solver_params.get("smooth", False))
        self.solver_type = solver_params.get("la_solve", "lu")
        self.dt = solver_params.get("dt")
        self.DEPLOY = False
        if mesh is None:
            mesh = read_xdmf_mesh(geometry_params["mesh"])  # flow_solver.py:58-62
        self.drag_probe = DragProbe(self.viscosity, self, tags=[1])
        self.lift_probe = LiftProbe(self.viscosity, self, tags=[1])
        self.accumulated_drag, self.accumulated_lift = [], []
        self.remesh(mesh)

    def remesh(self, mesh):
        """flow_solver.py:233-266,341-359 (non-DEPLOY): adopt the mesh, smooth, tag facets, removable mask."""
        if not isinstance(mesh, DeviceMesh):
            mesh = DeviceMesh(mesh[0], mesh[1], self.device)
        self.mesh = mesh
        if self.smooth:
            self.mesh.smooth(50)
        self.mark_boundaries()
        self.num_vertices = mesh.nv
        self._removable_host = None

    def mark_boundaries(self):
        """Facet tags (1 = airfoil) and the removable mask, one launch pair (flow_solver.py:194-226,247-250)."""
        m = self.mesh
        self.tags = torch.empty(m.ne, dtype=torch.int32, device=self.device)
        self.removable_dev = torch.empty(m.nv, dtype=torch.uint8, device=self.device)
        L = _lib.lib()
        p = _lib.ptr
        with torch.cuda.device(self.device):
            rc = L.mdq_mesh_tags_removable(p(m.coords), m.nv, p(m.edges), p(m.edge_ncells), m.ne, p(m.bverts), m.nb,
                                           p(self.tags), p(self.removable_dev), _lib.stream_ptr())
        _lib.check(rc, "mdq_mesh_tags_removable")
        return self.tags

    @property
    def removable(self):
        """bool ndarray [V] like the reference's python list (flow_solver.py:76-78)."""
        if self._removable_host is None:
            self._removable_host = self.removable_dev.cpu().numpy().astype(bool)
        return self._removable_host

    def deploy(self):
        raise NotImplementedError("DEPLOY re-assembles the Navier-Stokes system (flow_solver.py:268-339): out of scope, "
                                  "it stays in FEniCS on the host")

    def evolve(self):
        raise NotImplementedError("the one-time IPCS Navier-Stokes solve (flow_solver.py:362-396) stays in FEniCS; "
                                  "pass its snapshots through agent_params['u'] / ['p']")
