"""FEniCS snapshot files in DOLFIN's dof order <-> this package's P2 / P1 layout.

The reference stores its T velocity / pressure snapshots as ``u.vector().get_local()`` arrays
(/root/reference/Env2DAirfoil.py:139-150 read, :432-449 write: ``save_velocities.npy`` [T, 2 (V0+E0)],
``save_pressures.npy`` [T, V0]) -- DOLFIN's own dof numbering, which depends on the DOLFIN version and its
reordering options.  That numbering is NOT restated here (it could not be checked without DOLFIN).  Instead the
conversion uses what DOLFIN itself reports next to the values on the machine that wrote them:

    xy_u   = V.tabulate_dof_coordinates().reshape(-1, 2)        # VectorFunctionSpace(mesh, 'Lagrange', 2)
    comp_u = np.zeros(V.dim(), dtype=np.int8); comp_u[V.sub(1).dofmap().dofs()] = 1
    xy_p   = Q.tabulate_dof_coordinates().reshape(-1, 2)        # FunctionSpace(mesh, 'Lagrange', 1)
    np.savez("dofmap.npz", xy_u=xy_u, comp_u=comp_u, xy_p=xy_p)

Every P2 dof sits on a mesh vertex or an edge midpoint and every P1 dof on a vertex, so matching the reported
coordinates against this package's dof points (vertices, then ``0.5 a + 0.5 b`` of the lexicographic edges) gives the
permutation whatever DOLFIN's ordering was.  Layout here: ``U [T, V+E, 2]`` (vertex dofs then edge dofs, both
components adjacent), ``P [T, V]``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def p2_points(coords, edges):
    coords = np.asarray(coords, dtype=np.float64)
    edges = np.asarray(edges)
    return np.concatenate([coords, 0.5 * coords[edges[:, 0]] + 0.5 * coords[edges[:, 1]]], axis=0)


def _match(points, xy, what):
    d, idx = cKDTree(points).query(np.asarray(xy, dtype=np.float64))
    scale = max(1.0, float(np.abs(points).max()))
    if d.max() > 1e-9 * scale:
        raise ValueError(f"{what}: a dof coordinate is {d.max():.3e} away from every dof point of the mesh "
                         "(different mesh, or the mesh was smoothed after the dof map was written?)")
    return idx


class DolfinDofMap:
    """Permutation between DOLFIN's dof order and this package's layout, built from DOLFIN's dof coordinates."""

    def __init__(self, coords, edges, xy_u, comp_u, xy_p):
        coords = np.asarray(coords, dtype=np.float64)
        self.nv, self.ne = len(coords), len(edges)
        pts = p2_points(coords, edges)
        comp_u = np.asarray(comp_u).astype(np.int64).ravel()
        if len(xy_u) != 2 * len(pts) or len(comp_u) != len(xy_u) or set(np.unique(comp_u)) - {0, 1}:
            raise ValueError(f"P2 vector space: expected {2 * len(pts)} dofs with components in {{0, 1}}, got {len(xy_u)}")
        if len(xy_p) != self.nv:
            raise ValueError(f"P1 space: expected {self.nv} dofs, got {len(xy_p)}")
        self.u_slot = 2 * _match(pts, xy_u, "velocity space") + comp_u      # flat index into U[t].reshape(-1)
        self.p_slot = _match(coords, xy_p, "pressure space")
        if len(np.unique(self.u_slot)) != len(self.u_slot) or len(np.unique(self.p_slot)) != len(self.p_slot):
            raise ValueError("dof map is not a permutation (duplicate dof coordinates / components)")

    @classmethod
    def load(cls, path, coords, edges):
        z = np.load(path)
        return cls(coords, edges, z["xy_u"], z["comp_u"], z["xy_p"])

    # DOLFIN order -> package layout
    def velocities(self, values):
        values = np.atleast_2d(np.asarray(values, dtype=np.float64))
        out = np.empty((values.shape[0], 2 * (self.nv + self.ne)))
        out[:, self.u_slot] = values
        return out.reshape(values.shape[0], self.nv + self.ne, 2)

    def pressures(self, values):
        values = np.atleast_2d(np.asarray(values, dtype=np.float64))
        out = np.empty((values.shape[0], self.nv))
        out[:, self.p_slot] = values
        return out

    # package layout -> DOLFIN order (what set_plot_dir would have written)
    def dolfin_velocities(self, U):
        U = np.asarray(U, dtype=np.float64)
        return U.reshape(U.shape[0], -1)[:, self.u_slot]

    def dolfin_pressures(self, P):
        return np.asarray(P, dtype=np.float64)[:, self.p_slot]
