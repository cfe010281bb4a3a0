"""TEST INFRASTRUCTURE -- CPU oracle for the GEOMETRIC half of BASELINE.json configs[4] in "throughput mode".

NEW SEMANTICS, not the reference's: the reference evaluates a removal with a global Delaunay + 50 global smoothing
sweeps + a full re-interpolation (Env2DAirfoil.py:452-512,547-602).  Throughput mode evaluates a CANDIDATE removal of
vertex v on the current mesh M0 without moving any other vertex:

  1. the hole of v is re-triangulated locally (the cells come from meshdqn_b200/candidates.retriangulate_star);
  2. fields on the variant mesh M_v: every vertex dof and every edge dof whose edge survives keeps its value; a NEW
     edge (a diagonal of the hole) gets the ORIGINAL P2 velocity evaluated at its midpoint (point location among the
     old star cells, barycentric P2 evaluation -- the same arithmetic as the interpolation of the full step);
     the P1 pressure has vertex dofs only, so nothing is evaluated for it;
  3. drag / lift over the airfoil facets of M_v.

``variant_full`` rebuilds M_v and integrates over ALL its airfoil facets with the oracle's own routines;
``variant_delta`` is the algorithm a GPU kernel would run per candidate -- touch only the hole: evaluate the new
edges, subtract the traction of the airfoil facets whose cell disappears, add that of the cells that replace them.
tests/test_candidates_cpu.py checks delta == full to 1e-10 relative.  No GPU implementation exists yet.
"""
import numpy as np

from . import geom


def facet_traction(coords, cv, k, Udof, Pab, mu):
    """len * (sigma . n) of the facet opposite local vertex k of cell cv (ascending ids) for every snapshot.

    Udof [T, 6, 2]: velocity at the cell's dofs (3 vertices, then the edges opposite local vertices 0, 1, 2);
    Pab [T, 2]: pressure at the facet's two vertices.  Restates orc_drag_lift's per-facet arithmetic."""
    X, Y = coords[cv, 0], coords[cv, 1]
    det = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0])
    gx = np.array([(Y[1] - Y[2]) / det, (Y[2] - Y[0]) / det, (Y[0] - Y[1]) / det])
    gy = np.array([(X[2] - X[1]) / det, (X[0] - X[2]) / det, (X[1] - X[0]) / det])
    lam = np.array([0.5, 0.5, 0.5])
    lam[k] = 0.0
    bx, by = np.empty(6), np.empty(6)
    for a in range(3):
        s = 4.0 * lam[a] - 1.0
        bx[a], by[a] = s * gx[a], s * gy[a]
    bx[3] = 4.0 * (lam[1] * gx[2] + lam[2] * gx[1]); by[3] = 4.0 * (lam[1] * gy[2] + lam[2] * gy[1])
    bx[4] = 4.0 * (lam[0] * gx[2] + lam[2] * gx[0]); by[4] = 4.0 * (lam[0] * gy[2] + lam[2] * gy[0])
    bx[5] = 4.0 * (lam[0] * gx[1] + lam[1] * gx[0]); by[5] = 4.0 * (lam[0] * gy[1] + lam[1] * gy[0])
    i, j = (k + 1) % 3, (k + 2) % 3
    ex, ey = X[j] - X[i], Y[j] - Y[i]
    ln = np.sqrt(ex * ex + ey * ey)
    nx, ny = ey / ln, -ex / ln
    mx, my = 0.5 * X[i] + 0.5 * X[j], 0.5 * Y[i] + 0.5 * Y[j]
    if nx * (X[k] - mx) + ny * (Y[k] - my) > 0.0:
        nx, ny = -nx, -ny
    uxx = (Udof[:, :, 0] * bx).sum(1); uxy = (Udof[:, :, 0] * by).sum(1)
    uyx = (Udof[:, :, 1] * bx).sum(1); uyy = (Udof[:, :, 1] * by).sum(1)
    pm = 0.5 * Pab[:, 0] + 0.5 * Pab[:, 1]
    sxx = 2.0 * mu * uxx - pm
    sxy = mu * (uxy + uyx)
    syy = 2.0 * mu * uyy - pm
    return ln * (sxx * nx + sxy * ny), ln * (sxy * nx + syy * ny)


class Base:
    """Per-mesh data shared by all candidates: topology of M0, its facet tags, an edge lookup, base drag / lift."""

    def __init__(self, coords, cells, U0, P0, mu):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.nv = len(coords)
        self.topo = geom.Topology(cells, self.nv)
        self.U0, self.P0, self.mu = np.asarray(U0, dtype=np.float64), np.asarray(P0, dtype=np.float64), float(mu)
        self.tags = geom.facet_tags(self.coords, self.topo)
        e = self.topo.edges.astype(np.int64)
        self.edge_id = {(int(a), int(b)): i for i, (a, b) in enumerate(e)}
        self.drag, self.lift = geom.drag_lift(self.coords, self.topo, self.tags, self.U0, self.P0, self.mu)

    def star(self, v):
        return self.topo.vc_idx[self.topo.vc_ptr[v]:self.topo.vc_ptr[v + 1]].astype(np.int64)

    def new_edge_values(self, v, new_cells):
        """{(a, b): u [T, 2]} for the edges of the new cells that M0 does not have: P2 evaluation of the original field at
        the midpoint, located among the old star cells."""
        star = self.star(v)
        need = []
        for c in new_cells:
            for a, b in ((c[0], c[1]), (c[0], c[2]), (c[1], c[2])):
                key = (int(min(a, b)), int(max(a, b)))
                if key not in self.edge_id and key not in need:
                    need.append(key)
        if not need:
            return {}
        mids = np.array([0.5 * self.coords[a] + 0.5 * self.coords[b] for a, b in need])
        loc, nmiss, _ = geom.locate(mids, self.coords, self.topo.cells[star])
        u, _ = geom.eval_fields(mids, 0, star[loc].astype(np.int32), self.coords, self.topo, self.U0, self.P0)
        return {k: u[:, i, :] for i, k in enumerate(need)}

    def cell_dofs(self, c, new_vals):
        """Velocity at the six dofs of a cell with ascending ORIGINAL vertex ids c (old or new cell)."""
        out = np.empty((self.U0.shape[0], 6, 2))
        out[:, :3] = self.U0[:, list(c)]
        for li, (a, b) in enumerate(((c[1], c[2]), (c[0], c[2]), (c[0], c[1]))):      # edge opposite local vertex li
            key = (int(a), int(b))
            out[:, 3 + li] = self.U0[:, self.nv + self.edge_id[key]] if key in self.edge_id else new_vals[key]
        return out


def variant_delta(base: Base, v, new_cells):
    """(drag [T], lift [T], number of new-edge evaluations) of the variant, touching only the hole of v."""
    new_vals = base.new_edge_values(v, new_cells)
    drag, lift = base.drag.copy(), base.lift.copy()
    topo = base.topo
    for c in base.star(v):                                   # airfoil facets whose cell disappears
        for k in range(3):
            e = topo.cell_edges[c, k]
            if base.tags[e] == 1:
                cv = topo.cells[c].astype(np.int64)
                i, j = (k + 1) % 3, (k + 2) % 3
                fx, fy = facet_traction(base.coords, cv, k, base.cell_dofs(cv, {}), base.P0[:, [cv[i], cv[j]]], base.mu)
                drag -= fx
                lift -= fy
    for c in np.asarray(new_cells, dtype=np.int64):          # ... and the cells that replace them
        for k in range(3):
            i, j = (k + 1) % 3, (k + 2) % 3
            key = (int(c[i]), int(c[j])) if c[i] < c[j] else (int(c[j]), int(c[i]))
            e = base.edge_id.get(key)
            if e is not None and base.tags[e] == 1:
                fx, fy = facet_traction(base.coords, c, k, base.cell_dofs(c, new_vals), base.P0[:, [c[i], c[j]]], base.mu)
                drag += fx
                lift += fy
    return drag, lift, len(new_vals)


def variant_full(base: Base, v, new_cells):
    """The same quantities by rebuilding the whole variant mesh and integrating over all its airfoil facets."""
    keep = np.ones(base.nv, dtype=bool)
    keep[v] = False
    new_id = np.cumsum(keep) - 1
    old_cells = base.topo.cells[~(base.topo.cells == v).any(1)].astype(np.int64)
    cells_v = new_id[np.concatenate([old_cells, np.asarray(new_cells, dtype=np.int64)])]
    coords_v = base.coords[keep]
    topo_v = geom.Topology(cells_v, base.nv - 1)
    tags_v = geom.facet_tags(coords_v, topo_v)
    new_vals = base.new_edge_values(v, new_cells)
    old_of = np.nonzero(keep)[0]
    T = base.U0.shape[0]
    U = np.empty((T, topo_v.nv + topo_v.ne, 2))
    U[:, :topo_v.nv] = base.U0[:, old_of]
    for e, (a, b) in enumerate(topo_v.edges):
        key = (int(old_of[a]), int(old_of[b]))
        U[:, topo_v.nv + e] = base.U0[:, base.nv + base.edge_id[key]] if key in base.edge_id else new_vals[key]
    P = base.P0[:, old_of]
    return geom.drag_lift(coords_v, topo_v, tags_v, U, P, base.mu)
