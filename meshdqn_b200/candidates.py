"""Workload generator for BASELINE.json configs[4]: "batched candidate evaluation: 8192 one-vertex-removed variants of a
250k-triangle mesh, sharded over 8 B200".

A candidate is a pair (window offset o, action a): the state the reference would present after ``o`` do-nothing steps
(``do_nothing_offset``, Env2DAirfoil.py:330-333: the N-closest window slides along the distance ranking) with its a-th
vertex removed.  For each candidate this module builds the state graph of the VARIANT mesh:

* the hole left by the vertex is re-triangulated LOCALLY -- the Delaunay triangulation of the star polygon's vertices,
  clipped to the polygon.  For a Delaunay mesh this equals what a global ``Delaunay(points \\ v)`` (Env2DAirfoil.py:487)
  returns inside the hole and nothing changes outside it (``tests/test_candidates_cpu.py`` checks exactly that);
* the window is re-cut from the ranking without the vertex (the next-closest vertex enters);
* nodes = the window's vertices with features [x, y | u(T x 2) | p(T)], edges = the three directed pairs of every cell
  whose vertices all lie in the window, in the cell's ascending-id order (quirk B3).

This is "throughput mode": the variants skip the reference's 50 global smoothing sweeps and the re-interpolation they
cause (vertices that do not move keep their nodal values), so the graphs are NOT what ``Env2DAirfoil.step`` returns --
they are a faithful one-vertex-removed workload for the batched Q-evaluation, which is the half of configs[4] this
repository runs on the GPU.  Host-side numpy / scipy; untimed set-up in ``bench.py``.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import Data


def removable_vertices(coords, cells):
    """The reference's mask (flow_solver.py:75-78,247-250): not on the boundary and sharing neither coordinate with a
    boundary vertex."""
    nv = len(coords)
    e = np.concatenate([cells[:, [0, 1]], cells[:, [0, 2]], cells[:, [1, 2]]])
    e = np.sort(e, axis=1)
    key = e[:, 0].astype(np.int64) * nv + e[:, 1]
    uniq, cnt = np.unique(key, return_counts=True)
    b = uniq[cnt == 1]
    on_b = np.zeros(nv, dtype=bool)
    on_b[b // nv] = True
    on_b[b % nv] = True
    bx, by = np.unique(coords[on_b, 0]), np.unique(coords[on_b, 1])
    return ~on_b & ~np.isin(coords[:, 0], bx) & ~np.isin(coords[:, 1], by), on_b


def polygon_distance(pts, ring):
    """Distance of points OUTSIDE a closed polygon to it (vectorised; Env2DAirfoil.py:232-241 via shapely)."""
    a, b = ring, np.roll(ring, -1, axis=0)
    ab = b - a
    l2 = (ab * ab).sum(1)
    out = np.empty(len(pts))
    for lo in range(0, len(pts), 2048):
        p = pts[lo:lo + 2048, None, :]
        t = np.clip(((p - a) * ab).sum(2) / l2, 0.0, 1.0)
        d = p - (a + t[..., None] * ab)
        out[lo:lo + 2048] = np.sqrt((d * d).sum(2).min(1))
    return out


def _star_ring(cells_v, v):
    """Vertices of the star polygon of v in cyclic order, from the cells containing v (closed star)."""
    nxt = {}
    for c in cells_v:
        o = [int(u) for u in c if u != v]
        nxt.setdefault(o[0], []).append(o[1])
        nxt.setdefault(o[1], []).append(o[0])
    if any(len(w) != 2 for w in nxt.values()):
        return None                                   # open star (boundary vertex): not a removal candidate
    start = min(nxt)
    ring, prev, cur = [start], None, start
    while True:
        a, b = nxt[cur]
        n = a if a != prev else b
        if n == start:
            break
        ring.append(n)
        prev, cur = cur, n
        if len(ring) > len(nxt):
            return None
    return ring if len(ring) == len(nxt) else None


def _poly_area(p):
    return 0.5 * (np.dot(p[:, 0], np.roll(p[:, 1], -1)) - np.dot(p[:, 1], np.roll(p[:, 0], -1)))


def _inside(poly, q):
    x, y = q
    ins = False
    n = len(poly)
    for i in range(n):
        (x0, y0), (x1, y1) = poly[i], poly[(i + 1) % n]
        if (y0 > y) != (y1 > y) and x < x0 + (y - y0) * (x1 - x0) / (y1 - y0):
            ins = not ins
    return ins


def retriangulate_star(coords, cells_v, v):
    """Cells (ascending vertex ids, int64 [k, 3]) that fill the hole of vertex v, or None when the local Delaunay fails
    (degenerate ring) or does not tile the hole exactly."""
    from scipy.spatial import Delaunay
    ring = _star_ring(cells_v, v)
    if ring is None or len(ring) < 3:
        return None
    ids = np.asarray(ring, dtype=np.int64)
    pts = coords[ids]
    if len(ring) == 3:
        new = ids[None, :]
    else:
        try:
            tri = Delaunay(pts)
        except Exception:
            return None
        keep = [s for s in tri.simplices if _inside(pts, pts[s].mean(0))]
        if not keep:
            return None
        new = ids[np.asarray(keep)]
    area = sum(abs(_poly_area(coords[c])) for c in new)
    if abs(area - abs(_poly_area(pts))) > 1e-9 * abs(_poly_area(pts)):
        return None
    return np.sort(new, axis=1)


def candidate_state_graphs(coords, cells, ring_ids, vertex_u, vertex_p, n_candidates=8192, n_closest=180):
    """State graphs of ``n_candidates`` one-vertex-removed variants (module docstring).

    ``ring_ids``: the airfoil's vertex ids in curve order; ``vertex_u`` [T, V, 2], ``vertex_p`` [T, V]: nodal snapshots.
    Returns (list of ``Data`` with x float32 [n_closest, 2 + 3T] and edge_index int64 [2, E], meta int32 [n, 3] =
    (offset, action, removed vertex id)).  Candidates whose hole cannot be re-triangulated locally are skipped."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    T = int(vertex_u.shape[0])
    rem, _ = removable_vertices(coords, cells)
    ring = coords[np.asarray(ring_ids)]
    lo, hi = ring.min(0), ring.max(0)
    n_off = (n_candidates + n_closest - 1) // n_closest
    need = n_off + n_closest + 1
    pad = 0.05
    while True:                                        # distance only where it can matter: a growing box around the airfoil
        box = rem & (coords[:, 0] > lo[0] - pad) & (coords[:, 0] < hi[0] + pad) & (coords[:, 1] > lo[1] - pad) & \
            (coords[:, 1] < hi[1] + pad)
        cand = np.nonzero(box)[0]
        if len(cand) >= need:
            d = polygon_distance(coords[cand], ring)
            if np.count_nonzero(d < pad) >= need:      # everything closer than the box margin has been seen
                break
        if box.sum() == rem.sum():
            d = polygon_distance(coords[cand], ring)
            break
        pad *= 2.0
    if len(cand) < need:
        raise ValueError(f"mesh has {len(cand)} removable vertices, {need} needed for {n_candidates} candidates")
    rk = cand[np.argsort(d, kind="stable")][:need]     # vertex ids by increasing distance to the airfoil
    region = np.zeros(len(coords), dtype=bool)
    region[rk] = True
    rc = cells[region[cells].any(1)]                   # every cell touching a ranked vertex
    feats = np.concatenate([coords, np.asarray(vertex_u).transpose(1, 0, 2).reshape(len(coords), 2 * T),
                            np.asarray(vertex_p).T], axis=1).astype(np.float32)
    star_cache = {}
    graphs, meta = [], []
    for o in range(n_off):
        span = rk[o:o + n_closest + 1]
        for a in range(n_closest):
            if len(graphs) >= n_candidates:
                break
            v = int(span[a])
            if v not in star_cache:
                has_v = (rc == v).any(1)
                star_cache[v] = (has_v, retriangulate_star(coords, rc[has_v], v))
            has_v, new = star_cache[v]
            if new is None:
                continue
            win = np.delete(span, a)                   # the next-closest vertex enters the window
            local = np.full(len(coords), -1, dtype=np.int64)
            local[win] = np.arange(n_closest)
            vc = np.concatenate([rc[~has_v], new])
            lc = local[vc]
            lc = lc[(lc >= 0).all(1)]
            if len(lc):
                ei = np.stack([np.stack([lc[:, 0], lc[:, 0], lc[:, 1]], 1).ravel(),
                               np.stack([lc[:, 1], lc[:, 2], lc[:, 2]], 1).ravel()])
            else:
                ei = np.zeros((2, 0), dtype=np.int64)
            graphs.append(Data(x=torch.from_numpy(feats[win]), edge_index=torch.from_numpy(np.ascontiguousarray(ei))))
            meta.append((o, a, v))
    return graphs, np.asarray(meta, dtype=np.int32)
