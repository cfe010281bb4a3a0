"""BASELINE.json configs[4] workload generator (meshdqn_b200/candidates.py): the local star re-triangulation against a
global Delaunay of the remaining points, and the invariants of the candidate state graphs."""
import numpy as np
import torch
from scipy.spatial import Delaunay

from meshdqn_b200 import candidates as C
from meshdqn_b200.synthetic import field_values, synthetic_airfoil_mesh


def _mesh():
    coords, cells, n_ring = synthetic_airfoil_mesh(6000, seed=3)
    return coords, cells.astype(np.int64), np.arange(4, 4 + n_ring)


def test_local_star_retriangulation_equals_global_delaunay():
    coords, cells, ring_ids = _mesh()
    rem, on_b = C.removable_vertices(coords, cells)
    assert rem.sum() > 1000 and not rem[ring_ids].any()
    rng = np.random.RandomState(0)
    checked = tiled = 0
    area = lambda cs: sum(abs(C._poly_area(coords[c])) for c in cs)
    as_set = lambda a: {tuple(r) for r in a.tolist()}
    # vertices next to the airfoil (the ones the candidates actually remove) and vertices deep in the mesh
    d = C.polygon_distance(coords[rem], coords[ring_ids])
    near_airfoil = np.nonzero(rem)[0][np.argsort(d)[:40]]
    for v in np.concatenate([near_airfoil, rng.choice(np.nonzero(rem)[0], 40, replace=False)]):
        cv = cells[(cells == v).any(1)]
        new = C.retriangulate_star(coords, cv, int(v))
        if new is None:
            continue
        ring = np.array([u for u in np.unique(cv) if u != v])
        # the new cells tile the hole exactly and use its ring only
        assert abs(area(new) - area(cv)) <= 1e-12 * max(1.0, area(cv)) and np.isin(new, ring).all()
        tiled += 1
        if on_b[ring].any():
            continue        # a hole that touches the airfoil / walls: a global Delaunay of the point CLOUD ignores the boundary
                            # (it may bridge the airfoil); the local, polygon-constrained triangulation is the defined result
        keep = np.ones(len(coords), dtype=bool)
        keep[v] = False
        tri = Delaunay(coords[keep])                                   # what Env2DAirfoil.py:487 does after a removal
        ids = np.nonzero(keep)[0]
        g = np.sort(ids[tri.simplices], axis=1)
        assert as_set(new) <= as_set(g), v                             # every local cell is a cell of the global triangulation
        # outside the hole nothing changed: every old cell near it is still a global cell
        old = np.sort(cells[~(cells == v).any(1)], axis=1)
        near = old[np.isin(old, ring).any(1) & ~on_b[old].any(1)]
        assert as_set(near) <= as_set(g), v
        checked += 1
    assert checked >= 25 and tiled >= 60


def test_candidate_state_graphs_invariants():
    coords, cells, ring_ids = _mesh()
    u, p = field_values(coords, 5, 0)
    graphs, meta = C.candidate_state_graphs(coords, cells, ring_ids, u, p, n_candidates=400, n_closest=180)
    assert 380 <= len(graphs) <= 400 and len(meta) == len(graphs)
    assert meta[:, 0].max() == 2 and meta[0, 0] == 0                    # three window offsets for 400 candidates
    feats = np.concatenate([coords, u.transpose(1, 0, 2).reshape(len(coords), -1), p.T], 1).astype(np.float32)
    seen = set()
    for g, (o, a, v) in zip(graphs[::7], meta[::7]):
        assert g.x.shape == (180, 17) and g.x.dtype == torch.float32 and g.edge_index.dtype == torch.int64
        assert g.edge_index.shape[0] == 2 and g.edge_index.shape[1] % 3 == 0 and g.edge_index.shape[1] > 60
        assert int(g.edge_index.min()) >= 0 and int(g.edge_index.max()) < 180
        xy = g.x[:, :2].numpy()
        assert not (np.abs(xy - coords[v].astype(np.float32)).max(1) == 0).any()       # the removed vertex is gone
        # rows are mesh vertices with their own features
        row = np.nonzero((feats[:, :2] == xy[5]).all(1))[0]
        assert len(row) == 1 and np.array_equal(feats[row[0]], g.x[5].numpy())
        # quirk B3 edge pattern: (i1, i2), (i1, i3), (i2, i3) per cell
        e = g.edge_index.numpy().T.reshape(-1, 3, 2)
        assert (e[:, 0, 0] == e[:, 1, 0]).all() and (e[:, 0, 1] == e[:, 2, 0]).all() and (e[:, 1, 1] == e[:, 2, 1]).all()
        seen.add(g.edge_index.shape[1])
    assert len(seen) > 1                                                 # the variants differ
    g2, m2 = C.candidate_state_graphs(coords, cells, ring_ids, u, p, n_candidates=400, n_closest=180)
    assert np.array_equal(meta, m2) and all(torch.equal(a.edge_index, b.edge_index) for a, b in zip(graphs[:50], g2[:50]))


def test_throughput_mode_delta_equals_full_recompute():
    """oracle/candidates_ref.py (NEW semantics for configs[4]'s geometric half): per candidate, evaluating only the hole --
    new-edge dofs from the original field, minus / plus the traction of the airfoil facets whose cell changes -- gives the
    drag / lift of the fully rebuilt variant mesh to 1e-10 relative."""
    from oracle import candidates_ref as R
    from meshdqn_b200.synthetic import synthetic_fields
    coords, cells, ring_ids = _mesh()
    base_topo_edges = R.geom.Topology(cells, len(coords)).edges
    U0, P0 = synthetic_fields(coords, base_topo_edges, T=5, seed=0)
    base = R.Base(coords, cells, U0, P0, mu=1e-3)
    assert (base.tags == 1).sum() == len(ring_ids)                       # one airfoil facet per ring vertex
    rem, _ = C.removable_vertices(coords, cells)
    d = C.polygon_distance(coords[rem], coords[ring_ids])
    order = np.nonzero(rem)[0][np.argsort(d)]
    changed = same = 0
    for v in list(order[:40]) + list(order[400:410]):                    # first-layer vertices and vertices away from the airfoil
        cv = cells[(cells == v).any(1)]
        new = C.retriangulate_star(coords, cv, int(v))
        if new is None:
            continue
        dD, dL, n_eval = R.variant_delta(base, int(v), new)
        fD, fL = R.variant_full(base, int(v), new)
        scale = max(np.abs(fD).max(), np.abs(fL).max())
        assert np.abs(dD - fD).max() <= 1e-10 * scale and np.abs(dL - fL).max() <= 1e-10 * scale, v
        assert n_eval == len(new) - 1 or len(new) == 1                   # a k-gon hole has k-2 cells and k-3 new diagonals
        if np.abs(fD - base.drag).max() > 1e-12 * scale:
            changed += 1
        else:
            same += 1
    assert changed >= 10 and same >= 10       # removals over an airfoil facet change the integral, the others leave it alone
