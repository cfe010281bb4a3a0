"""Second opinions on the oracle ("parity hardening").

The reference's third-party arithmetic (DOLFIN point location, PyG layers, GEOS distance) cannot be run here, so every
GPU parity claim is "vs oracle/".  These tests check the oracle's restatements against INDEPENDENT implementations that
are available in this image -- the same Qhull triangulation's own point locator, torch's sort / topk, dense-matrix
formulations of the graph layers, a vectorised numpy point-to-polygon distance -- so that a slip in the restatement
cannot silently become the golden truth.
"""
import numpy as np
import torch
from scipy.spatial import Delaunay

from conftest import load_mesh
from oracle import geom, gnn_ref


# ---------------------------------------------------------------- point location
def _delaunay_fixture(short):
    coords, _ = load_mesh(short)
    tri = Delaunay(coords)
    return coords, tri


def test_locate_agrees_with_qhull_find_simplex_on_random_points():
    """Random points inside the hull are (almost surely) strictly inside one simplex: the oracle's brute-force
    lowest-index rule and Qhull's walk must name the same cell."""
    for short in ("ys930", "ah93w145"):
        coords, tri = _delaunay_fixture(short)
        rng = np.random.RandomState(3)
        lo, hi = coords.min(0), coords.max(0)
        pts = lo + rng.rand(4000, 2) * (hi - lo)
        ref = tri.find_simplex(pts)
        inside = ref >= 0
        assert inside.sum() > 3000
        got, nmiss, _ = geom.locate(pts[inside], coords, tri.simplices.astype(np.int32))
        assert nmiss == 0
        assert np.array_equal(got, ref[inside])


def test_locate_on_dof_points_is_a_containing_cell():
    """Vertices and edge midpoints sit on cell boundaries (several cells contain them): Qhull may name any of them,
    the oracle names the lowest index -- both must CONTAIN the point, and the oracle's index must be the smallest one
    that does (checked with Qhull's own barycentric transform)."""
    coords, tri = _delaunay_fixture("ys930")
    simp = tri.simplices.astype(np.int32)
    topo = geom.Topology(simp, len(coords))
    pts = topo.p2_points(coords)
    got, nmiss, _ = geom.locate(pts, coords, simp)
    assert nmiss == 0
    # barycentric coordinates of every point in every simplex from Qhull's affine transforms
    T = tri.transform                                    # [nsimplex, 3, 2]
    sub = np.random.RandomState(0).choice(len(pts), 400, replace=False)
    for i in sub:
        b = np.einsum("sij,sj->si", T[:, :2, :], pts[i] - T[:, 2, :])
        lam = np.concatenate([b, 1 - b.sum(1, keepdims=True)], axis=1)
        containing = np.nonzero(lam.min(1) >= -1e-12)[0]
        assert got[i] == containing.min(), (i, got[i], containing[:5])


# ---------------------------------------------------------------- TopK (pinned tie rule)
def test_topk_perm_against_torch_sort_and_topk():
    g = torch.Generator().manual_seed(0)
    for n, ratio in ((180, 0.1), (18, 0.1), (7, 0.5), (1, 0.1), (33, 0.5)):
        s = torch.randn(n, generator=g)
        batch = torch.zeros(n, dtype=torch.long)
        perm = gnn_ref.topk_perm(s, ratio, batch, 1)
        k = int(np.ceil(np.float32(ratio) * np.float32(n)))
        assert len(perm) == k
        assert torch.equal(perm, torch.topk(s, k).indices)                       # distinct scores: same order
    # ties (ReLU-dead rows all score tanh(0) = 0): pinned to the lower index, i.e. a stable descending sort
    s = torch.tensor([0.0, 0.5, 0.0, 0.5, 0.0, -1.0, 0.5])
    perm = gnn_ref.topk_perm(s, 0.7, torch.zeros(7, dtype=torch.long), 1)
    order = np.lexsort((np.arange(7), -s.numpy()))                               # independent: numpy lexsort
    assert perm.tolist() == order[:5].tolist() == [1, 3, 6, 0, 2]
    # two graphs in one batch: per-graph counts and offsets
    s = torch.randn(30, generator=g)
    batch = torch.cat([torch.zeros(12, dtype=torch.long), torch.ones(18, dtype=torch.long)])
    perm = gnn_ref.topk_perm(s, 0.25, batch, 2)
    want = torch.cat([torch.topk(s[:12], 3).indices, torch.topk(s[12:], 5).indices + 12])
    assert torch.equal(perm, want)


# ---------------------------------------------------------------- graph layers as dense matrices
def test_sage_and_gcn_against_dense_adjacency():
    g = torch.Generator().manual_seed(1)
    n, e, f, w = 23, 70, 5, 8
    x = torch.randn(n, f, generator=g, dtype=torch.float64)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei[:, :3] = torch.tensor([[4, 4, 4], [9, 9, 9]])          # duplicates are counted
    ei[:, 3] = torch.tensor([7, 7])                           # a self loop
    sage = gnn_ref.SAGEConv(f, w).double()
    gcn = gnn_ref.GCNConv(f, w).double()
    A = torch.zeros(n, n, dtype=torch.float64)
    for s_, d_ in ei.t().tolist():
        A[d_, s_] += 1.0
    deg = A.sum(1).clamp(min=1)
    want = (A @ x / deg[:, None]) @ sage.lin_l.weight.t() + sage.lin_l.bias + x @ sage.lin_r.weight.t()
    assert torch.allclose(sage(x, ei), want, atol=1e-12)
    A2 = A.clone()
    A2.fill_diagonal_(0.0)                                    # existing self loops dropped ...
    A2 += torch.eye(n, dtype=torch.float64)                   # ... one weight-1 loop per node added
    dis = A2.sum(1).pow(-0.5)
    want = (dis[:, None] * A2 * dis[None, :]) @ (x @ gcn.lin.weight.t()) + gcn.bias
    assert torch.allclose(gcn(x, ei), want, atol=1e-12)


def test_pooling_readout_against_loops():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(11, 4, generator=g)
    batch = torch.tensor([0] * 4 + [1] * 7)
    mx = gnn_ref.global_max_pool(x, batch, 2)
    mn = gnn_ref.global_mean_pool(x, batch, 2)
    assert torch.equal(mx[0], x[:4].max(0).values) and torch.equal(mx[1], x[4:].max(0).values)
    assert torch.allclose(mn[0], x[:4].mean(0)) and torch.allclose(mn[1], x[4:].mean(0))


# ---------------------------------------------------------------- polygon distance
def _poly_distance_numpy(pts, ring):
    """Independent formulation: clamp the projection parameter, distance to the closest point of each segment; 0 inside
    (crossing-number test)."""
    a = ring
    b = np.roll(ring, -1, axis=0)
    ab = b - a
    out = np.empty(len(pts))
    for i, p in enumerate(pts):
        t = np.clip(((p - a) * ab).sum(1) / (ab * ab).sum(1), 0.0, 1.0)
        d = np.linalg.norm(p - (a + t[:, None] * ab), axis=1).min()
        cross = ((a[:, 1] > p[1]) != (b[:, 1] > p[1])) & (p[0] < (b[:, 0] - a[:, 0]) * (p[1] - a[:, 1]) / (b[:, 1] - a[:, 1] + 1e-300) + a[:, 0])
        out[i] = 0.0 if cross.sum() % 2 else d
    return out


def test_polygon_distance_against_numpy_formulation():
    for short in ("ys930", "ah93w145"):
        coords, cells = load_mesh(short)
        topo = geom.Topology(cells, len(coords))
        rem = geom.removable_mask(coords, topo)
        x, y = coords[:, 0], coords[:, 1]
        ring = coords[(~rem) & (x > -0.5) & (x < 3.0) & (y > -0.5) & (y < 0.5)]   # the airfoil ring, vertex order (A.5)
        assert len(ring) in (120, 97)
        pts = coords[rem]
        got = geom.polygon_distance(pts, ring)
        want = _poly_distance_numpy(pts, ring)
        assert np.all(got > 0)                                    # removable vertices lie outside the airfoil
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, want.max())
        # ordering (what the environment uses): identical argsort
        assert np.array_equal(np.argsort(got, kind="stable"), np.argsort(want, kind="stable"))


# ---------------------------------------------------------------- drag / lift against an analytic field
def test_drag_lift_linear_velocity_field():
    """u = (a x + b y, c x + d y), p = 0: sigma is constant, so the closed-surface traction integral vanishes, and a
    constant pressure gradient p = g.x gives (drag, lift) = -area * g by the divergence theorem (outward normal of the
    FLUID domain points into the airfoil, hence the sign)."""
    coords, cells = load_mesh("ys930")
    topo = geom.Topology(cells, len(coords))
    tags = geom.facet_tags(coords, topo)
    pts2 = topo.p2_points(coords)
    U = np.stack([0.3 * pts2[:, 0] - 0.7 * pts2[:, 1], 1.1 * pts2[:, 0] + 0.2 * pts2[:, 1]], 1)[None]
    P = np.zeros((1, topo.nv))
    d, l = geom.drag_lift(coords, topo, tags, U, P, 1e-3)
    assert abs(d[0]) < 1e-12 and abs(l[0]) < 1e-12
    gx, gy = 0.8, -0.5
    P = (gx * coords[:, 0] + gy * coords[:, 1])[None]
    d, l = geom.drag_lift(coords, topo, tags, np.zeros_like(U), P, 1e-3)
    ring_pts = coords[np.unique(topo.edges[np.nonzero(tags == 1)[0]])]
    # polygon area of the airfoil from its ring (vertex order = curve order)
    x, y = coords[:, 0], coords[:, 1]
    rem = geom.removable_mask(coords, topo)
    ring = coords[(~rem) & (x > -0.5) & (x < 3.0) & (y > -0.5) & (y < 0.5)]
    area = 0.5 * abs(np.dot(ring[:, 0], np.roll(ring[:, 1], -1)) - np.dot(ring[:, 1], np.roll(ring[:, 0], -1)))
    assert len(ring_pts) == len(ring)
    # integral over the airfoil boundary of (-p n) with n pointing INTO the airfoil = +area * grad p
    assert abs(d[0] - area * gx) < 1e-10 and abs(l[0] - area * gy) < 1e-10
