"""CPU tests of the host-side tiled point-location index (meshdqn_b200/tile_index.py) against the oracle's
brute-force locate: the candidate sets the tiled CUDA kernel walks must contain the lowest-index containing cell."""
import numpy as np

from conftest import load_mesh
from oracle import geom
from meshdqn_b200.synthetic import synthetic_airfoil_mesh, synthetic_fields
from meshdqn_b200.tile_index import build_tile_index, emulate_locate


def _check(coords, cells, leaf_cells, pts):
    topo = geom.Topology(cells, len(coords))
    U, P = synthetic_fields(coords, topo.edges, 2, 0)
    ti = build_tile_index(coords, topo.cells, topo.cell_edges, topo.ne, U, P, leaf_cells=leaf_cells)
    ref, nmiss, flag = geom.locate(pts, coords, topo.cells)
    got, lam, lc, leaf = emulate_locate(ti, pts)
    inside = flag == 0
    assert np.array_equal(got[inside], ref[inside])
    assert (got[~inside] == -1).all() and int((~inside).sum()) == nmiss
    # payloads: local coordinates / coefficients are copies of the global ones
    for L in range(ti.n_leaves):
        inf = ti.leaf_info[L]
        n = int(inf[12])
        gid = ti.gidL[inf[4]: inf[4] + n]
        assert np.all(np.diff(gid) > 0)                       # ascending global ids -> first hit = lowest index
        cv = ti.cvL[inf[4]: inf[4] + n].astype(np.int64)
        assert np.array_equal(ti.coordsL[inf[0] + cv[:, :3]], coords[topo.cells[gid]])
        assert np.array_equal(ti.UL[:, inf[2] + cv[:, :3]], U[:, topo.cells[gid]])
        assert np.array_equal(ti.UL[:, inf[2] + cv[:, 3:]], U[:, len(coords) + topo.cell_edges[gid]])
        assert np.array_equal(ti.PL[:, inf[0] + cv[:, :3]], P[:, topo.cells[gid]])
    # every section offset / size is a 16-byte multiple (bulk-copy requirement)
    inf = ti.leaf_info
    assert not (inf[:, 1] % 2).any() and not (inf[:, 5] % 4).any() and not (inf[:, 7] % 8).any() and not (inf[:, 9] % 8).any()
    assert not (inf[:, 0] % 2).any() and not (inf[:, 4] % 4).any() and not (inf[:, 6] % 8).any() and not (inf[:, 8] % 8).any()
    return ti


def test_tile_index_fixture_mesh():
    coords, cells = load_mesh("ys930")
    topo = geom.Topology(cells, len(coords))
    xs = geom.smooth(coords, topo, 50)
    rng = np.random.RandomState(0)
    pts = np.concatenate([topo.p2_points(xs), np.stack([rng.uniform(-0.5, 3, 500), rng.uniform(-0.5, 0.5, 500)], 1)])
    ti = _check(xs, cells, 64, pts)
    assert ti.n_leaves == 32
    ti1 = _check(xs, cells, 4096, pts[:300])                  # single leaf (depth 0)
    assert ti1.n_leaves == 1 and ti1.depth == 0


def test_tile_index_synthetic_mesh_points_on_split_planes():
    coords, cells, _ = synthetic_airfoil_mesh(6000, seed=3)
    topo = geom.Topology(cells, len(coords))
    U, P = synthetic_fields(coords, topo.edges, 1, 0)
    ti = build_tile_index(coords, topo.cells, topo.cell_edges, topo.ne, U, P, leaf_cells=128)
    rng = np.random.RandomState(1)
    # points exactly on split planes and on leaf corners, plus the mesh's own P2 points
    on_planes = []
    for node in range(ti.n_leaves - 1):
        s = ti.tree[node]
        d = int(s.view(np.int64) & 1)
        q = np.stack([rng.uniform(-0.5, 3, 4), rng.uniform(-0.5, 0.5, 4)], 1)
        q[:, d] = s
        on_planes.append(q)
    pts = np.concatenate(on_planes + [topo.p2_points(coords)[::7]])
    _check(coords, cells, 128, pts)


def test_tri_rect_overlap_is_conservative_and_tighter_than_bounding_boxes():
    """The separating-axis filter may keep a pair that does not overlap, never drop one that does."""
    from meshdqn_b200.tile_index import tri_rect_overlap
    rng = np.random.RandomState(7)
    n = 20000
    tri = rng.uniform(0, 1, (n, 3, 2))
    lo = rng.uniform(0, 1, (n, 2))
    hi = lo + rng.uniform(0.01, 0.3, (n, 2))
    keep = tri_rect_overlap(tri, lo, hi, 1e-9)
    # ground truth by dense sampling: a rectangle point inside the triangle, or a triangle point inside the rectangle
    g = np.linspace(0, 1, 9)
    gx, gy = np.meshgrid(g, g)
    pts = lo[:, None, :] + np.stack([gx.ravel(), gy.ravel()], 1)[None] * (hi - lo)[:, None, :]       # [n, 81, 2]
    a, b, c = tri[:, 0][:, None], tri[:, 1][:, None], tri[:, 2][:, None]
    d = (b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (c[..., 0] - a[..., 0]) * (b[..., 1] - a[..., 1])
    l1 = ((pts[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (c[..., 0] - a[..., 0]) * (pts[..., 1] - a[..., 1])) / d
    l2 = ((b[..., 0] - a[..., 0]) * (pts[..., 1] - a[..., 1]) - (pts[..., 0] - a[..., 0]) * (b[..., 1] - a[..., 1])) / d
    inside_tri = ((l1 >= 0) & (l2 >= 0) & (1 - l1 - l2 >= 0)).any(1)
    w = rng.dirichlet([1, 1, 1], (n, 40))                                                              # triangle samples
    tp = (w[..., None] * tri[:, None, :, :]).sum(2)
    inside_rect = ((tp >= lo[:, None]) & (tp <= hi[:, None])).all(2).any(1)
    overlap = inside_tri | inside_rect
    assert not (overlap & ~keep).any()                       # conservative
    bbox = (tri.min(1) <= hi).all(1) & (tri.max(1) >= lo).all(1)
    assert not (keep & ~bbox).any() and keep.sum() < 0.9 * bbox.sum()   # and strictly tighter than the bbox test


def test_tile_index_shared_memory_sizing():
    from meshdqn_b200.tile_index import pick_smem
    coords, cells, _ = synthetic_airfoil_mesh(6000, seed=3)
    topo = geom.Topology(cells, len(coords))
    U, P = synthetic_fields(coords, topo.edges, 5, 0)
    ti = build_tile_index(coords, topo.cells, topo.cell_edges, topo.ne, U, P, leaf_cells=128)
    inf, T = ti.leaf_info, 5
    need = 16 + 16 * inf[:, 1] + T * (16 * inf[:, 3] + 8 * inf[:, 1]) + 16 * inf[:, 5] + 2 * inf[:, 7] + 2 * inf[:, 9]
    assert np.array_equal(need, ti.leaf_bytes)               # csrc/interp_tiled.cuh: tile_leaf_bytes
    assert 0 < ti.smem_bytes <= 227 * 1024 and (ti.leaf_bytes <= ti.smem_bytes).mean() >= 0.98
    assert ti.smem_bytes == ti.leaf_bytes[ti.leaf_bytes <= ti.smem_bytes].max() and ti.ctas_per_sm in (1, 2, 3, 4)
    b = np.array([1000] * 98 + [200000, 300000])
    assert pick_smem(b, 100.0) == (1000, 4) and pick_smem(b, 400.0) == (1000, 2)
    assert pick_smem(np.array([300000] * 10), 100.0) is None
