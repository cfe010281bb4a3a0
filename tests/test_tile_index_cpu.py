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
