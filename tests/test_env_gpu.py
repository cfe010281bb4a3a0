"""GPU parity: the per-action environment step (through the C-ABI) vs the CPU oracle and golden episodes.

Bars (BASELINE.json north_star): point-location cell indices, removable masks and the chosen action
bit-exact; interpolated fields and drag/lift within 1e-10 relative in float64.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, lively_state_dict, load_mesh, make_config, oracle_fields
from oracle import geom, gnn_ref
from oracle.env_ref import Env2DAirfoilRef

pytestmark = pytest.mark.gpu


def checksum(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.int64)[0]


def make_envs(short, dev, **kw):
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    coords, cells, U, P = oracle_fields(short)
    cfg = make_config(**kw)
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    return Env2DAirfoil(cfg, mesh=(coords, cells), device=dev), Env2DAirfoilRef(cfg, mesh=(coords, cells))


@pytest.mark.parametrize("short", ["ys930", "ah93w145"])
def test_mesh_services_bit_exact(cuda_device, short):
    env, renv = make_envs(short, cuda_device)
    m, rt = env.flow_solver.mesh, renv.flow_solver.topo
    assert (m.ne, m.nb) == (rt.ne, len(rt.boundary_vertices))
    assert np.array_equal(m.edges.cpu().numpy(), rt.edges)
    assert np.array_equal(m.cell_edges.cpu().numpy(), rt.cell_edges)
    assert np.array_equal(m.nbr_ptr.cpu().numpy(), rt.nbr_ptr) and np.array_equal(m.nbr_idx[: 2 * m.ne].cpu().numpy(), rt.nbr_idx)
    assert np.array_equal(m.vc_idx.cpu().numpy(), rt.vc_idx)
    assert np.array_equal(m.boundary_vertices(), rt.boundary_vertices)
    assert np.array_equal(m.coordinates(), renv.flow_solver.coords)          # 50 Gauss-Seidel sweeps, bit-identical
    assert np.array_equal(env.flow_solver.tags.cpu().numpy(), renv.flow_solver.tags)
    assert np.array_equal(env.flow_solver.removable, renv.flow_solver.removable)
    assert np.array_equal(env.distance_lookup, renv.distance_lookup)
    assert np.abs(env.gt_drag / renv.gt_drag - 1).max() < 1e-10 and np.abs(env.gt_lift / renv.gt_lift - 1).max() < 1e-10
    z = np.load(os.path.join(GOLDEN, f"episode_{short}.npz"))
    assert np.array_equal(m.coordinates(), z["coords_smoothed"]) and np.array_equal(env.flow_solver.removable, z["removable"])
    assert np.array_equal(env.flow_solver.tags.cpu().numpy(), z["tags"])
    s = env.get_state()
    assert np.array_equal(s.x.cpu().numpy(), z["x0"]) and np.array_equal(s.edge_index.cpu().numpy(), z["edge_index0"])
    renv.get_state()
    assert env.coord_map == renv.coord_map and env.inv_coord_map == renv.inv_coord_map
    assert np.array_equal(env.n_closest, renv.n_closest)


@pytest.mark.parametrize("short", ["ys930", "ah93w145"])
def test_greedy_episode_matches_oracle_and_golden(cuda_device, short):
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    env, renv = make_envs(short, cuda_device)
    z = np.load(os.path.join(GOLDEN, f"episode_{short}.npz"))
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    ref.load_state_dict(lively_state_dict(ref))
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda_device)
    s, rs = env.get_state(), renv.get_state()
    acts = []
    for i in range(len(z["actions"])):
        am, q = net.select_action(s)
        a = int(am[0])
        with torch.no_grad():
            ra = int(ref(rs).argmax())
        assert a == ra == int(z["actions"][i]), f"step {i}"
        s, r, done, _ = env.step(a)
        rs, rr, rdone, _ = renv.step(ra)
        acts.append(a)
        assert done == rdone == bool(z["dones"][i])
        assert abs(r - rr) < 1e-9 and abs(r - z["rewards"][i]) < 1e-9
        if a != env.action_space.n:
            cell_of = env.last["cell_of"].cpu().numpy()
            assert np.array_equal(cell_of, renv.last["cell_of"])                       # bit-exact point location
            assert checksum(cell_of) == z["cell_checksums"][i]
            assert int(env.last["miss"]) == renv.last["nmiss"]
            U, rU = env.U.cpu().numpy(), renv.U
            assert np.abs(U - rU).max() <= 1e-10 * np.abs(rU).max()
            assert np.abs(env.P.cpu().numpy() - renv.P).max() <= 1e-10 * np.abs(renv.P).max()
        assert np.array_equal(env.flow_solver.removable, renv.flow_solver.removable)
        assert np.abs(env.new_drags / renv.new_drags - 1).max() < 1e-10
        assert np.abs(env.new_lifts / renv.new_lifts - 1).max() < 1e-10
        assert np.abs(env.new_drags / z["drags"][i] - 1).max() < 1e-10
        assert torch.equal(s.x.cpu(), rs.x) and torch.equal(s.edge_index.cpu(), rs.edge_index)
        assert env.flow_solver.mesh.nv == z["nvs"][i]
        if done:
            break
    assert done and len(acts) == len(z["actions"])
    assert np.array_equal(s.x.cpu().numpy(), z["x_last"])


def test_do_nothing_action_and_attributes(cuda_device):
    env, renv = make_envs("ys930", cuda_device)
    n = env.action_space.n
    env.get_state()
    assert n == 180 and env.N_CLOSEST == 180 and len(env.coord_map) == 180
    s, r, done, info = env.step(n)          # do nothing: window shifts by one (quirk B5), reward recomputed
    rs, rr, rdone, _ = renv.step(n)
    assert env.do_nothing_offset == 1 and info == {} and not done
    assert abs(r - rr) < 1e-12 and abs(r - 1.0) < 1e-9   # unchanged mesh: drag error 0 -> reward 2*exp(0)-1
    assert torch.equal(s.x.cpu(), rs.x) and torch.equal(s.edge_index.cpu(), rs.edge_index)
    assert env.velocities.shape == (5, 876, 2) and env.pressures.shape == (5, 876, 1)
    assert np.array_equal(env.velocities, renv.velocities) and np.array_equal(env.pressures, renv.pressures)
    assert env.return_vals()[0] is env.gt_drag
    d = float(env.flow_solver.drag_probe.sample(env.U[0], env.P[0]))
    assert abs(d / renv.gt_drag[0] - 1) < 1e-10


def test_interpolation_on_synthetic_mesh_with_random_removals(cuda_device):
    """Size-independent properties on a larger synthetic mesh: P2 interpolation reproduces quadratics exactly,
    cell indices equal the oracle's brute-force search, repeated runs are bit-identical."""
    from meshdqn_b200.Env2DAirfoil import SourceField
    from meshdqn_b200.flow_solver import DeviceMesh
    from meshdqn_b200.synthetic import synthetic_airfoil_mesh
    from scipy.spatial import Delaunay
    coords, cells, _ = synthetic_airfoil_mesh(20000, seed=1)
    topo0 = geom.Topology(cells, len(coords))
    pts2 = topo0.p2_points(coords)
    f = lambda p: np.stack([1 + p[:, 0] ** 2 - p[:, 0] * p[:, 1], 2 * p[:, 1] ** 2 + p[:, 0]], 1)
    U0 = np.stack([f(pts2), -f(pts2)])
    P0 = np.stack([3 * coords[:, 0] - coords[:, 1], coords[:, 0] + 1])
    m0 = DeviceMesh(coords, cells, cuda_device)
    src = SourceField(m0, U0, P0)
    rng = np.random.RandomState(0)
    interior = np.nonzero(~topo0.on_boundary)[0]
    drop = rng.choice(interior, 200, replace=False)
    keep = np.ones(len(coords), bool)
    keep[drop] = False
    c2 = coords[keep]
    newid = np.cumsum(keep) - 1
    isb = np.zeros(len(c2), bool)
    isb[newid[topo0.boundary_vertices]] = True
    t2 = Delaunay(c2).simplices
    t2 = t2[isb[t2].sum(1) != 3]
    m1 = DeviceMesh(c2, t2, cuda_device)
    U, P, cell_of, miss = src.interpolate(m1)
    topo1 = geom.Topology(t2, len(c2))
    pts = topo1.p2_points(c2)
    ref_cells, nmiss, _ = geom.locate(pts, coords, topo0.cells)
    assert np.array_equal(cell_of.cpu().numpy(), ref_cells) and int(miss) == nmiss
    inside = np.ones(len(pts), bool)
    if nmiss:
        inside = geom.locate(pts, coords, topo0.cells)[2] == 0
    Un = U.cpu().numpy()
    assert np.abs(Un[0][inside] - f(pts)[inside]).max() < 1e-11      # exact for quadratics
    assert np.abs(P.cpu().numpy()[0] - (3 * c2[:, 0] - c2[:, 1]))[inside[: len(c2)]].max() < 1e-11
    assert np.array_equal(Un[1], -Un[0])                              # linearity, bit-exact
    U2, P2, cell2, _ = src.interpolate(m1)
    assert torch.equal(U2, U) and torch.equal(P2, P) and torch.equal(cell2, cell_of)


def _coarsened(coords, cells_topo, n_drop, seed, dev):
    """Target mesh for interpolation tests: drop interior vertices, re-triangulate, apply the reference's cell rule."""
    from meshdqn_b200.flow_solver import DeviceMesh
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(seed)
    interior = np.nonzero(~cells_topo.on_boundary)[0]
    keep = np.ones(len(coords), bool)
    keep[rng.choice(interior, n_drop, replace=False)] = False
    c2 = coords[keep]
    newid = np.cumsum(keep) - 1
    isb = np.zeros(len(c2), bool)
    isb[newid[cells_topo.boundary_vertices]] = True
    t2 = Delaunay(c2).simplices
    t2 = t2[isb[t2].sum(1) != 3]
    return c2, t2, DeviceMesh(c2, t2, dev)


@pytest.mark.parametrize("case", ["ys930_leaf64", "synthetic_leaf256", "synthetic_leaf128_overflow", "synthetic_leaf128_hbm_leaves",
                                  "synthetic_oversized_leaf"])
def test_tiled_interpolation_bit_identical_to_grid_path_and_oracle(cuda_device, case):
    """The tiled (k-d leaf, TMA-staged) kernel must return the same cell ids and the same field bits as the
    uniform-grid kernel, and the oracle's brute-force cell ids."""
    from meshdqn_b200.Env2DAirfoil import SourceField
    from meshdqn_b200.flow_solver import DeviceMesh
    from meshdqn_b200.synthetic import synthetic_airfoil_mesh, synthetic_fields
    if case == "ys930_leaf64":
        coords, cells = load_mesh("ys930")
        topo0 = geom.Topology(cells, len(coords))
        coords = geom.smooth(coords, topo0, 50)
        leaf, ndrop = 64, 12
    else:
        coords, cells, _ = synthetic_airfoil_mesh(20000, seed=2)
        topo0 = geom.Topology(cells, len(coords))
        leaf, ndrop = {"synthetic_leaf256": (256, 150), "synthetic_leaf128_overflow": (128, 150),
                       "synthetic_leaf128_hbm_leaves": (128, 150)}.get(case, (8192, 150))
    U0, P0 = synthetic_fields(coords, topo0.edges, 5, 1)
    m0 = DeviceMesh(coords, cells, cuda_device)
    if case == "synthetic_oversized_leaf":
        with pytest.raises(RuntimeError, match="shared memory"):       # 5k-cell leaves exceed 227 KB: loud failure
            SourceField(m0, U0, P0, tiled=True, leaf_cells=leaf).interpolate(m0)
        return
    grid = SourceField(m0, U0, P0, tiled=False)
    # bucket_factor=0: 32-slot buckets, so most points take the overflow path (served from HBM, same bits)
    tiled = SourceField(m0, U0, P0, tiled=True, leaf_cells=leaf, bucket_factor=0 if case.endswith("overflow") else 2)
    assert tiled.tile is not None and grid.tile is None
    if case.endswith("hbm_leaves"):
        # shrink the per-CTA shared memory to the median leaf: half of the leaves no longer fit and are served by
        # their CTA straight from the leaf arrays in HBM (the route the rare oversize leaf of a graded mesh takes)
        tiled.tile.smem_bytes = int(np.median(tiled.tile_host.leaf_bytes))
        assert (tiled.tile_host.leaf_bytes > tiled.tile.smem_bytes).sum() >= 2
    for seed in (0, 1):
        c2, t2, m1 = _coarsened(coords, topo0, ndrop, seed, cuda_device)
        if case == "ys930_leaf64" and seed == 1:
            m1.smooth(50)                                               # moved vertices: generic query points
            c2 = m1.coordinates()
        Ug, Pg, cg, mg = grid.interpolate(m1)
        Ut, Pt, ct, mt = tiled.interpolate(m1)
        assert torch.equal(ct, cg) and int(mt) == int(mg)
        assert torch.equal(Ut, Ug) and torch.equal(Pt, Pg)
        assert int(tiled._tile_counters.abs().sum()) == 0                # counters left zeroed for the next call
        topo1 = geom.Topology(t2, len(c2))
        ref_cells, nmiss, _ = geom.locate(topo1.p2_points(c2), coords, topo0.cells)
        assert np.array_equal(ct.cpu().numpy(), ref_cells) and int(mt) == nmiss
    # points outside the mesh (closest-cell fallback) go through the same miss path
    far = DeviceMesh(np.array([[-0.6, -0.6], [3.2, 0.0], [0.5, 0.7], [1.0, 0.0]]), np.array([[0, 1, 2], [0, 1, 3]]), cuda_device)
    Ug, Pg, cg, mg = grid.interpolate(far)
    Ut, Pt, ct, mt = tiled.interpolate(far)
    assert torch.equal(ct, cg) and int(mt) == int(mg) and int(mt) >= 3
    assert torch.equal(Ut, Ug) and torch.equal(Pt, Pg)


def test_snapshot_files_round_trip(cuda_device, tmp_path):
    """set_plot_dir writes the reference's four snapshot files (Env2DAirfoil.py:432-449); an environment built from
    the saved P2/P1 arrays (the ':126-133' load branch) reproduces the fields, the ground-truth drag and the state."""
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    env = Env2DAirfoil(cfg, mesh=(coords, cells), device=cuda_device)
    d = str(tmp_path / "plots")
    env.set_plot_dir(d)
    v, pr = np.load(d + "/snapshots/velocities.npy"), np.load(d + "/snapshots/pressures.npy")
    assert v.shape == (U.shape[0], len(coords), 2) and pr.shape == (U.shape[0], len(coords), 1)
    assert np.array_equal(v, env.velocities) and np.array_equal(pr, env.pressures)
    cfg2 = make_config()
    cfg2["agent_params"]["u"] = d + "/snapshots/save_velocities.npy"
    cfg2["agent_params"]["p"] = d + "/snapshots/save_pressures.npy"
    env2 = Env2DAirfoil(cfg2, mesh=(coords, cells), device=cuda_device)
    assert torch.equal(env2.U, env.U) and torch.equal(env2.P, env.P) and np.array_equal(env2.gt_drag, env.gt_drag)
    s1, s2 = env.get_state(), env2.get_state()
    assert torch.equal(s1.x, s2.x) and torch.equal(s1.edge_index, s2.edge_index)


def test_dolfin_ordered_snapshot_files(cuda_device, tmp_path):
    """Snapshot files in DOLFIN's dof order (what the reference's FEniCS run writes, Env2DAirfoil.py:139-150) plus the
    dof coordinates DOLFIN reports: the environment reproduces the one built from the package-layout arrays, and
    set_plot_dir writes the files back in the order they came in."""
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    from meshdqn_b200.snapshots import p2_points
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    env = Env2DAirfoil(cfg, mesh=(coords, cells), device=cuda_device)
    m = env.flow_solver.mesh
    pts = p2_points(m.coordinates(), m.edges.cpu().numpy())         # dof points of the smoothed mesh the "solve" ran on
    rng = np.random.RandomState(5)
    order, p_order = rng.permutation(len(pts)), rng.permutation(m.nv)
    np.savez(tmp_path / "dofmap.npz", xy_u=np.repeat(pts[order], 2, axis=0), comp_u=np.tile([0, 1], len(pts)),
             xy_p=m.coordinates()[p_order])
    np.save(tmp_path / "save_velocities.npy", U[:, order, :].reshape(U.shape[0], -1))
    np.save(tmp_path / "save_pressures.npy", P[:, p_order])
    cfg2 = make_config()
    cfg2["agent_params"].update(u=str(tmp_path / "save_velocities.npy"), p=str(tmp_path / "save_pressures.npy"),
                                dof_map=str(tmp_path / "dofmap.npz"))
    env2 = Env2DAirfoil(cfg2, mesh=(coords, cells), device=cuda_device)
    assert torch.equal(env2.U, env.U) and torch.equal(env2.P, env.P) and np.array_equal(env2.gt_drag, env.gt_drag)
    s1, s2 = env.get_state(), env2.get_state()
    assert torch.equal(s1.x, s2.x) and torch.equal(s1.edge_index, s2.edge_index)
    env2.set_plot_dir(str(tmp_path / "plots"))
    assert np.array_equal(np.load(tmp_path / "plots/snapshots/save_velocities.npy"), np.load(tmp_path / "save_velocities.npy"))
    assert np.array_equal(np.load(tmp_path / "plots/snapshots/save_pressures.npy"), np.load(tmp_path / "save_pressures.npy"))


@pytest.mark.parametrize("name", ["do_nothing", "strict_break", "out_of_vertices"])
def test_special_paths_match_oracle_and_golden(cuda_device, name):
    """Do-nothing steps (action 180: growing closest-window offset, reward recomputed on the unchanged mesh), a removal
    that breaks (strict interpolation: code 2 -> reward -1, terminal, the old mesh restored) and running out of vertices
    (Env2DAirfoil.py:330-364, 456-458, 569-573): the device environment against the oracle step by step and against the
    committed golden file."""
    z = np.load(os.path.join(GOLDEN, "special_ys930.npz"))
    cfgs = {"do_nothing": {}, "strict_break": dict(interp_strict_tol=-1.0), "out_of_vertices": dict(N_closest=700)}
    env, renv = make_envs("ys930", cuda_device, **cfgs[name])
    s, rs = env.get_state(), renv.get_state()
    assert torch.equal(s.x.cpu(), rs.x) and torch.equal(s.edge_index.cpu(), rs.edge_index)
    for i, a in enumerate(z[f"{name}/actions"]):
        s, r, done, _ = env.step(int(a))
        rs, rr, rdone, _ = renv.step(int(a))
        assert done == rdone == bool(z[f"{name}/dones"][i])
        assert abs(r - rr) < 1e-9 and abs(r - z[f"{name}/rewards"][i]) < 1e-9
        assert env.flow_solver.mesh.nv == renv.flow_solver.num_vertices == int(z[f"{name}/nvs"][i])
        assert env.do_nothing_offset == renv.do_nothing_offset == int(z[f"{name}/offsets"][i])
        assert torch.equal(s.x.cpu(), rs.x) and torch.equal(s.edge_index.cpu(), rs.edge_index)
        assert checksum(s.x.cpu().numpy()) == z[f"{name}/x_checksums"][i]
    if name == "strict_break":       # the broken removal left the environment on the old mesh
        assert np.array_equal(env.flow_solver.mesh.coordinates(), renv.flow_solver.coords)
        assert np.array_equal(env.flow_solver.removable, renv.flow_solver.removable)


def test_fast_div_sqrt_match_operators(cuda_device):
    """The smoothing sweep's branch-free division / square root (geom.cu: fast_div, fast_sqrt) are bit-identical to the
    compiler's ``/`` and ``sqrt()`` wherever they accept their operands -- 2^28 pairs of arbitrary finite bit patterns and
    2^28 pairs in the magnitude range of mesh coordinates -- and they accept essentially all of the latter."""
    from meshdqn_b200 import _lib
    L = _lib.lib()
    n = 1 << 28
    for mode in (0, 1):
        counts = torch.zeros(4, dtype=torch.int64, device=cuda_device)
        with torch.cuda.device(cuda_device):
            _lib.check(L.mdq_debug_fast_math_check(12345 + mode, n, mode, _lib.ptr(counts), _lib.stream_ptr()))
        bad_div, bad_sqrt, decl_div, decl_sqrt = counts.cpu().tolist()
        assert bad_div == 0 and bad_sqrt == 0, (mode, bad_div, bad_sqrt)
        if mode == 1:
            assert decl_div == 0 and decl_sqrt == 0, (decl_div, decl_sqrt)
        else:
            assert decl_div < n and decl_sqrt < n


def test_fused_interpolation_drag_lift_equals_separate_launches(cuda_device):
    """mdq_interpolate_drag_lift (the miss pass's last block integrates the airfoil facets) against mdq_interpolate followed by
    mdq_drag_lift: identical fields, bit-identical drag / lift (one reduction shape), and the environment's reward uses it."""
    from meshdqn_b200.probes import drag_lift_device
    env, renv = make_envs("ys930", cuda_device)
    env.get_state()
    for a in (5, 40):
        env.step(a)
    fs = env.flow_solver
    U1, P1, c1, m1 = env.source.interpolate(fs.mesh)                     # separate launches
    fs.__dict__.pop("_dl_cache", None)
    dl_sep = drag_lift_device(fs, U1, P1).clone()
    U2, P2, c2, m2 = env.source.interpolate(fs.mesh, probes_of=fs)       # fused
    cached = fs._dl_cache
    assert cached[1] is U2 and cached[2] is P2
    dl_fused = drag_lift_device(fs, U2, P2)
    assert dl_fused is cached[3]
    assert torch.equal(U1, U2) and torch.equal(P1, P2) and torch.equal(c1, c2)
    assert torch.equal(dl_sep, dl_fused)
    assert torch.isfinite(dl_fused).all() and dl_fused.abs().max() > 0
