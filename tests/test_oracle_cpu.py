"""CPU tests of the oracle itself: known facts of the shipped meshes (SURVEY.md probes), golden regression."""
import os

import numpy as np
import torch

from conftest import GOLDEN, lively_state_dict, load_mesh, make_config, oracle_fields
from oracle import geom, gnn_ref
from oracle.env_ref import Env2DAirfoilRef

# [probe] facts recorded in SURVEY.md / BASELINE.md for the two fixtures
FACTS = {
    "ys930": dict(nv=876, nc=1570, ne=2446, nb=182, airfoil_facets=120, removable=694),
    "ah93w145": dict(nv=797, nc=1431, ne=2228, nb=163, airfoil_facets=97, removable=634),
}


def test_mesh_facts():
    for short, f in FACTS.items():
        coords, cells = load_mesh(short)
        assert coords.shape == (f["nv"], 2) and cells.shape == (f["nc"], 3)
        topo = geom.Topology(cells, len(coords))
        assert topo.ne == f["ne"]
        assert len(topo.boundary_vertices) == f["nb"]
        xs = geom.smooth(coords, topo, 50)
        assert np.array_equal(xs[topo.on_boundary], coords[topo.on_boundary])  # boundary fixed
        tags = geom.facet_tags(xs, topo)
        assert int((tags == 1).sum()) == f["airfoil_facets"]
        assert int(geom.removable_mask(xs, topo).sum()) == f["removable"]
        # Euler: V - E + C = 1 - holes (one hole: the airfoil)
        assert f["nv"] - topo.ne + f["nc"] == 0


def test_param_census_and_state_dict_keys():
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    assert sum(p.numel() for p in net.parameters()) == 173493
    keys = set(net.state_dict())
    for k in ("conv1.lin_l.weight", "conv1.lin_l.bias", "conv1.lin_r.weight", "conv4.lin.weight", "conv4.bias",
              "pool1.weight", "lin3.bias"):
        assert k in keys


def test_topk_collapse():
    # 180 -> 18 -> 2 -> 1 -> 1 with ratio 0.1 in float32 (PyG's ceil(ratio * n.float()))
    n, seq = 180, []
    for _ in range(4):
        n = int((0.1 * torch.tensor([n]).to(torch.float)).ceil().long())
        seq.append(n)
    assert seq == [18, 2, 1, 1]


def test_polygon_distance_and_locate_small():
    ring = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    pts = np.array([[0.5, 0.5], [2.0, 0.5], [0.5, -1.0], [2.0, 2.0]])
    d = geom.polygon_distance(pts, ring)
    assert np.allclose(d, [0.0, 1.0, 1.0, np.sqrt(2.0)])
    coords = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    cells = np.array([[0, 1, 2], [1, 2, 3]], dtype=np.int32)
    c, nmiss, _ = geom.locate(np.array([[0.2, 0.2], [0.9, 0.9], [0.5, 0.5], [3.0, 3.0]]), coords, cells)
    assert c.tolist() == [0, 1, 0, 1] and nmiss == 1  # shared edge -> lowest index; outside -> closest cell


def test_p2_interpolation_reproduces_quadratics():
    coords, cells = load_mesh("ys930")
    topo = geom.Topology(cells, len(coords))
    pts2 = topo.p2_points(coords)
    f = lambda p: np.stack([1 + p[:, 0] ** 2 - p[:, 0] * p[:, 1], 2 * p[:, 1] ** 2 + p[:, 0]], 1)
    U = f(pts2)[None]
    P = (3 * coords[:, 0] - coords[:, 1])[None]
    rng = np.random.RandomState(0)
    cidx = rng.randint(0, len(cells), 500)
    w = rng.dirichlet([1, 1, 1], 500)
    q = (coords[topo.cells[cidx]] * w[:, :, None]).sum(1)
    cell_of, nmiss, _ = geom.locate(q, coords, topo.cells)
    assert nmiss == 0
    u, p = geom.eval_fields(q, len(q), cell_of, coords, topo, U, P)
    assert np.abs(u[0] - f(q)).max() < 1e-12
    assert np.abs(p[0] - (3 * q[:, 0] - q[:, 1])).max() < 1e-12


def test_drag_lift_hydrostatic():
    # u = 0, p = const: the closed airfoil surface integral of -p n vanishes
    coords, cells = load_mesh("ah93w145")
    topo = geom.Topology(cells, len(coords))
    tags = geom.facet_tags(coords, topo)
    U = np.zeros((1, topo.nv + topo.ne, 2))
    P = np.full((1, topo.nv), 2.5)
    d, l = geom.drag_lift(coords, topo, tags, U, P, 1e-3)
    assert abs(d[0]) < 1e-13 and abs(l[0]) < 1e-13


def test_golden_qnet_oracle():
    z = np.load(os.path.join(GOLDEN, "qnet_batch.npz"))
    from meshdqn_b200.data import Batch
    b = Batch(x=torch.from_numpy(z["x"]), edge_index=torch.from_numpy(z["edge_index"]))
    ptr = z["ptr"]
    b.batch = torch.repeat_interleave(torch.arange(len(ptr) - 1), torch.from_numpy(np.diff(ptr)))
    b.num_graphs = len(ptr) - 1
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")})
    q = net(b)
    assert np.allclose(q.detach().numpy(), z["q"], rtol=1e-5, atol=1e-9)
    assert np.allclose(q.sum(1).detach().numpy(), 1.0, atol=1e-5)


def test_golden_episode_oracle():
    for short in ("ys930", "ah93w145"):
        z = np.load(os.path.join(GOLDEN, f"episode_{short}.npz"))
        coords, cells, U, P = oracle_fields(short)
        cfg = make_config()
        cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
        env = Env2DAirfoilRef(cfg, mesh=(coords, cells))
        assert np.array_equal(env.flow_solver.coords, z["coords_smoothed"])
        assert np.array_equal(env.flow_solver.removable, z["removable"])
        torch.manual_seed(1370)
        net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
        net.set_num_nodes(17)
        net.load_state_dict(lively_state_dict(net))
        s = env.get_state()
        assert np.array_equal(s.x.numpy(), z["x0"]) and np.array_equal(s.edge_index.numpy(), z["edge_index0"])
        acts = []
        for i in range(len(z["actions"])):
            with torch.no_grad():
                a = int(net(s).argmax())
            s, r, done, _ = env.step(a)
            acts.append(a)
            assert abs(r - z["rewards"][i]) < 1e-9
            if done:
                break
        assert acts == z["actions"].tolist()
        assert done and bool(z["dones"][-1])


def test_golden_special_paths_oracle():
    """Scripted episodes through the branches the greedy policy never takes: do-nothing (action 180), a broken removal
    (strict interpolation -> code 2) and running out of vertices (tools/make_golden.py: SPECIAL)."""
    import hashlib
    sys_path_tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(sys_path_tools, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(GOLDEN, "special_ys930.npz"))
    for name, sp in mg.SPECIAL.items():
        coords, cells, U, P = oracle_fields("ys930")
        cfg = make_config(**sp["cfg"])
        cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
        env = Env2DAirfoilRef(cfg, mesh=(coords, cells))
        env.get_state()
        for i, a in enumerate(z[f"{name}/actions"]):
            s, r, done, _ = env.step(int(a))
            assert abs(r - z[f"{name}/rewards"][i]) < 1e-12 and done == bool(z[f"{name}/dones"][i])
            assert env.flow_solver.num_vertices == int(z[f"{name}/nvs"][i])
            assert mg.checksum(s.x.numpy()) == z[f"{name}/x_checksums"][i]
    assert bool(z["strict_break/dones"][-1]) and z["strict_break/rewards"][-1] == -1.0 and z["strict_break/nvs"][-1] == 876
    assert bool(z["out_of_vertices/dones"][-1]) and z["out_of_vertices/rewards"][-1] == -1.0
    assert list(z["do_nothing/offsets"]) == [0, 1, 2, 2, 3, 3, 3, 4, 4]
