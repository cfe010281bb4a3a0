"""CPU tests of the host logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_mesh
from meshdqn_b200 import _lib
from meshdqn_b200.data import Batch, Data, DataLoader


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "meshdqn_b200.h")).read()
    declared = set(re.findall(r"\b(mdq_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = ctypes.CDLL(_lib.build())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/meshdqn_b200.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert _lib.lib().mdq_version() >= 100


def test_batch_collation_matches_pyg_rule():
    g = torch.Generator().manual_seed(0)
    ds = [Data(x=torch.randn(n, 3, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))
          for n, e in ((4, 5), (2, 0), (7, 9))]
    b = Batch.from_data_list(ds)
    assert b.x.shape == (13, 3) and b.edge_index.shape == (2, 14) and b.num_graphs == 3
    assert b.batch.tolist() == [0] * 4 + [1] * 2 + [2] * 7
    assert b.ptr.tolist() == [0, 4, 6, 13] and b.eptr.tolist() == [0, 5, 5, 14]
    assert torch.equal(b.edge_index[:, 5:], ds[2].edge_index + 6)
    loader = DataLoader(ds, batch_size=2)
    sizes = [bb.num_graphs for bb in loader]
    assert sizes == [2, 1] and len(loader) == 2


def test_flat_parameter_buffer_aliases_state_dict():
    from meshdqn_b200.airfoilgcnn import AirfoilGCNN, NodeRemovalNet
    from oracle import gnn_ref
    torch.manual_seed(1370)
    net = NodeRemovalNet(181, conv_width=128, topk=0.1)
    net.set_num_nodes(17)
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    assert set(net.state_dict()) == set(ref.state_dict())
    assert sum(p.numel() for p in net.parameters()) == 173493
    net.load_state_dict(ref.state_dict())
    net._pack()
    assert net._n_used == 123832 and net._net.n_params >= 173493  # 123,829 used params + alignment padding
    sd = net.state_dict()
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k], v), k
    # the kernel layout is the transpose: conv2.lin_l.weight^T occupies rows 0..127 of a [256][128] block
    off = net._views["conv2.lin_l.weight"][0]
    assert torch.equal(net._flat[off:off + 128 * 128].view(128, 128).t(), ref.conv2.lin_l.weight.detach())
    # writes through the parameter land in the flat buffer
    with torch.no_grad():
        net.lin3.bias.fill_(0.25)
    boff = net._views["lin3.bias"][0]
    assert torch.all(net._flat[boff:boff + 181] == 0.25)
    a = AirfoilGCNN(64)
    a._pack()
    assert a._net.n_blocks == 6 and a._net.in_col0 == 2 and a._net.softmax == 0


def test_fused_kernel_shared_memory_budget():
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net._pack()
    L = _lib.lib()
    fwd = L.mdq_qnet_smem_bytes(net._net, 180, 512, 1, 0)
    bwd = L.mdq_qnet_smem_bytes(net._net, 180, 512, 1, 1)
    assert 0 < fwd <= 227 * 1024 and 0 < bwd <= 227 * 1024
    assert L.mdq_qnet_smem_bytes(net._net, 180, 512, 2, 1) == -1  # backward is one graph per CTA
    assert L.mdq_qnet_bwd_workspace_floats(net._net, 256, 180) > 0


def test_no_cpu_path():
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    d = Data(x=torch.zeros(4, 17), edge_index=torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(d)
    if not torch.cuda.is_available():
        from meshdqn_b200.flow_solver import FlowSolver
        with pytest.raises(RuntimeError):
            FlowSolver({"mu": 1e-3}, {"mesh": None}, {"smooth": True}, mesh=load_mesh("ys930"))


def test_replay_batch_collation_and_lr_schedule():
    from meshdqn_b200.replay import ReplayBatch, multistep_lr
    g = torch.Generator().manual_seed(1)
    mk = lambda: Data(x=torch.randn(5, 17, generator=g), edge_index=torch.randint(0, 5, (2, 6), generator=g))
    tr = [(mk(), 3, mk(), 0.5), (mk(), 180, None, -1.0), (mk(), 7, mk(), 0.25)]
    rb = ReplayBatch.from_transitions(tr)
    assert rb.actions.tolist() == [3, 180, 7] and rb.next_slot.tolist() == [0, -1, 1]
    assert rb.states.num_graphs == 3 and rb.next_states.num_graphs == 2
    assert rb.h2d_bytes() > 0
    assert multistep_lr(1e-5, 0) == 1e-5 and abs(multistep_lr(1e-5, 500000) - 1e-6) < 1e-18
    assert abs(multistep_lr(1e-5, 1500000) - 1e-8) < 1e-20


def test_xdmf_reader_on_reference_files():
    ref = "/root/reference/xdmf_files/ys930_0.15000_triangle.xdmf"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    from meshdqn_b200.xdmf import read_xdmf_mesh
    coords, cells = read_xdmf_mesh(ref)
    gc, gcells = load_mesh("ys930")
    assert np.array_equal(coords, gc) and np.array_equal(cells, gcells)


def test_synthetic_mesh_generator():
    from meshdqn_b200.synthetic import synthetic_airfoil_mesh
    from oracle import geom
    coords, cells, n_ring = synthetic_airfoil_mesh(4000, seed=0)
    assert 2500 < len(cells) < 5000
    topo = geom.Topology(cells, len(coords))
    tags = geom.facet_tags(coords, topo)
    # the hole boundary is the airfoil ring, up to the few all-ring-vertex cells the reference's rule
    # (Env2DAirfoil.py:496) also drops just outside a concave stretch of the contour
    assert abs(int((tags == 1).sum()) - n_ring) <= 3
    assert len(coords) - topo.ne + len(cells) == 0


def test_replay_batch_arena_packing_round_trip(monkeypatch):
    """ReplayBatch.pin_memory packs the whole minibatch into one arena; the views must equal the collated tensors
    and survive `.to()` (one copy) with the kernel-facing metadata intact."""
    import torch
    from meshdqn_b200.data import Data
    from meshdqn_b200.replay import ReplayBatch
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)      # no CUDA driver in the CPU suite
    g = torch.Generator().manual_seed(5)
    mk = lambda n, e: Data(x=torch.randn(n, 17, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))
    tr = [(mk(180, 372), 3, mk(179, 370), 0.5), (mk(150, 300), 180, None, -1.0), (mk(7, 0), 0, mk(6, 2), 0.25)]
    rb = ReplayBatch.from_transitions(tr)
    rp = rb.pin_memory()
    assert rp._arena is not None and rp._arena.numel() % 256 == 0 and rp.h2d_bytes() == rp._arena.numel()
    for rr in (rp, rp.to("cpu")):
        for k in ("actions", "next_slot", "rewards", "owner"):
            assert torch.equal(getattr(rr, k), getattr(rb, k))
        for name in ("states", "next_states"):
            a, b = getattr(rr, name), getattr(rb, name)
            for k in ("x", "edge_index", "batch", "ptr", "eptr"):
                assert torch.equal(getattr(a, k), getattr(b, k)) and getattr(a, k).dtype == getattr(b, k).dtype
            m, m0 = a._host_meta(), b._host_meta()
            assert torch.equal(m[0], m0[0]) and torch.equal(m[1], m0[1]) and m[2:] == m0[2:]
            assert a.edge_index.data_ptr() % 16 == 0 and a.x.data_ptr() % 16 == 0
        assert rr.tensors()[0] is rr._arena and len(rr.tensors()) == 1


def test_env_replicas_host_logic_on_cpu():
    """run_env_replicas with the CPU oracle environment: every replica takes its steps, episodes restart on `done`."""
    import io, contextlib
    import torch
    from conftest import make_config, oracle_fields
    from meshdqn_b200.parallel import run_env_replicas
    from oracle.env_ref import Env2DAirfoilRef
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    made = []

    def mk():
        e = Env2DAirfoilRef(cfg, mesh=(coords, cells))
        made.append(e)
        return e
    with contextlib.redirect_stdout(io.StringIO()):
        tot, wall = run_env_replicas(mk, lambda env, s, k: (3 + 5 * k) % 180, 2, 2, torch.device("cpu"))
    assert tot == 4 and wall > 0 and len(made) >= 2
    assert all(e.steps >= 1 for e in made[:2])


def test_device_replay_memory_has_no_cpu_path():
    import pytest
    import torch
    from meshdqn_b200.replay import DeviceReplayMemory
    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceReplayMemory(16, 180, 400, 17, torch.device("cpu"))


def test_dqn_driver_host_pieces(tmp_path):
    """epsilon schedule (airfoil_dqn.py:454), the DataHandler's five .npy files (:128-133) and the policy-net
    checkpoint names / PyG state_dict keys (:214-218) -- the artefacts deploy_dqn.py and training_results/ read."""
    import math
    import torch
    from meshdqn_b200 import dqn
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    assert dqn.epsilon_threshold(0) == 1.0
    assert abs(dqn.epsilon_threshold(10000) - (0.01 + 0.99 * math.exp(-1.0))) < 1e-15
    pre = str(tmp_path / "run" / "ys930_")
    h = dqn.DataHandler(pre)
    h.add_eps(1.0); h.add_eps(0.99); h.add_loss(0.5)
    h.add_episode([0.1, -1.0], [3, 180])
    h.add_episode([0.2], [7])
    h.write()
    assert np.allclose(np.load(pre + "reward.npy"), [-0.9, 0.2]) and len(np.load(pre + "eps.npy")) == 2
    assert list(np.load(pre + "actions.npy", allow_pickle=True)[0]) == [3, 180]
    assert len(np.load(pre + "rewards.npy", allow_pickle=True)) == 2 and np.load(pre + "losses.npy")[0] == 0.5
    h2 = dqn.DataHandler(pre, restart=True)
    assert h2.num_eps() == 2 and h2.save_dir.endswith("RESTART_")
    torch.manual_seed(0)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n)
    dqn.save_policy_nets(pre, nets[0], nets[1])
    sd = torch.load(pre + "policy_net_1.pt")
    assert {"conv1.lin_l.weight", "conv1.lin_l.bias", "conv1.lin_r.weight", "conv4.lin.weight", "conv4.bias",
            "pool1.weight", "lin3.weight", "lin3.bias"} <= set(sd)
    assert tuple(sd["conv1.lin_l.weight"].shape) == (128, 17) and tuple(sd["lin3.weight"].shape) == (181, 64)
    fresh = NodeRemovalNet(181, conv_width=128, topk=0.1)
    fresh.set_num_nodes(17)
    dqn.load_policy_nets(pre, fresh, fresh)
    assert all(torch.equal(fresh.state_dict()[k], nets[1].state_dict()[k]) for k in sd)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) must print ONE JSON line with the
    contract's keys; ranks other than 0 print nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "replay_train_graphs_per_s" and d["unit"] == "graphs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_dolfin_dof_order_ingest_round_trip():
    """Snapshot arrays in an arbitrary (DOLFIN-like) dof order come back in the package layout through the dof
    coordinates DOLFIN reports (meshdqn_b200/snapshots.py); wrong meshes and non-permutations are refused."""
    from conftest import load_mesh
    from meshdqn_b200.snapshots import DolfinDofMap, p2_points
    from meshdqn_b200.synthetic import synthetic_fields
    from oracle import geom
    coords, cells = load_mesh("ah93w145")
    topo = geom.Topology(cells, len(coords))
    U, P = synthetic_fields(coords, topo.edges, 3, 1)
    pts = p2_points(coords, topo.edges)
    rng = np.random.RandomState(0)
    # a DOLFIN-like numbering: nodes in a random order, the two components of a node adjacent (x then y), plus an
    # independent random order for the pressure space
    node_order = rng.permutation(len(pts))
    xy_u = np.repeat(pts[node_order], 2, axis=0)
    comp_u = np.tile([0, 1], len(pts))
    p_order = rng.permutation(len(coords))
    xy_p = coords[p_order]
    u_dolfin = U[:, node_order, :].reshape(3, -1)
    p_dolfin = P[:, p_order]
    dm = DolfinDofMap(coords, topo.edges, xy_u, comp_u, xy_p)
    assert np.array_equal(dm.velocities(u_dolfin), U) and np.array_equal(dm.pressures(p_dolfin), P)
    assert np.array_equal(dm.dolfin_velocities(U), u_dolfin) and np.array_equal(dm.dolfin_pressures(P), p_dolfin)
    with pytest.raises(ValueError):
        DolfinDofMap(coords + 1e-3, topo.edges, xy_u, comp_u, xy_p)           # another (moved) mesh
    bad = xy_u.copy()
    bad[1] = bad[3]
    with pytest.raises(ValueError):
        DolfinDofMap(coords, topo.edges, bad, comp_u, xy_p)                   # two dofs on one (point, component)


def test_bind_to_gpu_numa_node_reads_sysfs(tmp_path):
    """parallel.bind_to_gpu_numa_node: PCI address -> NUMA node -> that node's CPUs (intersected with what the
    container allows); anything missing leaves the affinity alone."""
    import os
    from meshdqn_b200 import parallel
    before = os.sched_getaffinity(0)
    try:
        cpus = sorted(before)
        root = tmp_path / "sys"
        (root / "bus/pci/devices/0000:1b:00.0").mkdir(parents=True)
        (root / "bus/pci/devices/0000:1b:00.0/numa_node").write_text("1\n")
        (root / "devices/system/node/node1").mkdir(parents=True)
        half = cpus[: max(1, len(cpus) // 2)]
        (root / "devices/system/node/node1/cpulist").write_text(",".join(str(c) for c in half) + ",100000-100003\n")
        assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
        assert parallel._bind_numa_of_pci("0000:1b:00.0", str(root)) == 1
        assert os.sched_getaffinity(0) == set(half)
        os.sched_setaffinity(0, before)
        assert parallel._bind_numa_of_pci("0000:ff:00.0", str(root)) is None          # unknown device
        (root / "bus/pci/devices/0000:1b:00.0/numa_node").write_text("-1\n")            # single-node machine
        assert parallel._bind_numa_of_pci("0000:1b:00.0", str(root)) is None
        assert os.sched_getaffinity(0) == before
    finally:
        os.sched_setaffinity(0, before)


def test_env_workers_step_in_replica_order():
    """dqn._EnvWorkers: R environments stepped on R threads, results in replica order whatever the thread timing; errors
    in a worker surface in the caller."""
    import time
    import pytest
    from meshdqn_b200 import dqn

    class FakeEnv:
        made = 0

        def __init__(self):
            self.k = FakeEnv.made
            FakeEnv.made += 1
            self.t = 0

        def get_state(self):
            return (self.k, self.t)

        def step(self, action):
            if action < 0:
                raise ValueError("bad action")
            time.sleep(0.002 * ((7 * self.k) % 3))      # finish out of order
            self.t += 1
            return (self.k, self.t), float(action), self.t >= 2, {}

    w = dqn._EnvWorkers(FakeEnv, 4, "cpu")
    try:
        assert w.states == [(0, 0), (1, 0), (2, 0), (3, 0)]
        r = w.step_all([5, 6, 7, 8])
        assert [x[0] for x in r] == [(0, 1), (1, 1), (2, 1), (3, 1)] and [x[1] for x in r] == [5.0, 6.0, 7.0, 8.0]
        r = w.step_all([1, 1, 1, 1])
        assert all(x[2] for x in r)
        w.reset(2)
        assert w.states[2] == (4, 0) and w.envs[2].k == 4
        with pytest.raises(ValueError):
            w.step_all([0, -1, 0, 0])
    finally:
        w.close()


def test_torch_library_ops_are_registered_with_fake_kernels():
    """ops.py: the hot-path entry points live under torch.ops.meshdqn_b200 with schemas and fake (meta) kernels, so a traced
    program sees opaque ops of known output shape; there is no CPU kernel behind them."""
    import pytest
    from torch._subclasses.fake_tensor import FakeTensorMode
    from meshdqn_b200 import ops
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    names = {"qnet_forward", "qnet_backward", "qnet_select_action", "mesh_smooth", "polygon_distance", "drag_lift"}
    assert names <= set(dir(torch.ops.meshdqn_b200))
    assert "Tensor[] params" in str(torch.ops.meshdqn_b200.qnet_forward.default._schema)
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    h = ops.net_handle(net)
    assert ops.net_handle(net) == h
    with FakeTensorMode():
        x = torch.empty(360, 17, device="cuda")
        ei = torch.empty(2, 700, dtype=torch.int64, device="cuda")
        nptr = torch.empty(3, dtype=torch.int32, device="cuda")
        q = torch.ops.meshdqn_b200.qnet_forward([], x, ei, nptr, nptr, h, 2, 180, 350)
        am, q2 = torch.ops.meshdqn_b200.qnet_select_action(x, ei, nptr, nptr, h, 2, 180, 350)
        c = torch.empty(50, 2, dtype=torch.float64, device="cuda")
        d = torch.ops.meshdqn_b200.polygon_distance(c, torch.empty(7, dtype=torch.int32, device="cuda"), c[:5])
        dl = torch.ops.meshdqn_b200.drag_lift(c, nptr, nptr, nptr, nptr, torch.empty(5, 80, 2, dtype=torch.float64, device="cuda"),
                                             torch.empty(5, 50, dtype=torch.float64, device="cuda"), 1e-3, 30)
    assert q.shape == (2, 181) and am.shape == (2,) and am.dtype == torch.int32 and q2.shape == (2, 181)
    assert d.shape == (7,) and d.dtype == torch.float64 and dl.shape == (2, 5)
    with pytest.raises(NotImplementedError):
        torch.ops.meshdqn_b200.polygon_distance(torch.zeros(4, 2, dtype=torch.float64), torch.zeros(2, dtype=torch.int32),
                                                torch.zeros(3, 2, dtype=torch.float64))
