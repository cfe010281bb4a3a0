"""First-contact diagnostics on the GPU box (verbose; the real assertions live in tests/)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lively_state_dict, load_mesh, make_config, oracle_fields  # noqa: E402

from meshdqn_b200 import _lib  # noqa: E402
from meshdqn_b200.airfoilgcnn import AirfoilGCNN, NodeRemovalNet  # noqa: E402
from meshdqn_b200.data import Batch, Data  # noqa: E402
from oracle import gnn_ref  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), "lib version", _lib.lib().mdq_version())


def rand_graph(g, n=180, e=369, f=17):
    return Data(x=torch.randn(n, f, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))


def section(name):
    print(f"\n===== {name} =====", flush=True)


what = sys.argv[1:] or ["gnn", "bwd", "geom", "env"]
import traceback

def run_gnn():
    section("GNN forward parity")
    global net, g
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    for sd_name, sd in (("default", ref.state_dict()), ("lively", lively_state_dict(ref))):
        ref.load_state_dict(sd)
        net.load_state_dict(sd)
        net = net.to(dev)
        g = torch.Generator().manual_seed(3)
        d = rand_graph(g)
        q_ref = ref(d)
        q = net(d.to(dev)).cpu()
        rel = ((q - q_ref).abs() / q_ref.abs().clamp_min(1e-30)).max().item()
        print(sd_name, "B=1 max rel", rel, "argmax", int(q.argmax()), int(q_ref.argmax()), "sum", float(q.sum()))
        graphs = [rand_graph(g, n=int(torch.randint(60, 181, (1,), generator=g)), e=int(torch.randint(0, 500, (1,), generator=g)))
                  for _ in range(64)]
        b = Batch.from_data_list(graphs)
        q_ref = ref(b)
        with torch.no_grad():
            q = net(b.to(dev)).cpu()
        rel = ((q - q_ref).abs() / q_ref.abs().clamp_min(1e-30)).max().item()
        print(sd_name, "B=64 ragged max rel", rel, "argmax equal", int((q.argmax(1) == q_ref.argmax(1)).sum()), "/ 64")
        am, _ = net.select_action(b.to(dev))
        print("   fused argmax equal", int((am.cpu().long() == q_ref.argmax(1)).sum()))
        e_ref = ref(b, embedding=True)
        with torch.no_grad():
            e = net(b.to(dev), embedding=True).cpu()
        print("   embedding max abs", (e - e_ref).abs().max().item(), "scale", e_ref.abs().max().item())
    torch.manual_seed(5)
    aref = gnn_ref.AirfoilGCNN(64)
    anet = AirfoilGCNN(64)
    anet.load_state_dict(aref.state_dict())
    anet = anet.to(dev)
    b = Batch.from_data_list([rand_graph(g) for _ in range(8)])
    with torch.no_grad():
        print("AirfoilGCNN", (anet(b.to(dev)).cpu() - aref(b)).abs().max().item(), aref(b).abs().max().item())
    # timing
    b256 = Batch.from_data_list([rand_graph(g) for _ in range(256)]).to(dev)
    with torch.no_grad():
        for _ in range(5):
            net(b256)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            net(b256)
        t1.record(); torch.cuda.synchronize()
        print("fwd B=256: %.1f us/launch" % (t0.elapsed_time(t1) * 1000 / 20))
        d1 = rand_graph(g).to(dev)
        for _ in range(5):
            net(d1)
        t0.record()
        for _ in range(20):
            net(d1)
        t1.record(); torch.cuda.synchronize()
        print("fwd B=1: %.1f us/launch" % (t0.elapsed_time(t1) * 1000 / 20))

def run_bwd():
    section("GNN backward parity")
    torch.manual_seed(1370)
    ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    ref.set_num_nodes(17)
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    sd = lively_state_dict(ref)
    ref.load_state_dict(sd)
    net.load_state_dict(sd)
    net = net.to(dev)
    g = torch.Generator().manual_seed(11)
    graphs = [rand_graph(g, n=int(torch.randint(100, 181, (1,), generator=g)), e=int(torch.randint(100, 500, (1,), generator=g)))
              for _ in range(32)]
    b = Batch.from_data_list(graphs)
    wsel = torch.randn(32, 181, generator=g)
    (ref(b) * wsel).sum().backward()
    (net(b.to(dev)) * wsel.to(dev)).sum().backward()
    rp = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        gr = rp[k].grad
        if gr is None:
            print(f"{k:24s} ref None   mine {None if p.grad is None else float(p.grad.abs().max())}")
            continue
        gm = p.grad.cpu()
        err = (gm - gr).abs().max().item()
        print(f"{k:24s} max|g| {gr.abs().max().item():.3e}  max err {err:.3e}  rel {err / max(gr.abs().max().item(), 1e-30):.2e}")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    b256 = Batch.from_data_list([rand_graph(g) for _ in range(256)]).to(dev)
    w256 = torch.randn(256, 181, device=dev)
    for _ in range(3):
        (net(b256) * w256).sum().backward()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(10):
        (net(b256) * w256).sum().backward()
    t1.record(); torch.cuda.synchronize()
    print("fwd+bwd (autograd path) B=256: %.1f us" % (t0.elapsed_time(t1) * 1000 / 10))

if "geom" in what or "env" in what:
    from oracle import geom as og
    from oracle.env_ref import Env2DAirfoilRef
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil

def run_geom():
    section("geometry parity (ys930)")
    coords, cells, U, P = oracle_fields("ys930")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    renv = Env2DAirfoilRef(cfg, mesh=(coords, cells))
    env = Env2DAirfoil(cfg, mesh=(coords, cells), device=dev)
    m = env.flow_solver.mesh
    rt = renv.flow_solver.topo
    print("ne", m.ne, rt.ne, "nb", m.nb, len(rt.boundary_vertices))
    print("edges equal", np.array_equal(m.edges.cpu().numpy(), rt.edges), "cell_edges equal",
          np.array_equal(m.cell_edges.cpu().numpy(), rt.cell_edges))
    print("nbr equal", np.array_equal(m.nbr_idx[: 2 * m.ne].cpu().numpy(), rt.nbr_idx), "vc equal",
          np.array_equal(m.vc_idx.cpu().numpy(), rt.vc_idx))
    dc = np.abs(m.coordinates() - renv.flow_solver.coords).max()
    print("smoothed coords max abs diff", dc, "bit-equal", np.array_equal(m.coordinates(), renv.flow_solver.coords))
    print("tags equal", np.array_equal(env.flow_solver.tags.cpu().numpy(), renv.flow_solver.tags), "removable equal",
          np.array_equal(env.flow_solver.removable, renv.flow_solver.removable))
    print("gt_drag", env.gt_drag, "rel err", np.abs(env.gt_drag / renv.gt_drag - 1).max(), np.abs(env.gt_lift / renv.gt_lift - 1).max())
    print("dist max abs diff", np.abs(env.distance_lookup - renv.distance_lookup).max())
    s, rs = env.get_state(), renv.get_state()
    print("state x equal", torch.equal(s.x.cpu(), rs.x), "edge_index equal", torch.equal(s.edge_index.cpu(), rs.edge_index),
          tuple(s.edge_index.shape))
    print("coord_map equal", env.coord_map == renv.coord_map)

def run_env():
    section("episode parity")
    for short in ("ys930", "ah93w145"):
        coords, cells, U, P = oracle_fields(short)
        cfg = make_config()
        cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
        renv = Env2DAirfoilRef(cfg, mesh=(coords, cells))
        env = Env2DAirfoil(cfg, mesh=(coords, cells), device=dev)
        torch.manual_seed(1370)
        ref = gnn_ref.NodeRemovalNet(181, 128, 0.1)
        ref.set_num_nodes(17)
        ref.load_state_dict(lively_state_dict(ref))
        net = NodeRemovalNet(181, 128, 0.1)
        net.set_num_nodes(17)
        net.load_state_dict(ref.state_dict())
        net = net.to(dev)
        s, rs = env.get_state(), renv.get_state()
        acts, racts = [], []
        t_env = 0.0
        for i in range(80):
            with torch.no_grad():
                q = net(s)
                a = int(q.argmax())
                rq = ref(rs)
                ra = int(rq.argmax())
            qrel = ((q.cpu() - rq).abs() / rq.abs().clamp_min(1e-30)).max().item()
            t0 = time.time()
            s, r, done, _ = env.step(a)
            torch.cuda.synchronize()
            t_env += time.time() - t0
            rs, rr, rdone, _ = renv.step(ra)
            acts.append(a); racts.append(ra)
            ok_cells = np.array_equal(env.last["cell_of"].cpu().numpy(), renv.last["cell_of"]) if "cell_of" in renv.last and a != 180 else True
            xeq = torch.equal(s.x.cpu(), rs.x)
            if i < 4 or not (ok_cells and xeq and a == ra):
                print(short, i, "a", a, ra, "qrel %.2e" % qrel, "rew", r, rr, "done", done, rdone, "cells_eq", ok_cells, "x_eq", xeq,
                      "miss", int(env.last["miss"]) if "miss" in env.last else None, renv.last.get("nmiss"),
                      "drag rel", np.abs(env.new_drags / renv.new_drags - 1).max())
            if a != ra:
                break
            if done or rdone:
                break
        print(short, "steps", len(acts), "actions equal", acts == racts, "env step ms", 1000 * t_env / len(acts), "distinct actions", len(set(acts)))
for name, fn in (("gnn", run_gnn), ("bwd", run_bwd), ("geom", run_geom), ("env", run_env)):
    if name in what:
        try:
            fn()
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()
print("\nlaunches", _lib.lib().mdq_launch_count())
