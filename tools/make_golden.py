"""Generate tests/golden/*.npz from the CPU oracle (run in the build container).

PARITY UNPINNED: the reference cannot run offline (torch_geometric / DOLFIN / shapely absent) and
ships no golden vectors, so these files pin the oracle restatement, seeded:
  * episode_{ys930,ah93w145}.npz : greedy episode with seeded "lively" NodeRemovalNet weights --
    actions, rewards, terminal flags, per-step drags/lifts, vertex counts, cell-location checksums,
    first state x / edge_index, removable mask, smoothed coords, facet tags
  * qnet_batch.npz               : Q-values, embeddings and parameter gradients for a seeded ragged batch
Usage: python tools/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN, lively_state_dict, make_config, oracle_fields  # noqa: E402

from meshdqn_b200.data import Batch, Data  # noqa: E402
from oracle import gnn_ref  # noqa: E402
from oracle.env_ref import Env2DAirfoilRef  # noqa: E402


def checksum(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.int64)[0]


def episode(short):
    coords, cells, U, P = oracle_fields(short)
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    env = Env2DAirfoilRef(cfg, mesh=(coords, cells))
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net.load_state_dict(lively_state_dict(net))
    s = env.get_state()
    out = dict(coords_smoothed=env.flow_solver.coords.copy(), tags=env.flow_solver.tags.copy(),
               removable=env.flow_solver.removable.copy(), x0=s.x.numpy().copy(), edge_index0=s.edge_index.numpy().copy(),
               gt_drag=env.gt_drag.copy(), gt_lift=env.gt_lift.copy(), distance0=env.distance_lookup.copy())
    acts, rews, dones, drags, lifts, nvs, cellsum, margins = [], [], [], [], [], [], [], []
    for _ in range(200):
        with torch.no_grad():
            q = net(s)[0]
        top2 = torch.topk(q, 2).values
        margins.append(float(top2[0] - top2[1]))
        a = int(q.argmax())
        s, r, done, _ = env.step(a)
        acts.append(a); rews.append(r); dones.append(done)
        drags.append(env.new_drags.copy()); lifts.append(env.new_lifts.copy()); nvs.append(env.flow_solver.num_vertices)
        cellsum.append(checksum(env.last["cell_of"]) if a != 180 else 0)
        if done:
            break
    out.update(actions=np.array(acts), rewards=np.array(rews), dones=np.array(dones), drags=np.array(drags),
               lifts=np.array(lifts), nvs=np.array(nvs), cell_checksums=np.array(cellsum), top2_margin=np.array(margins),
               x_last=s.x.numpy().copy(), edge_index_last=s.edge_index.numpy().copy())
    np.savez_compressed(os.path.join(GOLDEN, f"episode_{short}.npz"), **out)
    print(short, "steps", len(acts), "distinct actions", len(set(acts)), "min top-2 margin", min(margins))


def qnet_batch():
    g = torch.Generator().manual_seed(2024)
    graphs = []
    for _ in range(24):
        n = int(torch.randint(40, 181, (1,), generator=g))
        e = int(torch.randint(0, 500, (1,), generator=g))
        graphs.append(Data(x=torch.randn(n, 17, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g)))
    b = Batch.from_data_list(graphs)
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    q = net(b)
    emb = net(b, embedding=True)
    w = torch.randn(24, 181, generator=g)
    (q * w).sum().backward()
    out = dict(x=b.x.numpy(), edge_index=b.edge_index.numpy(), ptr=b.ptr.numpy(), eptr=b.eptr.numpy(), q=q.detach().numpy(),
               emb=emb.detach().numpy(), w=w.numpy())
    for k, p in net.named_parameters():
        out["param/" + k] = p.detach().numpy()
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "qnet_batch.npz"), **out)
    print("qnet_batch", q.shape)


SPECIAL = {
    # scripted episodes for the branches a greedy policy does not visit:
    #   do-nothing (action N_closest = 180, Env2DAirfoil.py:330-332: the closest-window offset grows, reward recomputed),
    #   a broken removal (strict interpolation -> code 2, :569-573: old mesh back, reward -1, terminal),
    #   running out of vertices (N_closest larger than the removable set, :355-357 / :456-458)
    "do_nothing": dict(cfg={}, actions=[7, 180, 180, 23, 180, 0, 179, 180, 3]),
    "strict_break": dict(cfg=dict(interp_strict_tol=-1.0), actions=[11]),
    "out_of_vertices": dict(cfg=dict(N_closest=700), actions=[5]),
}


def special(short="ys930"):
    out = {}
    for name, spec in SPECIAL.items():
        coords, cells, U, P = oracle_fields(short)
        cfg = make_config(**spec["cfg"])
        cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
        env = Env2DAirfoilRef(cfg, mesh=(coords, cells))
        s = env.get_state()
        rews, dones, nvs, xsum, esum, offs = [], [], [], [], [], []
        for a in spec["actions"]:
            s, r, done, _ = env.step(a)
            rews.append(r); dones.append(done); nvs.append(env.flow_solver.num_vertices)
            xsum.append(checksum(s.x.numpy())); esum.append(checksum(s.edge_index.numpy())); offs.append(env.do_nothing_offset)
            if done:
                break
        out.update({f"{name}/actions": np.array(spec["actions"][:len(rews)]), f"{name}/rewards": np.array(rews),
                    f"{name}/dones": np.array(dones), f"{name}/nvs": np.array(nvs), f"{name}/x_checksums": np.array(xsum),
                    f"{name}/edge_checksums": np.array(esum), f"{name}/offsets": np.array(offs)})
        print(short, name, "rewards", rews, "dones", dones, "nv", nvs)
    np.savez_compressed(os.path.join(GOLDEN, f"special_{short}.npz"), **out)


if __name__ == "__main__":
    special("ys930")
    if "--special-only" in sys.argv:
        sys.exit(0)
    episode("ys930")
    episode("ah93w145")
    qnet_batch()
