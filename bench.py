#!/usr/bin/env python
"""bench.py -- headline benchmark of the MeshDQN hot path on B200 (one process per GPU).

Workload (BASELINE.json configs[2], the configuration the 1/2/4/8-GPU metric is quoted on):
replay training on a minibatch of 256 ys930-sized state graphs PER GPU (weak scaling) with
NodeRemovalNet(181, conv_width=128, topk=0.1): forward Q1(s), forward Q2(s'), Huber, backward,
gradient all-reduce (NCCL, N>1), Adam.  A "step" is one such replay step; `value` = graphs
(transitions) per second over all ranks with the batch resident in HBM; `e2e` = the same through
ReplayBatch.to(device) from pinned host memory plus the loss read-back every step.
Extras on the same JSON line: forward-only Q-evals, ys930 env steps, re-interpolation vertices/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--small]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

--impl reference times the CPU oracle restatement of the same replay step on the host cores (the
reference stack -- torch_geometric + FEniCS -- cannot be installed offline).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NOTE: the product arm imports nothing from oracle/ or tests/: its inputs come from meshdqn_b200.synthetic and the
# fixture mesh files; only the `cpu_baseline` legs and `--impl reference` (below) execute the oracle.

BATCH = 256
# dram__bytes_read.sum + dram__bytes_write.sum per launch summed over the six kernels of the backward group, from the
# `ncu --set full` capture summarised in profiles/r02_ncu_staged.txt: k_stage0<save> 5.44 + k_stage1<save> 2.73 + k_tail<bwd>
# 1.77 + k_bwd1 3.49 + wgrad_partial 6.21 + wgrad_reduce 1.01 MB.  ncu invalidates the caches before every kernel, so each
# one re-reads from DRAM the intermediates its predecessor left in L2 -- an upper bound on the step's real DRAM traffic.
STAGED_GROUP_TRAFFIC = 20_645_000
METRIC = "replay_train_graphs_per_s"
UNIT = "graphs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(n_gpus):
    return {"workload": "replay training, 256 ys930-sized state graphs per GPU (BASELINE.json configs[2]): "
                        "NodeRemovalNet(181,128,0.1) fwd Q1 + fwd Q2 + Huber + bwd + allreduce + Adam",
            "graphs_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "nodes_per_graph": 180, "features": 17,
            "parallelism": f"dp{n_gpus}", "l2": "256 MiB write between timed steps (L2 flush)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def harvest_transitions(make_env, n, seed):
    """Seeded random-policy episodes (epsilon = 1 branch of airfoil_dqn.py:455-459) -> n transitions on the host."""
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        with quiet():
            env = make_env()
            s = env.get_state()
            for _ in range(60):
                a = int(rng.randint(0, env.N_CLOSEST + 1))
                s2, r, done, _ = env.step(a)
                out.append((s.to("cpu"), a, None if done else s2.to("cpu"), float(r)))
                s = s2
                if done or len(out) >= n:
                    break
    return out[:n]


def random_transitions(n, seed):
    """Seeded random graphs of ys930 state size (180 x 17, ~372 edges): ONLY for short profiler runs
    (--fast-setup); the benchmark proper harvests real transitions."""
    from meshdqn_b200.data import Data
    g = torch.Generator().manual_seed(seed)
    mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 372), generator=g))
    return [(mk(), int(torch.randint(0, 181, (1,), generator=g)), None if i % 16 == 15 else mk(), float(torch.rand(1, generator=g)))
            for i in range(n)]


MESH_YS930 = os.path.join(ROOT, "tests", "golden", "mesh_ys930.npz")      # the reference's xdmf fixture, re-encoded


def env_factory(device=None):
    """device given: the product environment, inputs built by the package itself.  device None: the CPU oracle
    environment (cpu_baseline / --impl reference legs only), fields smoothed by the oracle's own smoother."""
    from meshdqn_b200.synthetic import reference_config, synthetic_fields
    cfg = reference_config()
    if device is None:
        from oracle import geom
        from oracle.env_ref import Env2DAirfoilRef
        z = np.load(MESH_YS930)
        coords, cells = z["coords"].astype(np.float64), z["cells"].astype(np.int32)
        topo = geom.Topology(cells, len(coords))
        U, P = synthetic_fields(geom.smooth(coords, topo, 50), topo.edges, 5, 0)
        cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
        return lambda: Env2DAirfoilRef(cfg, mesh=(coords, cells))
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    from meshdqn_b200.synthetic import fixture_environment_inputs
    coords, cells, U, P = fixture_environment_inputs(MESH_YS930, device)
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    return lambda: Env2DAirfoil(cfg, mesh=(coords, cells), device=device)


# --------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle restatement of the same replay step
# --------------------------------------------------------------------------------------------------
def cpu_replay_steps(transitions, max_seconds, max_steps, warmup):
    from meshdqn_b200.replay import ReplayBatch
    from oracle import gnn_ref
    rb = ReplayBatch.from_transitions(transitions)
    torch.manual_seed(1370)
    nets = []
    for _ in range(2):
        n = gnn_ref.NodeRemovalNet(181, 128, 0.1)
        n.set_num_nodes(17)
        nets.append(n)
    opt = torch.optim.Adam(nets[0].parameters(), lr=1e-5, weight_decay=1e-6)
    mask = rb.next_slot >= 0

    def step():
        opt.zero_grad()
        loss = gnn_ref.replay_loss(nets[0], nets[1], rb.states, rb.actions.long(), (rb.next_states, mask), rb.rewards, 1.0, True)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    k = 0
    while k < max_steps and (time.perf_counter() - t0) < max_seconds:
        step()
        k += 1
    dt = time.perf_counter() - t0
    return k, dt


def cpu_extras(transitions):
    """cpu_baseline objects for the secondary metrics: the CPU oracle timed on this host
    in the same run, bounded samples (~10 s in total).  GNN on all torch threads, geometry (scalar C) on one core."""
    from meshdqn_b200.data import Batch
    from oracle import geom, gnn_ref
    torch.manual_seed(1370)
    net = gnn_ref.NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    states = [t[0] for t in transitions]
    b256 = Batch.from_data_list(states)
    out = {}

    def timed(fn, budget, max_n, warm=1):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        n = 0
        while n < max_n and time.perf_counter() - t0 < budget:
            fn()
            n += 1
        return n, time.perf_counter() - t0
    with torch.no_grad():
        n, dt = timed(lambda: net(b256), 3.0, 100)
        out["q_eval_b256"] = {"value": len(states) * n / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                              "sample": f"{n} forward passes of the same {len(states)}-graph batch, CPU oracle (torch fp32)"}
        n, dt = timed(lambda: net(states[0]), 2.0, 2000)
        out["q_eval_b1"] = {"value": n / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{n} single-graph forward passes, CPU oracle (torch fp32)"}
    with quiet():
        env = env_factory(None)()
        rng = np.random.RandomState(0)
        s = env.get_state()
        t0 = time.perf_counter()
        n = 0
        while n < 400 and time.perf_counter() - t0 < 4.0:
            with torch.no_grad():
                net(s)
            s, r, done, _ = env.step(int(rng.randint(0, 180)))
            n += 1
            if done:
                env = env_factory(None)()
                s = env.get_state()
        dt = time.perf_counter() - t0
    out["env_step_ys930"] = {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
                             "sample": f"{n} steps (Q-eval + Qhull + smoothing + brute-force locate + P2 evaluation + drag/lift), CPU oracle"}
    fs = env.flow_solver
    pts = fs.topo.p2_points(fs.coords)

    def reinterp():
        cell_of, _, _ = geom.locate(pts, env.coords0, env.topo0.cells)
        geom.eval_fields(pts, fs.num_vertices, cell_of, env.coords0, env.topo0, env.U0, env.P0)
    n, dt = timed(reinterp, 2.0, 500)
    out["reinterp_ys930"] = {"value": len(pts) * n / dt, "unit": "vertices/s", "cores": 1, "kind": "port",
                             "sample": f"{n} re-interpolations of {len(pts)} dof points x 5 snapshots (brute-force locate), CPU oracle C"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for multi-process launches; the reference arm is rank 0 alone on the host's
    # cores, so give it all of them back
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncpu))
    tr = harvest_transitions(env_factory(None), BATCH, 1000)
    k, dt = cpu_replay_steps(tr, max_seconds=120.0, max_steps=args.steps, warmup=min(args.warmup, 2))
    val = BATCH * k / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": min(args.warmup, 2),
            "ms_per_step": 1000 * dt / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{k} replay steps of the {BATCH}-graph batch, CPU oracle (torch fp32 restatement of PyG)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "torch_geometric/FEniCS cannot be installed offline; the reference arm is the CPU oracle port"}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    from meshdqn_b200 import _lib
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    from meshdqn_b200.parallel import bind_to_gpu_numa_node, init_from_env, max_over_ranks
    from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
    import torch.distributed as dist

    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None    # pinned staging pages next to this rank's GPU
    L = _lib.lib()
    K, W = args.steps, max(args.warmup, 3)

    tr = random_transitions(BATCH, 1000 + rank) if args.fast_setup else harvest_transitions(env_factory(dev), BATCH, 1000 + rank)
    rb_host = ReplayBatch.from_transitions(tr).pin_memory(slim=True)   # features + 32-bit edges + 32-bit offsets only
    rb_dev = rb_host.to(dev).mark_static()       # stepped K times in place: the trainer may capture the step on it
    torch.manual_seed(1370)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n.to(dev))
    # graphs=True: the step's launches are captured once per (select branch, minibatch buffers) and replayed
    trainer = ReplayTrainer(nets[0], nets[1], lr=1e-5, weight_decay=1e-6, gamma=1.0, target_update=50, graphs=not args.no_graphs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    clk.__enter__()  # sampled from warm-up to the end of the per-kernel pass: all of it is the same replay step under load
    # warm-up covers BOTH `select` branches (airfoil_dqn.py:185-186 flips the trained net every target_update
    # gradients): each branch allocates its own workspace / Adam state on first use, which must not land in a timed step
    for sel in (False, True):
        trainer.select = sel
        for _ in range(W):
            trainer.step(rb_dev)
    trainer.select, trainer.num_grads = True, 0
    barrier()
    # ---- timed: K steps, inputs resident in HBM, L2 flushed between steps, CUDA events per step ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = L.mdq_launch_count()
    barrier()
    for k, (e0, e1) in enumerate(evs):
        flush.fill_(1)
        e0.record()
        trainer.step(rb_dev)
        if k == K - 1:
            trainer.flush()     # the last step's optimizer segment runs on the update stream: it belongs to the K steps
        e1.record()
    barrier()
    launches = int(L.mdq_launch_count() - l0)
    ms_local = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    ms_total = max_over_ranks(ms_local, dev)
    ms_step = ms_total / K
    value = world * BATCH / (ms_step * 1e-3)

    # ---- e2e: every step copies its (pinned) host batch to the device and reads the step's loss back.  The copy of
    # step k+1 runs on a side stream while step k computes (DevicePrefetcher), and the loss of step k is read (pinned
    # D2H + event) while step k+1 is already enqueued -- the public training-loop API a user would call ----
    from meshdqn_b200.replay import DevicePrefetcher
    pf = DevicePrefetcher(dev, static=True)   # two persistent device arenas: fixed addresses for the captured step
    MEM_STEPS = 400                              # device-memory loop: long enough for the per-n_next graphs to be captured
    loss_host = torch.zeros(max(K, 8, MEM_STEPS) + 4, dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        pf.submit(rb_host)
        evs = []
        for k in range(n):
            rb = pf.take()
            if k + 1 < n:
                pf.submit(rb_host)
            loss = trainer.step(rb)
            pf.release()
            loss_host[k:k + 1].copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            evs.append(ev)
            if k >= 1:
                evs[k - 1].synchronize()      # the previous step's loss is on the host now
                float(loss_host[k - 1])
        evs[-1].synchronize()
        return float(loss_host[n - 1])
    e2e_loop(8)      # covers the capture of the step on both arenas
    barrier()
    t0 = time.perf_counter()
    e2e_loop(K)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    e2e_val = world * BATCH * K / e2e_s

    # ---- strong scaling (SURVEY.md 7 hard part 5): the SAME global minibatch of 256 graphs split over the
    # ranks.  32 graphs per GPU at N = 8 are far below one wave of anything, so this is expected to stay well under 2x:
    # every kernel of the step is latency-bound at that size and the all-reduce is a fixed ~20-40 us on top.
    strong = None
    if world > 1:
        per = BATCH // world
        rb_s = ReplayBatch.from_transitions(tr[:per]).pin_memory(slim=True).to(dev).mark_static()
        for sel in (False, True):
            trainer.select = sel
            for _ in range(W + 2):
                trainer.step(rb_s)
        trainer.select = True
        barrier()
        evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k, (e0, e1) in enumerate(evs2):
            flush.fill_(1)
            e0.record()
            trainer.step(rb_s)
            if k == K - 1:
                trainer.flush()
            e1.record()
        barrier()
        ms_s = max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs2), dev) / K
        strong = {"global_batch": per * world, "graphs_per_gpu": per, "ms_per_step": ms_s, "graphs_per_s": per * world / (ms_s * 1e-3)}

    # ---- the same loop fed from the DEVICE-resident replay memory -- sample (host-side index
    # draw + ~12 KB of offsets over PCIe + one gather launch) + step + lagged loss read-back; no minibatch H2D ----
    from meshdqn_b200.replay import DeviceReplayMemory
    e_cap = max(max(int(t[0].edge_index.shape[1]), int(t[2].edge_index.shape[1]) if t[2] is not None else 0) for t in tr)
    mem = DeviceReplayMemory(capacity=BATCH, n_max=180, e_max=e_cap, n_features=int(tr[0][0].x.shape[1]), device=dev)
    for s_, a_, s2_, r_ in tr:
        mem.push(s_, a_, s2_, r_)
    mrng = np.random.RandomState(1234 + rank)

    sampler = mem.static_sampler(BATCH)          # fixed device buffers: the step replays as graphs keyed by the draw's n_next

    def mem_loop(n):
        evs = []
        for k in range(n):
            loss = trainer.step(sampler.sample(rng=mrng))
            loss_host[k:k + 1].copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            evs.append(ev)
            if k >= 1:
                evs[k - 1].synchronize()
                float(loss_host[k - 1])
        evs[-1].synchronize()
    mem_loop(MEM_STEPS // 2)                     # warm-up: every (select, n_next) pair seen twice gets its graphs captured
    barrier()
    t0 = time.perf_counter()
    mem_loop(MEM_STEPS)
    barrier()
    mem_s = max_over_ranks(time.perf_counter() - t0, dev)
    mem_val = world * BATCH * MEM_STEPS / mem_s
    # the loop the reference runs (airfoil_dqn.py:240-257): a NEW sample every step, collated on the host, copied over
    hrng = np.random.RandomState(99 + rank)

    def host_resample_loop(n):
        for _ in range(n):
            pick = hrng.choice(len(tr), BATCH, replace=False)
            rb = ReplayBatch.from_transitions([tr[i] for i in pick]).pin_memory().to(dev)
            float(trainer.step(rb))
    host_resample_loop(2)
    barrier()
    t0 = time.perf_counter()
    host_resample_loop(6)
    barrier()
    host_val = world * BATCH * 6 / max_over_ranks(time.perf_counter() - t0, dev)

    # ---- per-kernel durations (CUDA events around each launch group) for the roofline object ----
    trainer.timers = {}
    for _ in range(K):
        flush.fill_(1)
        trainer.step(rb_dev)
    torch.cuda.synchronize()
    kern = {k: float(np.median([a.elapsed_time(b) for a, b in v])) * 1e3 for k, v in trainer.timers.items()}  # us per launch (median)
    trainer.timers = None
    clk.__exit__()

    # extras: the single-GPU secondary numbers run at N = 1 only; at N > 1 the one extra with a collective (the
    # sharded candidate batch, BASELINE.json configs[4]) runs on EVERY rank -- never inside the rank-0 block
    extras = {}
    if not args.no_extras:
        if world == 1:
            extras = measure_extras(dev, nets[0], rb_dev, flush, args)
        else:
            extras = {"candidate_batch": bench_candidates(nets[0], rb_dev, dev, rank, world, flush),
                      "candidate_variants_250k": bench_candidate_variants(nets[0], dev, rank, world, flush)}
    line = None
    if rank == 0:
        hbm, how = peaks()
        n_nodes = int(rb_dev.states.x.shape[0])
        n_edges = int(rb_dev.states.edge_index.shape[1])
        net = nets[0]
        # algorithmic bytes of the dominant kernel (backward = recompute + gradients): inputs once, parameters once,
        # gradient written once (DESIGN.md "Algorithmic bytes")
        alg_bytes = 4 * n_nodes * 17 + 16 * n_edges + 8 * (BATCH + 1) + 4 * BATCH * 181 + 2 * 4 * net._n_used
        dom = "qnet_bwd+wgrad"
        dom_us = kern.get(dom, float("nan"))
        achieved = alg_bytes / (dom_us * 1e-6) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(world), **({"setup": "fast-setup: random graphs (profiling run, not a bench value)"} if args.fast_setup else {})),
                "clocks": clk.summary(),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(rb_host.h2d_bytes()), "d2h_bytes_per_step": 4,
                        "host_numa_node_rank0": numa},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm",
                             "kernel": "backward group of the selected net: k_stage0<save> + k_stage1<save> (tcgen05) + k_tail<bwd> + "
                                       "k_bwd1 + wgrad_partial + wgrad_reduce",
                             "achieved": achieved, "peak": hbm, "peak_source": how, "unit": "GB/s", "frac": achieved / hbm,
                             # dram__bytes_read+write per launch, summed over the group, from one `ncu --set full` capture
                             # (cold caches; profiles/r02_ncu_staged.txt)
                             "traffic": STAGED_GROUP_TRAFFIC,
                             "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": dom_us,
                             "note": "latency-bound by construction at 256 x 180-node graphs (15 KB per graph, 0.5 MB of weights): "
                                     "the fraction is reported as the contract asks, the per-kernel cycle traces in profiles/ are the "
                                     "optimisation signal; see DESIGN.md 4.1"},
                "kernel_us": kern, "extras": extras}
        line["extras"]["strong_scaling"] = dict(
            strong if strong is not None else {"global_batch": BATCH, "graphs_per_gpu": BATCH, "ms_per_step": ms_step,
                                               "graphs_per_s": value},
            note="same 256-graph global minibatch split over the ranks; expected < 2x at 8 GPUs (32 graphs per GPU: every "
                 "kernel is latency-bound and the all-reduce is a fixed cost) -- the >= 6x target refers to weak scaling")
        line["extras"]["replay_device_memory"] = {
            "graphs_per_s": mem_val, "unit": UNIT, "host_resample_collate_graphs_per_s": host_val,
            "steps": MEM_STEPS, "graphs_cached": len(trainer._graphs),
            "note": "same step, a NEW minibatch every step sampled from the device-resident replay memory into fixed buffers "
                    "(StaticSampler: one gather launch, ~12 KB of offsets over PCIe per step instead of the ~8 MB minibatch; the "
                    "step replays as CUDA graphs keyed by the draw's number of non-terminal transitions); host_resample_collate = a new sample collated on "
                    "the host and copied every step, the loop the reference runs; SURVEY.md 8(f) row 1"}
        if world == 1 and not args.no_cpu:
            k, dt = cpu_replay_steps(tr, max_seconds=15.0, max_steps=1000, warmup=1)
            line["cpu_baseline"] = {"value": BATCH * k / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{k} replay steps of the same {BATCH}-graph batch in {dt:.1f} s, CPU oracle (torch fp32)"}
            if not args.no_extras:
                for name, cb in cpu_extras(tr).items():
                    if name in line["extras"]:
                        line["extras"][name]["cpu_baseline"] = cb
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # Captured CUDA graphs hold NCCL kernels of this communicator; tearing the process group down under them hung
        # (round 2, N = 2).  Drop the graphs, drain the device, meet once more, then leave without running destructors.
        trainer._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def time_events(fn, iters, flush=None):
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters  # ms


def bench_candidates(net, rb_dev, dev, rank, world, flush, total=8192):
    """8192 candidate state graphs (the harvested ys930-sized states, tiled), this rank's shard scored in one fused
    launch, then the all_gather.  Returns candidates/s over all ranks (max over ranks, device-timed)."""
    from meshdqn_b200.data import Batch
    from meshdqn_b200.parallel import max_over_ranks, shard_range
    import torch.distributed as dist
    from meshdqn_b200.airfoilgcnn import graph_ptrs
    st = rb_dev.states
    nptr32, eptr32, nb, _, _ = graph_ptrs(st)
    lo, hi = shard_range(total, rank, world)
    nloc = hi - lo
    reps = (nloc + nb - 1) // nb
    N, E = int(st.x.shape[0]), int(st.edge_index.shape[1])
    x = st.x.repeat(reps, 1)
    rr = torch.arange(reps, device=dev)
    ei = (st.edge_index.unsqueeze(0) + (rr * N).view(-1, 1, 1)).permute(1, 0, 2).reshape(2, -1).contiguous()
    p0, e0 = nptr32.long(), eptr32.long()
    ptr = torch.cat([(p0[:-1].unsqueeze(0) + (rr * N).view(-1, 1)).reshape(-1), torch.tensor([reps * N], device=dev)])
    eptr = torch.cat([(e0[:-1].unsqueeze(0) + (rr * E).view(-1, 1)).reshape(-1), torch.tensor([reps * E], device=dev)])
    b = Batch(x=x, edge_index=ei)
    b.batch = torch.repeat_interleave(torch.arange(reps * nb, device=dev), ptr[1:] - ptr[:-1])
    b.ptr, b.eptr, b.num_graphs = ptr, eptr, reps * nb
    table = torch.empty((world, reps * nb, 2), dtype=torch.float32, device=dev) if world > 1 else None

    def run():
        am, q = net.select_action(b)
        loc = torch.stack([am.float(), q.max(1).values], 1)
        if world > 1:
            dist.all_gather_into_tensor(table.view(-1, 2), loc)
        return loc
    with torch.no_grad():
        for _ in range(3):
            run()
        ms = time_events(run, 10, flush)
    ms = max_over_ranks(ms, dev)
    n_eval = reps * nb * world
    return {"candidates": n_eval, "per_gpu": reps * nb, "ms": ms, "candidates_per_s": n_eval / (ms * 1e-3),
            "nodes_per_graph": 180, "collective": "one all_gather of (action, q) per candidate" if world > 1 else "none (1 GPU)"}


def bench_candidate_variants(net, dev, rank, world, flush, total=8192, n_tri=250_000):
    """BASELINE.json configs[4] as written: `total` one-vertex-removed variants of a ~250k-triangle mesh (local star
    re-triangulation on the host, meshdqn_b200/candidates.py), this rank's shard scored in one launch, then the all_gather."""
    import torch.distributed as dist
    from meshdqn_b200 import candidates as C
    from meshdqn_b200.data import Batch
    from meshdqn_b200.parallel import max_over_ranks, shard_range
    from meshdqn_b200.synthetic import field_values, synthetic_airfoil_mesh
    t0 = time.perf_counter()
    coords, cells, n_ring = synthetic_airfoil_mesh(n_tri, seed=0, n_airfoil=120, order="morton")
    u, p = field_values(coords, 5, 0)
    graphs, meta = C.candidate_state_graphs(coords, cells, np.arange(4, 4 + n_ring), u, p, total, 180)
    gen_s = time.perf_counter() - t0
    n = len(graphs)
    lo, hi = shard_range(n, rank, world)
    b = Batch.from_data_list(graphs[lo:hi]).to(dev)
    cap = (n + world - 1) // world
    loc_buf = torch.zeros((cap, 2), dtype=torch.float32, device=dev)
    table = torch.empty((world * cap, 2), dtype=torch.float32, device=dev) if world > 1 else None

    def run():
        am, q = net.select_action(b)
        loc_buf[: hi - lo] = torch.stack([am.float(), q.max(1).values], 1)
        if world > 1:
            dist.all_gather_into_tensor(table, loc_buf)
        return loc_buf
    with torch.no_grad():
        for _ in range(3):
            run()
        ms = max_over_ranks(time_events(run, 10, flush), dev)
    edges = np.array([int(g.edge_index.shape[1]) for g in graphs])
    return {"candidates": n, "per_gpu": hi - lo, "ms": ms, "candidates_per_s": n / (ms * 1e-3), "triangles": int(len(cells)),
            "nodes_per_graph": 180, "edges_per_graph_mean": float(edges.mean()), "window_offsets": int(meta[:, 0].max()) + 1,
            "host_generation_s": gen_s,
            "collective": "one all_gather of (action, q) per candidate" if world > 1 else "none (1 GPU)",
            "note": "the reference's 180-closest window on a 250k-triangle mesh is a sparse shell around the airfoil (few cells have "
                    "all three vertices inside it), hence the low edge count; `candidate_batch` keeps the ys930-density graphs"}


def measure_extras(dev, net, rb_dev, flush, args):
    """Secondary numbers of BASELINE.json's metric: Q-eval graphs/s, ys930 env steps/s, re-interp vertices/s."""
    from meshdqn_b200.Env2DAirfoil import SourceField
    from meshdqn_b200.flow_solver import DeviceMesh
    from meshdqn_b200.synthetic import synthetic_airfoil_mesh, synthetic_fields
    hbm, _ = peaks()
    out = {}
    with torch.no_grad():
        sargs = net._prep(rb_dev.states)
        f = lambda: net._launch_forward(*sargs, False, True)
        for _ in range(3):
            f()
        ms = time_events(f, 20, flush)
        out["q_eval_b256"] = {"graphs_per_s": BATCH / (ms * 1e-3), "us_per_launch": ms * 1e3}
        from meshdqn_b200.data import Data
        from meshdqn_b200.airfoilgcnn import graph_ptrs
        nptr32, eptr32, *_ = graph_ptrs(rb_dev.states)
        n0 = int(nptr32[1])
        e0 = int(eptr32[1])
        one = Data(x=rb_dev.states.x[:n0].contiguous(), edge_index=rb_dev.states.edge_index[:, :e0].contiguous())
        oargs = net._prep(one)
        f1 = lambda: net._launch_forward(*oargs, False, True)
        for _ in range(3):
            f1()
        ms = time_events(f1, 20, None)
        out["q_eval_b1"] = {"graphs_per_s": 1.0 / (ms * 1e-3), "us_per_launch": ms * 1e3}
    # ys930 episode steps (Qhull on the host inside, as in the reference)
    mk = env_factory(dev)
    with quiet():
        env = mk()
        s = env.get_state()
        rng = np.random.RandomState(0)
        for _ in range(3):              # warm-up: first-use allocations, the torch.ops layer's one-off set-up
            am, _ = net.select_action(s)
            s, r, done, _ = env.step(int(rng.randint(0, 180)))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for _ in range(30):
            am, _ = net.select_action(s)
            s, r, done, _ = env.step(int(rng.randint(0, 180)))
            n += 1
            if done:
                break
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        m1 = env.flow_solver.mesh
        fi = lambda: env.source.interpolate(m1)
        for _ in range(3):
            fi()
        ms = time_events(fi, 20, flush)
    out["env_step_ys930"] = {"steps_per_s": n / dt, "ms_per_step": 1e3 * dt / n, "note": "Q-eval + Qhull (host) + device step"}
    # replicas: the single-episode step does not shard, so rollout throughput comes from independent
    # environments -- one host thread + one CUDA stream each on the same GPU (the reference uses 12 Ray workers)
    from meshdqn_b200.parallel import run_env_replicas
    with quiet():
        rngs = {}

        def pol(env, s, k):
            r = rngs.setdefault(id(env), np.random.RandomState(len(rngs)))
            return int(r.randint(0, 180))
        run_env_replicas(mk, pol, 2, 3, dev)                       # warm-up (per-thread streams, allocator pools)
        nrep = 16
        tot, wall = run_env_replicas(mk, pol, nrep, 16, dev)
    out["env_step_ys930_replicas"] = {"replicas": nrep, "steps_per_s": tot / wall, "steps": tot,
                                      "note": "independent environments on one GPU, one host thread + CUDA stream each"}
    npt = m1.nv + m1.ne
    out["reinterp_ys930"] = {"vertices_per_s": npt / (ms * 1e-3), "us_per_launch": ms * 1e3, "target_points": npt, "T": 5}
    # large synthetic mesh: full-field re-interpolation onto a coarsened copy (throughput mode, no smoothing)
    ntri = 250_000 if args.small else 1_000_000   # BASELINE.json configs[3] names the ~1M-triangle mesh
    coords, cells, _ = synthetic_airfoil_mesh(ntri, seed=0, order="morton")
    m0 = DeviceMesh(coords, cells, dev)
    U0, P0 = synthetic_fields(coords, m0.edges.cpu().numpy(), 5, 0)
    src = SourceField(m0, U0, P0)
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(1)
    isb = m0.on_boundary.cpu().numpy().astype(bool)
    drop = rng.choice(np.nonzero(~isb)[0], max(1, len(coords) // 100), replace=False)
    keep = np.ones(len(coords), bool)
    keep[drop] = False
    c2 = coords[keep]
    t2 = Delaunay(c2).simplices
    b2 = isb[keep]
    t2 = t2[b2[t2].sum(1) != 3]
    m2 = DeviceMesh(c2, t2, dev)
    fi = lambda: src.interpolate(m2)
    for _ in range(3):
        fi()
    # the step's launches (classify, per-leaf interpolation, overflow, closest-cell fallback) replayed as one CUDA
    # graph, so the host's launch latency is not part of the device number
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fi()
    ms = time_events(graph.replay, 10, flush)
    npt = m2.nv + m2.ne
    # algorithmic bytes: target vertices + edge list, source coords/cells/cell->dof map, source
    # coefficients, written dofs + cell ids
    T = 5
    alg = 16 * m2.nv + 8 * m2.ne + 16 * m0.nv + 36 * m0.nc + 8 * T * (2 * (m0.nv + m0.ne) + m0.nv) + \
        8 * T * (2 * npt + m2.nv) + 4 * npt
    # ---- BASELINE.json configs[3]: Q-evaluation of ONE large state graph (all vertices of the synthetic mesh) through the
    # layered path: CSR message passing + tcgen05 3xTF32 node GEMMs + radix-select TopK (gnn_layered.cu)
    from meshdqn_b200 import _lib
    from meshdqn_b200.data import Data
    from meshdqn_b200.synthetic import field_values
    u, pr = field_values(coords, 5, 0)
    xg = np.concatenate([coords, u.transpose(1, 0, 2).reshape(len(coords), -1), pr.T], axis=1).astype(np.float32)
    cc = cells.astype(np.int64)
    eig = np.stack([np.stack([cc[:, 0], cc[:, 0], cc[:, 1]], 1).ravel(), np.stack([cc[:, 1], cc[:, 2], cc[:, 2]], 1).ravel()])
    big = Data(x=torch.from_numpy(xg), edge_index=torch.from_numpy(eig)).to(dev)
    Ng, Eg = int(big.x.shape[0]), int(big.edge_index.shape[1])
    lay = {"nodes": Ng, "directed_edges": Eg, "features": 17}
    with torch.no_grad():
        for gemm in ("tf32x3", "fp32"):
            net.layered_gemm = gemm
            for _ in range(3):
                net.select_action(big)
            gq = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gq):
                net.select_action(big)
            msq = time_events(gq.replay, 10, flush)
            lay[f"q_eval_ms_{gemm}"] = msq
            lay[f"nodes_per_s_{gemm}"] = Ng / (msq * 1e-3)
        net.layered_gemm = "tf32x3"
    # message-passing kernel alone (SAGEConv aggregation over the CSR), against the HBM roofline
    L, p = _lib.lib(), _lib.ptr
    src32 = big.edge_index[0].to(torch.int32).contiguous()
    dst32 = big.edge_index[1].to(torch.int32).contiguous()
    ecount = torch.tensor([Eg], dtype=torch.int32, device=dev)
    row_ptr = torch.empty(Ng + 1, dtype=torch.int32, device=dev)
    colx = torch.empty(Eg, dtype=torch.int32, device=dev)
    scr = torch.empty(int(L.mdq_csr_build_scratch_words(Eg, Ng)), dtype=torch.int32, device=dev)
    _lib.check(L.mdq_csr_build(p(src32), p(dst32), p(ecount), Eg, Ng, p(row_ptr), p(colx), p(scr), _lib.stream_ptr()))
    Amat = torch.empty(Ng, 40, device=dev)
    fa = lambda: _lib.check(L.mdq_sage_aggregate(p(big.x), 17, 0, 17, p(row_ptr), p(colx), Ng, p(Amat), 40, _lib.stream_ptr()))
    fa()
    msa = time_events(fa, 10, flush)
    alg_mp = 4 * (2 * Ng * 17 + Eg + Ng + 1)          # features read once, aggregate written once, CSR indices
    lay["message_passing_17"] = {"us": msa * 1e3, "algorithmic_bytes": alg_mp, "achieved_GBps": alg_mp / (msa * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": alg_mp / (msa * 1e-3) / 1e9 / hbm}
    n1 = Ng // 10                                       # the pooled level's shape: 128-wide rows, N/10 nodes
    keep = (big.edge_index[0] < n1) & (big.edge_index[1] < n1)
    s1, d1 = src32[keep].contiguous(), dst32[keep].contiguous()
    e1 = int(s1.numel())
    ec1 = torch.tensor([e1], dtype=torch.int32, device=dev)
    rp1 = torch.empty(n1 + 1, dtype=torch.int32, device=dev)
    cl1 = torch.empty(max(e1, 1), dtype=torch.int32, device=dev)
    sc1 = torch.empty(int(L.mdq_csr_build_scratch_words(e1, n1)), dtype=torch.int32, device=dev)
    _lib.check(L.mdq_csr_build(p(s1), p(d1), p(ec1), e1, n1, p(rp1), p(cl1), p(sc1), _lib.stream_ptr()))
    x1 = torch.randn(n1, 128, device=dev)
    A1 = torch.empty(n1, 256, device=dev)
    fb = lambda: _lib.check(L.mdq_sage_aggregate(p(x1), 128, 0, 128, p(rp1), p(cl1), n1, p(A1), 256, _lib.stream_ptr()))
    fb()
    msb = time_events(fb, 10, flush)
    alg_mp1 = 4 * (2 * n1 * 128 + e1 + n1 + 1)
    lay["message_passing_128"] = {"us": msb * 1e3, "rows": n1, "algorithmic_bytes": alg_mp1,
                                  "achieved_GBps": alg_mp1 / (msb * 1e-3) / 1e9, "frac_of_hbm_peak": alg_mp1 / (msb * 1e-3) / 1e9 / hbm}
    out["q_eval_large_graph"] = lay
    del big, Amat, A1, x1
    # ---- BASELINE.json configs[4]: batched candidate evaluation, 8192 candidate state graphs sharded by graph over
    # the ranks (1024 per GPU at 8 GPUs), fused Q-kernel + one all_gather of (action, q) per candidate
    from meshdqn_b200.parallel import evaluate_candidates
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    out["candidate_batch"] = bench_candidates(net, rb_dev, dev, rank, world, flush)
    out["candidate_variants_250k"] = bench_candidate_variants(net, dev, rank, world, flush)
    out["reinterp_synthetic"] = {"triangles": int(m0.nc), "target_points": npt, "vertices_per_s": npt / (ms * 1e-3),
                                 "ms_per_launch": ms, "algorithmic_bytes": alg, "achieved_GBps": alg / (ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / hbm,
                                 "kernel": "tiled (k-d leaves staged in shared memory by TMA bulk copies)" if src.tile is not None else "uniform grid",
                                 "leaf_cells": src.leaf_cells, "leaves": src.tile_host.n_leaves if src.tile is not None else 0,
                                 "index_bytes": src.tile_bytes if src.tile is not None else 0,
                                 "vertex_order": "morton", "timing": "CUDA graph replay, L2 flushed between replays"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="250k-triangle synthetic mesh for the large-mesh extras (default: the ~1M-triangle mesh of configs[3])")
    ap.add_argument("--big", action="store_true", help="(default now) kept for compatibility")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--fast-setup", action="store_true", help="random graphs instead of harvested transitions (profiler runs only)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
