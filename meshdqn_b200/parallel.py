"""Data-parallel plumbing for replay training: one process per GPU, torch.distributed (NCCL).

Graphs are independent, so the replay minibatch shards by graph with no data-path collective
(SURVEY.md 8e); the single exchange is an all-reduce of the flat fp32 gradient.  These helpers are
backend-agnostic so the host logic is tested on CPU with gloo (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's RANK / WORLD_SIZE / MASTER_* variables."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index=None, sysfs="/sys"):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are allocated.

    One process per GPU streams ~8 MB of minibatch per replay step from pinned host memory; with eight ranks on one
    node that is the whole job's bottleneck unless every rank's staging pages live on the socket next to its GPU
    (first-touch allocation follows the CPU affinity set here).  Returns the NUMA node, or None when it cannot be
    determined (no GPU, no sysfs entry, a single-node machine, or an affinity mask the container does not allow) --
    in which case nothing is changed.
    """
    try:
        if device_index is None:
            device_index = torch.cuda.current_device()
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"      # sysfs name, e.g. 0000:1b:00.0
    except Exception:
        return None
    return _bind_numa_of_pci(bus, sysfs)


def _bind_numa_of_pci(bus, sysfs="/sys"):
    try:
        with open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def allreduce_mean_(flat: torch.Tensor, world: int, group=None):
    """In-place mean over ranks of a flat gradient buffer (sum all-reduce, then scale)."""
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat.mul_(1.0 / world)
    return flat


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a scalar over ranks (used for device-timed multi-GPU measurements)."""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def evaluate_candidates(q_eval, candidates, rank: int, world: int, group=None, device=None):
    """Batched candidate-action evaluation, sharded by graph (BASELINE.json configs[4]).

    Candidates (one state graph per one-vertex-removed mesh variant) are independent: every rank scores its
    contiguous slice with ``q_eval(list_of_graphs) -> (best_action int [n], best_q float [n])`` (on GPU:
    ``NodeRemovalNet.select_action`` over a ``Batch``, i.e. the fused Q-kernel), and ONE ``all_gather`` of the
    per-candidate (action, q) pairs -- 8 bytes per candidate -- puts the full table on every rank.
    Returns (actions int64 [N], q float32 [N]) in candidate order.  ``device``: where the exchanged table lives
    (default: the device of the candidates' ``x``); a rank whose shard is empty (N < world) must still post a
    tensor on the collective's device (NCCL accepts CUDA tensors only).
    """
    n = len(candidates)
    if device is None:
        device = candidates[0].x.device if n and hasattr(candidates[0], "x") else torch.device("cpu")
    device = torch.device(device)
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        act, q = q_eval(candidates[lo:hi])
        local = torch.stack([act.to(torch.float32).reshape(-1), q.to(torch.float32).reshape(-1)], dim=1).to(device)
    else:
        local = torch.zeros((0, 2), dtype=torch.float32, device=device)
    if world == 1:
        return local[:, 0].long(), local[:, 1]
    cap = (n + world - 1) // world                      # shard sizes differ by at most one: pad to the largest
    buf = torch.zeros((cap, 2), dtype=torch.float32, device=device)
    buf[: hi - lo] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(n, r, world)
        parts.append(out[r][: rhi - rlo])
    table = torch.cat(parts, dim=0)
    return table[:, 0].long(), table[:, 1]


def run_env_replicas(make_env, policy, n_envs: int, n_steps: int, device=None):
    """Independent environment replicas on ONE GPU (the single-episode step does not shard --
    "replicas only"; the reference gets its rollout parallelism from 12 Ray worker processes,
    airfoil_dqn.py:428-503).

    One host thread and one CUDA stream per replica: a replica's device work (re-triangulated mesh topology, the
    ordered smoothing sweep -- a single-CTA, latency-bound kernel -- re-interpolation, drag/lift, state build) runs
    concurrently with the other replicas' kernels on other SMs and with their host-side Qhull calls (PyTorch
    releases the GIL while a thread blocks on its stream).  ``make_env()`` builds an environment on ``device``;
    ``policy(env, state, k)`` returns the action of step k.  Episodes restart on termination.  Returns
    (total environment steps, wall seconds).
    """
    import threading
    import time

    import contextlib

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    on_gpu = dev.type == "cuda"          # a CPU device only exercises the host logic (tests); no streams there
    envs = [make_env() for _ in range(n_envs)]
    states = [e.get_state() for e in envs]
    if on_gpu:
        torch.cuda.synchronize(dev)
    counts = [0] * n_envs
    errors = []
    start = threading.Barrier(n_envs + 1)

    def worker(i):
        try:
            stream = torch.cuda.Stream(dev) if on_gpu else None
            with (torch.cuda.device(dev) if on_gpu else contextlib.nullcontext()), \
                    (torch.cuda.stream(stream) if on_gpu else contextlib.nullcontext()):
                env, s = envs[i], states[i]
                start.wait()
                for k in range(n_steps):
                    s, _, done, _ = env.step(policy(env, s, k))
                    counts[i] += 1
                    if done:
                        env = make_env()
                        s = env.get_state()
                if on_gpu:
                    stream.synchronize()
        except Exception as exc:  # surfaced by the caller
            errors.append(exc)
            try:
                start.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n_envs)]
    for t in threads:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    if errors:
        raise errors[0]
    return sum(counts), dt
