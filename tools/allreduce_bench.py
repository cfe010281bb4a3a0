"""Fused all-reduce + Adam (mdq_allreduce_adam) alone, every rank entering at the same time, against NCCL all_reduce +
mdq_adam_step_dev on the same flat gradient.  torchrun --nproc-per-node N tools/allreduce_bench.py"""
import ctypes, os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200 import _lib
from meshdqn_b200.parallel import init_from_env
import torch.distributed._symmetric_memory as symm

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
L, p = _lib.lib(), _lib.ptr
n = int(os.environ.get("AR_N", "502000"))
g = torch.randn(n, device=dev); par = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
step_dev = torch.zeros(2, dtype=torch.int32, device=dev)
stage = symm.empty(int(L.mdq_allreduce_stage_floats(n, world)), dtype=torch.float32, device=dev); stage.zero_()
hdl = symm.rendezvous(stage, dist.group.WORLD)
ptrs = (ctypes.c_uint64 * world)(*[int(x) for x in hdl.buffer_ptrs])
counter = torch.zeros(2, dtype=torch.int32, device=dev)
tiny = torch.zeros(1, device=dev)
torch.cuda.synchronize(); dist.barrier()


def fused():
    _lib.check(L.mdq_allreduce_adam(p(par), p(g), p(m), p(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, p(step_dev), ptrs, rank, world,
                                    p(counter), _lib.stream_ptr()))


def nccl():
    dist.all_reduce(g)
    _lib.check(L.mdq_adam_step_dev(p(par), p(g), p(m), p(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0 / world, p(step_dev),
                                   _lib.stream_ptr()))


def timeit(fn, reps=40):
    ts = []
    for _ in range(reps):
        dist.all_reduce(tiny)                 # device-side rendezvous: every rank's stream reaches fn together
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([float(np.median(ts[5:]))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for name, fn in (("fused two-shot all-reduce + Adam", fused), ("NCCL all_reduce + Adam kernel", nccl), ("fused again", fused)):
    us = timeit(fn)
    if rank == 0:
        print(f"world {world}, n {n}: {name}: {us:.1f} us (median, max over ranks)", flush=True)
        if fn is fused:
            st = stage[-16:].view(torch.int64).cpu().numpy()
            d = np.diff(st[:6])
            print("    block 0 phases (ns): scatter %d | fence+flag+wait %d | reduce+publish %d | fence+flag+wait %d | Adam %d" % tuple(d), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
