"""Phase-by-phase cycle trace of the fused Q-network kernels (CTA 0, thread 0, clock64)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from meshdqn_b200 import _lib
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from meshdqn_b200.data import Batch, Data
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 372), generator=g))
net = NodeRemovalNet(181, 128, 0.1); net.set_num_nodes(17); net = net.to(dev)
FWD = ["start", "L1 load", "L1 csr", "L1 agg", "L1 dense", "L1 scores", "L1 rank", "L1 gather+readout+filter",
       "B1 csr", "B1 conv", "B1 pool", "B2 csr", "B2 conv", "B2 pool", "B3 csr", "B3 conv", "B3 pool", "lin1", "lin2", "lin3 (MLP done)", "softmax"]
BWD = FWD + ["bwd huber", "bwd lin3T", "bwd lin2T", "bwd lin1T (MLP done)", "bwd B3 start", "bwd B2 start", "bwd B1 start", "bwd B0 start", "bwd done"]
L = _lib.lib()
net._ensure_packed(); net._net.x_stride = 17
print('occupancy CTAs/SM fwd', L.mdq_qnet_occupancy(net._net, 180, 372, 0), 'bwd', L.mdq_qnet_occupancy(net._net, 180, 372, 1))
for B in (1, 256):
    b = Batch.from_data_list([mk() for _ in range(B)]).to(dev)
    tr = torch.zeros(512, dtype=torch.int64, device=dev)
    for mode, names in (("fwd", FWD), ("bwd", BWD)):
        for rep in range(3):
            tr.zero_()
            L.mdq_qnet_set_trace(_lib.ptr(tr))
            if mode == "fwd":
                with torch.no_grad():
                    net(b)
            else:
                q = net(b); q.sum().backward()
            torch.cuda.synchronize()
        L.mdq_qnet_set_trace(None)
        t = tr.cpu().numpy()
        n = len(names)
        print(f"--- {mode} B={B}: total {(t[n-1]-t[0])} cycles")
        for i in range(1, n):
            print(f"   {names[i]:28s} {t[i]-t[i-1]:8d}")
        if mode == "fwd" and B == 1:
            f = t[128:128 + 3 * 30].reshape(30, 3)
            print("   chunk: wait+sync+issue  compute  gap-to-next")
            for j in range(30):
                nxt = f[j + 1, 0] - f[j, 2] if j + 1 < 30 else 0
                print(f"   {j:3d} {f[j,1]-f[j,0]:8d} {f[j,2]-f[j,1]:8d} {nxt:8d}")

# ---- launch-level timing: back-to-back vs interleaved with an unrelated (small-carveout) kernel ----
b = Batch.from_data_list([mk() for _ in range(256)]).to(dev)
args = net._prep(b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
small = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, pre=None, n=20):
    tot = 0.0
    for _ in range(n):
        if pre is not None:
            pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1000 * tot / n
with torch.no_grad():
    f = lambda: net._launch_forward(*args, False, True)
    for _ in range(3): f()
    print("fwd B=256 back-to-back        : %.1f us" % timeit(f))
    print("fwd B=256 after small fill    : %.1f us" % timeit(f, lambda: small.fill_(1)))
    print("fwd B=256 after 256MB L2 flush: %.1f us" % timeit(f, lambda: flush.fill_(1)))
