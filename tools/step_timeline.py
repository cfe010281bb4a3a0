"""Wall-clock timeline (%globaltimer of CTA 0) of the staged kernels of ONE graph-replayed replay step: which kernels of
the two streams really overlap.  MDQ_EARLY_TAIL=1 for the variant with the early tail launch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200 import _lib
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from meshdqn_b200.data import Data
from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 369), generator=g))
L = _lib.lib()
nets = []
for _ in range(2):
    net = NodeRemovalNet(181, 128, 0.1); net.set_num_nodes(17); nets.append(net.to(dev))
trans = []
for i in range(256):
    s = mk(); nx = None if i % 9 == 0 else mk()
    trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
rb = ReplayBatch.from_transitions(trans).pin_memory(slim=True).to(dev).mark_static()
tr = torch.zeros(512, dtype=torch.int64, device=dev)
L.mdq_qnet_set_trace(_lib.ptr(tr))               # baked into the captured launches
trn = ReplayTrainer(nets[0], nets[1], graphs=True, target_update=1000)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["stage0 other", "stage1 other", "tail other (fwd)", "stage0 selected", "stage1 selected", "tail selected (bwd)", "bwd1"]
slots = [(0, 1), (4, 5), (8, 9), (2, 3), (6, 7), (10, 11), (14, 15)]
for rep in range(6):
    flush.fill_(1)
    tr.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); trn.step(rb); trn.flush(); e1.record()
    torch.cuda.synchronize()
t = tr.cpu().numpy()[400:420]
t0 = min(int(v) for v in t if v > 0)
print(f"early tail {'on' if trn.early_tail else 'off'}; step (events, incl. update) {e0.elapsed_time(e1) * 1e3:.1f} us; CTA 0 of each kernel, us from the first stamp:")
for n, (a, b) in zip(names, slots):
    print(f"  {n:22s} {(t[a] - t0) / 1e3:7.1f} -> {(t[b] - t0) / 1e3:7.1f}")
if t[12] > 0:
    print(f"  tail selected: reached the wait at {(t[12] - t0) / 1e3:.1f}, released at {(t[13] - t0) / 1e3:.1f}")
L.mdq_qnet_set_trace(None)
