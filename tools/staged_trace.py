"""Phase trace (clock64 of CTA 0 / thread 0) of the staged Q-path kernels + CUDA-graph timings without host overhead."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200 import _lib
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from meshdqn_b200.data import Batch, Data
from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 369), generator=g))
L = _lib.lib()
nets = []
for _ in range(2):
    net = NodeRemovalNet(181, 128, 0.1); net.set_num_nodes(17); net = net.to(dev); net.qpath = "staged"; nets.append(net)
net = nets[0]
S0 = ["start", "x loaded, TMEM", "edges counted", "scan", "fill", "sort", "A tile 0", "W1 landed", "tile0 MMA done", "tile1 done",
      "scores", "ranked", "kept rows+filter", "outputs", "(means in regs: absolute)"]
S1 = ["start", "setup", "cat2 built", "GEMM+epi", "scores", "outputs"]
S2 = ["start", "setup", "inputs", "block2", "block3+readout", "MLP", "softmax/lossgrad", "end"]
B1 = ["start", "inputs staged", "pool2 bwd", "conv2T", "dX1", "pool1 bwd (end)"]


def dump(t, base, names, pipe_base=None, pipe_end=None):
    v = t[base:base + len(names)]
    print(f"   total {v[-1] - v[0]} cycles")
    for i in range(1, len(names)):
        print(f"     {names[i]:22s} {v[i] - v[i - 1]:8d}")
    if pipe_base is not None:
        p = t[pipe_base:pipe_end]
        p = p[p != 0]
        if len(p):
            d = np.diff(np.concatenate([[v[0]], p]))
            print("     pipe stamps (delta cycles; per block U W B S F M, per GEMM +acc +epi):", " ".join(str(int(x)) for x in d))


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NOFLUSH = "--noflush" in sys.argv
_fill = flush.fill_
if NOFLUSH:
    flush.fill_ = lambda v: None
    print("*** L2 NOT flushed between runs")
for B in (1, 256):
    b = Batch.from_data_list([mk() for _ in range(B)]).to(dev)
    tr = torch.zeros(512, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for rep in range(3):
            flush.fill_(1)
            tr.zero_()
            L.mdq_qnet_set_trace(_lib.ptr(tr))
            net.select_action(b)
            torch.cuda.synchronize()
    L.mdq_qnet_set_trace(None)
    t = tr.cpu().numpy()
    print(f"=== forward B={B} (L2 flushed)")
    print("  stage0"); dump(t, 0, S0)
    print("  stage1"); dump(t, 32, S1, 40, 96)
    print("  stage2"); dump(t, 96, S2, 112, 256)
trans = []
for i in range(256):
    s = mk(); nx = None if i % 9 == 0 else mk()
    trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
rb = ReplayBatch.from_transitions(trans).to(dev)
trn = ReplayTrainer(nets[0], nets[1])
tr = torch.zeros(512, dtype=torch.int64, device=dev)
for rep in range(3):
    flush.fill_(1)
    tr.zero_()
    L.mdq_qnet_set_trace(_lib.ptr(tr))
    trn.step(rb)
    torch.cuda.synchronize()
L.mdq_qnet_set_trace(None)
t = tr.cpu().numpy()
print("=== replay step B=256: backward kernels of the selected net (stage 0/1 stamps are overwritten by them)")
print("  stage0<save>"); dump(t, 0, S0)
print("  stage1<save>"); dump(t, 32, S1, 40, 96)
print("  stage2<bwd>"); dump(t, 96, S2, 112, 256)
v = t[96:110]
print("   tail backward detail (cycles): d2 %d | d1 %d | dR %d | pool5 bwd %d | conv5^T %d | pool4 bwd + A^T %d | conv4^T + dX2 %d" %
      (v[8] - v[6], v[9] - v[8], v[10] - v[9], v[11] - v[10], v[12] - v[11], v[13] - v[12], v[7] - v[13]))
print("  bwd1"); dump(t, 256, B1, 264, 320)

# ---- CUDA-graph timings (no host launch overhead) ----
def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

for path in ("staged", "fused"):
    for n_ in nets:
        n_.qpath = path
    for B in (1, 256):
        b = Batch.from_data_list([mk() for _ in range(B)]).to(dev)
        with torch.no_grad():
            print(f"[graph] forward B={B} {path}: {graph_time(lambda: net.select_action(b)):.1f} us")
    trn = ReplayTrainer(nets[0], nets[1])
    trn.overlap = False          # one stream: this tool captures the whole step into its own graph
    print(f"[graph] replay step B=256 {path}: {graph_time(lambda: trn.step(rb)):.1f} us")
