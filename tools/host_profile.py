"""Host-side cost of one replay step (cProfile) -- the e2e number is bounded by it once the GPU step is short."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from meshdqn_b200.data import Data
from meshdqn_b200.replay import ReplayBatch, ReplayTrainer
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mk = lambda: Data(x=torch.randn(180, 17, generator=g), edge_index=torch.randint(0, 180, (2, 369), generator=g))
nets = []
for _ in range(2):
    n = NodeRemovalNet(181, 128, 0.1); n.set_num_nodes(17); nets.append(n.to(dev))
trans = [(mk(), int(torch.randint(0, 181, (1,), generator=g)), None if i % 9 == 0 else mk(), float(torch.randn(1, generator=g))) for i in range(256)]
rb = ReplayBatch.from_transitions(trans).pin_memory(slim=True).to(dev).mark_static()
tr = ReplayTrainer(nets[0], nets[1], graphs=True, target_update=1000)
for _ in range(10):
    tr.step(rb)
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    tr.step(rb)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6 * (t1 - t0) / N:.1f} us/step, with drain {1e6 * (t2 - t0) / N:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    tr.step(rb)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
