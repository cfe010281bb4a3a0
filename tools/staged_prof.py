"""Short driver for ncu: staged forward at B=1 / 256 and replay steps (both select branches)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from meshdqn_b200.data import Batch, Data
from meshdqn_b200.replay import ReplayBatch, ReplayTrainer

dev = torch.device("cuda:0")
path = sys.argv[1] if len(sys.argv) > 1 else "staged"
g = torch.Generator().manual_seed(9)


def rand_graph(n=180, e=369, f=17):
    return Data(x=torch.randn(n, f, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))


nets = []
for _ in range(2):
    net = NodeRemovalNet(181, 128, 0.1)
    net.set_num_nodes(17)
    net = net.to(dev)
    net.qpath = path
    nets.append(net)
b1 = rand_graph().to(dev)
b256 = Batch.from_data_list([rand_graph() for _ in range(256)]).to(dev)
trans = []
for i in range(256):
    s = rand_graph()
    nx = None if i % 9 == 0 else rand_graph()
    trans.append((s, int(torch.randint(0, 181, (1,), generator=g)), nx, float(torch.randn(1, generator=g))))
rb = ReplayBatch.from_transitions(trans).to(dev)
tr = ReplayTrainer(nets[0], nets[1], target_update=3)
with torch.no_grad():
    for _ in range(3):
        nets[0].select_action(b1)
    for _ in range(3):
        nets[0].select_action(b256)
for _ in range(6):
    tr.step(rb)
torch.cuda.synchronize()
print("ok")
