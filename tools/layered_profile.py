"""One layered Q-evaluation of the ~1M-triangle state graph, for `ncu --metrics gpu__time_duration.sum`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from layered_bench import state_graph
from meshdqn_b200.airfoilgcnn import NodeRemovalNet
from conftest import lively_state_dict
from oracle import gnn_ref
ntri = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
d = state_graph(ntri).to(dev)
torch.manual_seed(1370)
ref = gnn_ref.NodeRemovalNet(181, 128, 0.1); ref.set_num_nodes(17)
net = NodeRemovalNet(181, 128, 0.1); net.set_num_nodes(17); net.load_state_dict(lively_state_dict(ref)); net = net.to(dev)
net.layered_gemm = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
with torch.no_grad():
    for _ in range(3):
        am, q = net.select_action(d)
torch.cuda.synchronize()
print("argmax", int(am[0]))
