"""Replay step timed per `select` branch (airfoil_dqn.py:258-276): mode 1 = backward through Q1(s)[a],
mode 2 = backward through max_a Q2(s').  CUDA events, L2 flushed between steps.

usage: python tools/mode_bench.py [harvest]   (default: random ys930-sized graphs)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from meshdqn_b200.airfoilgcnn import NodeRemovalNet  # noqa: E402
from meshdqn_b200.replay import ReplayBatch, ReplayTrainer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    harvest = len(sys.argv) > 1 and sys.argv[1] == "harvest"
    tr = bench.harvest_transitions(bench.env_factory(dev), 256, 1000) if harvest else bench.random_transitions(256, 1000)
    rb = ReplayBatch.from_transitions(tr).to(dev)
    torch.manual_seed(1370)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n.to(dev))
    trainer = ReplayTrainer(nets[0], nets[1], lr=1e-5, weight_decay=1e-6, gamma=1.0, target_update=1 << 30)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    iters = int(os.environ.get("ITERS", "20"))
    for sel in (True, False, True, False):
        trainer.select = sel
        for _ in range(3):
            trainer.step(rb)
        torch.cuda.synchronize()
        ts = []
        trainer.timers = {}
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            trainer.step(rb)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        kern = {k: round(float(np.median([a.elapsed_time(b) for a, b in v])) * 1e3, 1) for k, v in trainer.timers.items()}
        trainer.timers = None
        print(f"select={sel}: step median {np.median(ts):.1f} us  min {np.min(ts):.1f}  groups {kern}", flush=True)


if __name__ == "__main__":
    main()
