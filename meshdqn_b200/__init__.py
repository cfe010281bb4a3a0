"""meshdqn_b200 -- B200-native (sm_100a) implementation of MeshDQN's data-parallel hot path.

Public surface mirrors the reference's modules for that path:
``airfoilgcnn`` (NodeRemovalNet, AirfoilGCNN), ``Env2DAirfoil`` (Env2DAirfoil),
``flow_solver`` (FlowSolver mesh services), ``probes`` (DragProbe, LiftProbe), plus
``data`` (Data/Batch/DataLoader stand-ins for torch_geometric) and ``replay``
(replay-minibatch training step, NCCL data parallel).
"""
from .data import Batch, Data, DataLoader  # noqa: F401

__all__ = ["Data", "Batch", "DataLoader"]
