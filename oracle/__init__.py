"""CPU oracle for the MeshDQN hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (BaratiLab/MeshDQN) ships no tests, golden
vectors, weights or recorded trajectories, and its stack (torch_geometric,
FEniCS/DOLFIN, shapely) cannot be installed offline, so this package restates
the reference's algorithm per SURVEY.md Appendix A/B.  Golden files under
tests/golden/ pin this restatement, not a run of the original stack.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  meshdqn_b200/ never does.
"""
