"""GPU smoke of the Ray-free DQN driver (meshdqn_b200/dqn.py <-> airfoil_dqn.py:428-503): environment, device replay
memory, fused replay step and the reference's on-disk artefacts working together."""
import contextlib
import io

import numpy as np
import pytest
import torch

from conftest import make_config, oracle_fields

pytestmark = pytest.mark.gpu


def test_dqn_training_loop_runs_and_writes_reference_artefacts(cuda_device, tmp_path):
    from meshdqn_b200 import dqn
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    mk = lambda: Env2DAirfoil(cfg, mesh=(coords, cells), device=cuda_device)
    torch.manual_seed(1370)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n.to(cuda_device))
    w0 = nets[0].state_dict()["lin3.weight"].clone()
    pre = str(tmp_path / "out" / "ah93w145_")
    with contextlib.redirect_stdout(io.StringIO()):
        h = dqn.train(mk, nets[0], nets[1], episodes=2, batch_size=4, eps_decay=6.0, target_update=3, memory_capacity=64,
                      device=cuda_device, save_prefix=pre, max_steps_per_episode=9)
    assert len(h.rewards) == 2 and len(h.actions) == 2 and 2 <= len(h.epss) <= 18
    assert len(h.losses) == len(h.epss) - 3 and all(np.isfinite(h.losses))      # optimisation starts with the 4th transition
    assert h.epss[0] == 1.0 and h.epss[-1] < 0.2                                  # both branches of the epsilon test ran
    assert all(0 <= a <= 180 for ep in h.actions for a in ep)
    assert not torch.equal(nets[0].state_dict()["lin3.weight"], w0)               # the selected net was trained
    for name in ("reward", "rewards", "losses", "actions", "eps"):
        assert len(np.load(pre + name + ".npy", allow_pickle=True)) > 0
    sd = torch.load(pre + "policy_net_2.pt")
    assert torch.equal(sd["lin1.weight"], nets[1].state_dict()["lin1.weight"].cpu())


def _mk_env_and_nets(cuda_device, seed=1370):
    from meshdqn_b200.airfoilgcnn import NodeRemovalNet
    from meshdqn_b200.Env2DAirfoil import Env2DAirfoil
    coords, cells, U, P = oracle_fields("ah93w145")
    cfg = make_config()
    cfg["agent_params"]["u"], cfg["agent_params"]["p"] = U, P
    mk = lambda: Env2DAirfoil(cfg, mesh=(coords, cells), device=cuda_device)
    torch.manual_seed(seed)
    nets = []
    for _ in range(2):
        n = NodeRemovalNet(181, conv_width=128, topk=0.1)
        n.set_num_nodes(17)
        nets.append(n.to(cuda_device))
    return mk, nets


def test_replica_actor_loop_is_reproducible_and_trains(cuda_device, tmp_path):
    """dqn.train_replicas (airfoil_dqn.py:428-520 without Ray): 3 environment replicas stepped on their own threads /
    streams, one batched Q-evaluation per round, one optimisation step per environment step.  Thread timing must not
    leak into the result: two runs from the same seeds give identical actions, losses and weights."""
    from meshdqn_b200 import dqn
    runs = []
    for rep in range(3):
        mk, nets = _mk_env_and_nets(cuda_device)
        pre = str(tmp_path / f"run{rep}" / "ah93w145_")
        with contextlib.redirect_stdout(io.StringIO()):
            h = dqn.train_replicas(mk, nets[0], nets[1], n_envs=3, rounds=6, batch_size=4, eps_decay=8.0, target_update=3,
                                   memory_capacity=64, device=cuda_device, save_prefix=pre, graphs=rep == 2)   # run 2: static sampler + graph replay
        runs.append((h, {k: v.clone() for k, v in nets[0].state_dict().items()}))
        assert len(h.epss) == 18 and h.epss[0] == 1.0 and h.epss[-1] < 0.2
        assert len(h.losses) == 18 - 3 and all(np.isfinite(h.losses))            # memory reaches 4 in round 2 (3 + 3 pushes)
        assert sum(len(a) for a in h.actions) == 18
        assert torch.load(pre + "policy_net_1.pt")["lin3.weight"].shape[0] == 181
    (h0, w0), (h1, w1), (h2, w2) = runs
    assert h0.actions == h1.actions == h2.actions and h0.losses == h1.losses == h2.losses
    assert all(torch.equal(w0[k], w1[k]) and torch.equal(w0[k], w2[k]) for k in w0)
