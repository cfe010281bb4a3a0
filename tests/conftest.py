import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_mesh(short):
    z = np.load(os.path.join(GOLDEN, f"mesh_{short}.npz"))
    return z["coords"].astype(np.float64), z["cells"].astype(np.int32)


def make_config(N_closest=180, timesteps=10000, threshold=0.001, smooth=True, **extra):
    """The reference's YAML (configs/ray_ys930.yaml:1-39) as a dict; mesh and fields are injected."""
    cfg = {
        "flow_config": {
            "flow_params": {"mu": 1e-3, "rho": 1.0, "inflow": "constant"},
            "geometry_params": {"mesh": None},
            "solver_params": {"dt": 0.001, "solver_type": "lu", "smooth": smooth},
        },
        "agent_params": dict(solver_steps=5000, episodes=1000000, timesteps=timesteps, threshold=threshold,
                             N_closest=N_closest, gt_drag=-1, gt_time=-1, u=-1, p=-1, do_nothing=True, time_reward=0.005,
                             smoothing=True, save_steps=1000, goal_vertices=0.95, plot_dir=""),
        "optimizer": {"lr": 1e-5, "weight_decay": 1e-6, "batch_size": 32},
        "epsilon": {"decay": 10000, "start": 1.0, "end": 0.01, "gamma": 1.0},
    }
    cfg["agent_params"].update(extra)
    return cfg


def oracle_fields(short, T=5, seed=0):
    """Synthetic snapshots on the smoothed fixture mesh, built with the oracle's smoother."""
    from oracle import geom
    from meshdqn_b200.synthetic import synthetic_fields
    coords, cells = load_mesh(short)
    topo = geom.Topology(cells, len(coords))
    xs = geom.smooth(coords, topo, 50)
    U, P = synthetic_fields(xs, topo.edges, T, seed)
    return coords, cells, U, P


def lively_state_dict(net, seed=7, scale=4.0):
    """Seeded weights whose argmax actually depends on the state (default init is bias-dominated)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in net.state_dict().items():
        t = torch.randn(v.shape, generator=g)
        if k.endswith("bias") or ".bias" in k:
            sd[k] = 0.05 * t
        elif k.startswith("pool"):
            sd[k] = t / v.shape[-1] ** 0.5
        else:
            fan_in = v.shape[-1]
            sd[k] = scale * t / fan_in ** 0.5 if k.startswith("lin3") else 1.5 * t / fan_in ** 0.5
    return sd


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from meshdqn_b200 import _lib
    _lib.lib()  # fail loudly if the extension is missing on a GPU box
    return torch.device("cuda:0")
