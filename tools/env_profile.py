"""Host-side profile of the ys930 environment step (cProfile, cumulative) -- where the 13-14 ms per step go."""
import contextlib, cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
dev = torch.device("cuda:0")
mk = bench.env_factory(dev)
with contextlib.redirect_stdout(io.StringIO()):
    env = mk(); s = env.get_state()
    for a in (3, 7):
        env.step(a)
rng = np.random.RandomState(0)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter(); n = 0
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(25):
        s, r, done, _ = env.step(int(rng.randint(0, 180))); n += 1
        if done: break
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("env steps/s", n / dt, "ms/step", 1e3 * dt / n)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
