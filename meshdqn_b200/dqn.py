"""Ray-free DQN rollout / training driver.

Mirrors the per-worker loop of /root/reference/airfoil_dqn.py:428-503 -- epsilon-greedy action selection with
``eps = EPS_END + (EPS_START - EPS_END) * exp(-steps_done / EPS_DECAY)`` (:454), ``env.step``, push of the
``Transition`` (terminal -> ``next_state = None``, :475-481), one optimisation step per environment step once the
memory holds a minibatch (:314-318) -- with the Ray actors replaced by in-process objects:

    ReplayMemory actor      -> ``replay.DeviceReplayMemory`` (transitions stay on the GPU)
    DataWorker + ParameterServer (gradients over RPC, :150-310) -> ``replay.ReplayTrainer`` (fused step, NCCL when multi-GPU)
    DataHandler actor (:70-147) -> ``DataHandler`` here, writing the same files

and the same on-disk artefacts: ``{prefix}policy_net_{1,2}.pt`` state_dicts with the PyG parameter names
(:214-218) and ``{prefix}reward.npy / rewards.npy / losses.npy / actions.npy / eps.npy`` (:128-133), so
``deploy_dqn.py:177-226`` and ``training_results/*.py`` can consume a run made here.
"""
from __future__ import annotations

import math
import os
import random

import numpy as np
import torch


def epsilon_threshold(steps_done, start=1.0, end=0.01, decay=10000.0):
    """airfoil_dqn.py:454."""
    return end + (start - end) * math.exp(-float(steps_done) / float(decay))


class DataHandler:
    """The reference's ``DataHandler`` (airfoil_dqn.py:70-147) without Ray and without the matplotlib plot."""

    def __init__(self, save_prefix, restart=False):
        self.save_dir = save_prefix
        self.rewards, self.ep_rewards, self.losses, self.actions, self.epss = [], [], [], [], []
        if restart:                                        # :88-112
            for name, attr in (("reward", "rewards"), ("rewards", "ep_rewards"), ("losses", "losses"),
                               ("actions", "actions"), ("eps", "epss")):
                try:
                    setattr(self, attr, list(np.load(self.save_dir + name + ".npy", allow_pickle=True)))
                except OSError:
                    setattr(self, attr, [])
            self.save_dir += "RESTART_"
            self.write()

    def add_eps(self, eps):
        self.epss.append(eps)

    def num_eps(self):
        return len(self.epss)

    def add_loss(self, loss):
        self.losses.append(loss)

    def add_episode(self, ep_rew, ep_action):
        self.rewards.append(sum(ep_rew))
        self.ep_rewards.append(ep_rew)
        self.actions.append(ep_action)

    def write(self):                                       # :128-133
        d = os.path.dirname(self.save_dir)
        if d:
            os.makedirs(d, exist_ok=True)
        np.save(self.save_dir + "reward.npy", self.rewards)
        np.save(self.save_dir + "rewards.npy", np.array(self.ep_rewards, dtype=object), allow_pickle=True)
        np.save(self.save_dir + "losses.npy", self.losses)
        np.save(self.save_dir + "actions.npy", np.array(self.actions, dtype=object), allow_pickle=True)
        np.save(self.save_dir + "eps.npy", self.epss)


def save_policy_nets(save_prefix, net1, net2):
    """``ParameterServer.write`` (airfoil_dqn.py:214-218): ``{prefix}policy_net_{1,2}.pt`` = CPU state_dicts."""
    d = os.path.dirname(save_prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    for i, net in ((1, net1), (2, net2)):
        torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, f"{save_prefix}policy_net_{i}.pt")


def load_policy_nets(save_prefix, net1, net2, map_location="cpu"):
    """The RESTART branch (airfoil_dqn.py:163-169) / deploy_dqn.py:196-199."""
    for i, net in ((1, net1), (2, net2)):
        net.load_state_dict(torch.load(f"{save_prefix}policy_net_{i}.pt", map_location=map_location))


def train(make_env, net1, net2, *, episodes, batch_size=32, lr=1e-5, weight_decay=1e-6, gamma=1.0, target_update=50,
          eps_start=1.0, eps_end=0.01, eps_decay=10000.0, memory_capacity=10000, device=None, save_prefix=None,
          seed=137, max_steps_per_episode=None, handler=None, memory=None, verbose=False, graphs=False):
    """Single-process epsilon-greedy DQN (the loop of airfoil_dqn.py:428-503).  ``make_env()`` builds an
    ``Env2DAirfoil``; ``net1`` / ``net2`` are the two policy nets (on ``device``).  Returns the ``DataHandler``."""
    from .replay import DeviceReplayMemory, ReplayTrainer
    dev = torch.device(device) if device is not None else next(net1.parameters()).device
    rng = np.random.RandomState(seed)                       # np.random.seed(137) at :424, one stream here
    random.seed(seed)
    handler = handler or DataHandler(save_prefix or "./")
    trainer = ReplayTrainer(net1, net2, lr=lr, weight_decay=weight_decay, gamma=gamma, target_update=target_update, graphs=graphs)
    sampler = None                                          # graphs=True: minibatches in fixed buffers (StaticSampler)
    env = make_env()
    n_actions = int(env.N_CLOSEST)
    steps_done = handler.num_eps() / 14 if handler.num_eps() else 0       # :436 (restart quirk kept)
    for episode in range(episodes):
        if episode != 0:
            env = make_env()                                # :448-449: a fresh environment per episode
        state = env.get_state()
        if memory is None:
            n_feat = int(state.x.shape[1])
            # the state graph holds only cells whose three vertices are among the N closest (Env2DAirfoil.py:259-280):
            # 3 directed edges per such cell and <= 2 N cells among N planar points, so 6 N bounds it for every
            # later (re-triangulated) mesh as well -- not the 3 * n_cells of the whole mesh
            e_max = max(6 * int(state.x.shape[0]), 2 * int(state.edge_index.shape[1]))
            memory = DeviceReplayMemory(memory_capacity, int(state.x.shape[0]), e_max, n_feat, dev)
        ep_actions, ep_rewards = [], []
        t = 0
        while True:
            sample = rng.random_sample()
            eps = epsilon_threshold(steps_done, eps_start, eps_end, eps_decay)
            steps_done += 1
            if sample > eps:                                # exploit: policy_net_1's argmax (:208-209, :458-461)
                with torch.no_grad():
                    am, _ = net1.select_action(state)
                action = int(am[0])
            else:                                           # explore (:463)
                action = random.sample(range(n_actions + 1), 1)[0]
            next_state, reward, done, _ = env.step(action)
            ep_actions.append(action)
            ep_rewards.append(reward)
            memory.push(state, action, None if done else next_state, reward)
            state = next_state
            loss = None
            if len(memory) >= batch_size:                   # optimize_model (:314-335)
                if graphs and sampler is None:
                    sampler = memory.static_sampler(batch_size)
                mb = sampler.sample(rng=rng) if sampler is not None else memory.sample(batch_size, rng=rng)
                loss = float(trainer.step(mb))
                handler.add_loss(loss)
            handler.add_eps(eps)
            t += 1
            if verbose:
                print(f"episode {episode} step {t}: action {action} reward {reward:.4f} eps {eps:.3f} loss {loss}")
            if done or (max_steps_per_episode is not None and t >= max_steps_per_episode):
                break
        handler.add_episode(ep_rewards, ep_actions)
        if save_prefix is not None:
            handler.write()                                 # :490-491
            save_policy_nets(save_prefix, net1, net2)
    return handler


class _EnvWorkers:
    """R environments stepped concurrently: one host thread and one CUDA stream each (the device work of a step --
    topology, the single-CTA smoothing sweep, re-interpolation, drag / lift, state build -- overlaps the other
    replicas' kernels and their host-side Qhull calls, as in ``parallel.run_env_replicas``).  ``step_all(actions)``
    returns when every environment has taken its step; results come back in replica order, so a run is reproducible
    whatever the thread timing."""

    def __init__(self, make_env, n_envs, device):
        import threading
        self.dev = torch.device(device)
        self.on_gpu = self.dev.type == "cuda"
        self.make_env = make_env
        self.envs = [make_env() for _ in range(n_envs)]
        self.states = [e.get_state() for e in self.envs]
        if self.on_gpu:
            torch.cuda.synchronize(self.dev)
        self._go = threading.Barrier(n_envs + 1)
        self._done = threading.Barrier(n_envs + 1)
        self._actions = [0] * n_envs
        self._results = [None] * n_envs
        self._errors = []
        self._stop = False
        self._threads = [threading.Thread(target=self._run, args=(i,), daemon=True) for i in range(n_envs)]
        for t in self._threads:
            t.start()

    def _run(self, i):
        import contextlib
        stream = torch.cuda.Stream(self.dev) if self.on_gpu else None
        ctx = (lambda: torch.cuda.stream(stream)) if self.on_gpu else contextlib.nullcontext
        if self.on_gpu:
            torch.cuda.set_device(self.dev)
        while True:
            try:
                self._go.wait()
            except Exception:
                return
            if self._stop:
                return
            try:
                with ctx():
                    self._results[i] = self.envs[i].step(self._actions[i])
                    if stream is not None:
                        stream.synchronize()
            except Exception as exc:              # surfaced by step_all
                self._errors.append(exc)
            try:
                self._done.wait()
            except Exception:
                return

    def step_all(self, actions):
        self._actions[:] = [int(a) for a in actions]
        self._go.wait()
        self._done.wait()
        if self._errors:
            raise self._errors[0]
        return list(self._results)

    def reset(self, i):
        self.envs[i] = self.make_env()
        self.states[i] = self.envs[i].get_state()

    def close(self):
        self._stop = True
        try:
            self._go.wait(timeout=5)
        except Exception:
            pass
        for t in self._threads:
            t.join(timeout=5)


def train_replicas(make_env, net1, net2, *, n_envs, rounds, batch_size=32, lr=1e-5, weight_decay=1e-6, gamma=1.0,
                   target_update=50, eps_start=1.0, eps_end=0.01, eps_decay=10000.0, memory_capacity=10000, device=None,
                   save_prefix=None, seed=137, pg=None, handler=None, graphs=False, verbose=False):
    """The reference's multi-worker actor loop (airfoil_dqn.py:428-503 run by ``NUM_WORKERS`` Ray actors, :508-520)
    without Ray: ``n_envs`` environment replicas per process and -- under ``torchrun`` -- one process per GPU.

    Every round (a) ONE batched Q-evaluation of ``net1`` scores all replicas' states (the reference's workers each
    call ``select_action`` on the shared parameter server, :458-461), (b) the replicas take their epsilon-greedy
    step concurrently (``_EnvWorkers``), (c) the transitions go to this process's device replay memory in replica
    order, and (d) one optimisation step runs per environment step taken (:483, each worker's ``optimize_model``).
    ``steps_done`` advances once per environment step of this process, as the reference's per-worker counter does.

    Multi-GPU: with an initialised process group (``pg`` or the default group) ``ReplayTrainer`` averages the
    gradients of all ranks every optimisation step, so every rank keeps bit-identical nets; ranks seed their
    exploration and replay sampling with ``seed + rank``.  All ranks run the same number of rounds and start
    optimising in the same round (same ``n_envs`` and ``batch_size``), so the collectives line up.

    Returns this process's ``DataHandler`` (rank 0 writes the policy nets; every rank writes its own curves under
    ``{save_prefix}rank{r}_`` when there is more than one).
    """
    import torch.distributed as dist
    from .data import Batch
    from .replay import DeviceReplayMemory, ReplayTrainer
    dev = torch.device(device) if device is not None else next(net1.parameters()).device
    world = dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(pg) if world > 1 else 0
    rng = np.random.RandomState(seed + rank)
    pyrng = random.Random(seed + rank)
    prefix = save_prefix
    if prefix is not None and world > 1:
        prefix = f"{save_prefix}rank{rank}_"
    handler = handler or DataHandler(prefix or "./")
    trainer = ReplayTrainer(net1, net2, lr=lr, weight_decay=weight_decay, gamma=gamma, target_update=target_update,
                            process_group=pg, graphs=graphs)       # world > 1: gradients averaged over the ranks
    workers = _EnvWorkers(make_env, n_envs, dev)
    try:
        n_actions = int(workers.envs[0].N_CLOSEST)
        s0 = workers.states[0]
        e_max = max(6 * int(s0.x.shape[0]), 2 * int(s0.edge_index.shape[1]))
        memory = DeviceReplayMemory(memory_capacity, int(s0.x.shape[0]), e_max, int(s0.x.shape[1]), dev)
        sampler = None                                      # graphs=True: minibatches in fixed buffers (StaticSampler)
        ep_actions = [[] for _ in range(n_envs)]
        ep_rewards = [[] for _ in range(n_envs)]
        steps_done = 0
        for rnd in range(rounds):
            with torch.no_grad():
                greedy, _ = net1.select_action(Batch.from_data_list(workers.states).to(dev))
            greedy = greedy.cpu().tolist()
            actions, epss = [], []
            for i in range(n_envs):
                sample = rng.random_sample()
                eps = epsilon_threshold(steps_done, eps_start, eps_end, eps_decay)
                steps_done += 1
                epss.append(eps)
                actions.append(int(greedy[i]) if sample > eps else pyrng.sample(range(n_actions + 1), 1)[0])
            results = workers.step_all(actions)
            for i, (next_state, reward, done, _) in enumerate(results):
                ep_actions[i].append(actions[i])
                ep_rewards[i].append(reward)
                memory.push(workers.states[i], actions[i], None if done else next_state, reward)
                workers.states[i] = next_state
                if done:
                    handler.add_episode(ep_rewards[i], ep_actions[i])
                    ep_actions[i], ep_rewards[i] = [], []
                    workers.reset(i)
            for i in range(n_envs):
                if len(memory) >= batch_size:
                    if graphs and sampler is None:
                        sampler = memory.static_sampler(batch_size)
                    mb = sampler.sample(rng=rng) if sampler is not None else memory.sample(batch_size, rng=rng)
                    loss = float(trainer.step(mb))
                    handler.add_loss(loss)
                handler.add_eps(epss[i])
            if verbose and rank == 0:
                print(f"round {rnd}: actions {actions} eps {epss[-1]:.3f} memory {len(memory)}")
        trainer.flush()
        for i in range(n_envs):                             # unfinished episodes are recorded as they stand
            if ep_actions[i]:
                handler.add_episode(ep_rewards[i], ep_actions[i])
        if save_prefix is not None:
            handler.write()
            if rank == 0:
                save_policy_nets(save_prefix, net1, net2)
    finally:
        workers.close()
    return handler
