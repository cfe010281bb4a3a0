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
          seed=137, max_steps_per_episode=None, handler=None, memory=None, verbose=False):
    """Single-process epsilon-greedy DQN (the loop of airfoil_dqn.py:428-503).  ``make_env()`` builds an
    ``Env2DAirfoil``; ``net1`` / ``net2`` are the two policy nets (on ``device``).  Returns the ``DataHandler``."""
    from .replay import DeviceReplayMemory, ReplayTrainer
    dev = torch.device(device) if device is not None else next(net1.parameters()).device
    rng = np.random.RandomState(seed)                       # np.random.seed(137) at :424, one stream here
    random.seed(seed)
    handler = handler or DataHandler(save_prefix or "./")
    trainer = ReplayTrainer(net1, net2, lr=lr, weight_decay=weight_decay, gamma=gamma, target_update=target_update)
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
                loss = float(trainer.step(memory.sample(batch_size, rng=rng)))
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
