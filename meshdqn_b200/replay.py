"""Replay-minibatch training step on device (data parallel over NCCL).

Restates the math of ``DataWorker._get_data`` / ``compute_gradients`` and the parameter
server's Adam step (/root/reference/airfoil_dqn.py:172-176, 240-310):

    pred   = Q1(s)[a]                                      (:264)
    target = r + gamma * max_a Q2(s')   (0 when s' is terminal)   (:267-281)
    loss   = HuberLoss(delta=1, mean)(pred, target)        (:303-304)
    backward through the *selected* net only               (:258-262, 272-276)
    Adam(lr, weight_decay) + MultiStepLR([500k, 1M, 1.5M], 0.1)   (:172-176)

The reference routes this through Ray actors (replay actor -> worker -> parameter server); its
apply_gradients steps the optimizer before the gradients are set and re-creates the optimizer on
every call, losing Adam's moments (:188-199).  Those are defects, not semantics (SURVEY.md App. B,
"do not replicate"): here each net keeps one Adam state and the step is
forward(other net) -> backward(selected net; recomputes its forward and evaluates the Huber term in
place) -> weight-gradient reduce -> loss -> [all-reduce] -> Adam.

Multi-GPU: graphs are independent, so each rank takes its own B transitions; the only exchange is
ONE all-reduce (sum) of the flat fp32 gradient (the first ``n_used`` floats of the flat buffer,
~0.5 MB) over NCCL/NVLink; the mean over ranks is folded into the Adam kernel (grad_scale).
"""
from __future__ import annotations

import torch

from . import _lib
from .airfoilgcnn import graph_ptrs


class ReplayBatch:
    """A collated minibatch resident on one device.

    states / next_states: ``Batch`` objects; ``next_slot`` i32 [B] = row of ``next_states`` holding
    transition b's next state, -1 if terminal; ``actions`` i32 [B]; ``rewards`` f32 [B].
    """

    def __init__(self, states, actions, next_states, next_slot, rewards, owner=None):
        self.states, self.actions, self.next_states, self.next_slot, self.rewards = \
            states, actions, next_states, next_slot, rewards
        if owner is None:  # owner[g] = transition whose next state is row g of next_states (inverse of next_slot)
            ns = next_slot.cpu()
            idx = torch.nonzero(ns >= 0, as_tuple=False)[:, 0]
            owner = torch.zeros(max(int(idx.numel()), 1), dtype=torch.int32)
            owner[ns[idx].long()] = idx.to(torch.int32)
            owner = owner.to(next_slot.device)
        self.owner = owner

    @classmethod
    def from_transitions(cls, transitions):
        """transitions: iterable of (state Data, action int, next_state Data | None, reward float)
        -- the reference's ``Transition`` tuples (airfoil_dqn.py:46-47)."""
        from .data import Batch
        states, actions, nxt, rewards = zip(*transitions)
        slot, non_final = [], []
        for s in nxt:
            if s is None:
                slot.append(-1)
            else:
                slot.append(len(non_final))
                non_final.append(s)
        return cls(Batch.from_data_list(states), torch.tensor([int(a) for a in actions], dtype=torch.int32),
                   Batch.from_data_list(non_final) if non_final else None, torch.tensor(slot, dtype=torch.int32),
                   torch.tensor([float(r) for r in rewards], dtype=torch.float32))

    # -- host staging: ONE pinned arena, ONE host->device copy per minibatch ---------------------------
    _arena = None      # uint8 tensor holding every tensor of the batch (pinned host, or its device copy)
    _layout = None     # [(slot, offset, dtype, shape)], slot = ("states", "x") / ("", "actions") / ...

    def _named_tensors(self):
        out = [(("", k), getattr(self, k)) for k in ("actions", "next_slot", "rewards", "owner")]
        for name in ("states", "next_states"):
            b = getattr(self, name)
            if b is None:
                continue
            m = b._host_meta()
            if m is None:
                raise RuntimeError("ReplayBatch.pin_memory: collate on the host first (Batch.from_data_list)")
            out += [((name, k), getattr(b, k)) for k in ("x", "edge_index", "batch", "ptr", "eptr")]
            out += [((name, "_ptr32"), m[0]), ((name, "_eptr32"), m[1])]
        return out

    def _from_arena(self, arena):
        """Rebuild the batch as views into `arena` (same layout on host and device)."""
        from .data import Batch
        parts = {"": {}, "states": {}, "next_states": {}}
        for (grp, key), off, dtype, shape in self._layout:
            n = 1
            for d in shape:
                n *= d
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            parts[grp][key] = arena[off:off + nbytes].view(dtype).view(shape)
        batches = {}
        for name in ("states", "next_states"):
            src = getattr(self, name)
            if src is None:
                batches[name] = None
                continue
            d = parts[name]
            b = Batch(x=d["x"], edge_index=d["edge_index"])
            b.batch, b.ptr, b.eptr, b.num_graphs = d["batch"], d["ptr"], d["eptr"], src.num_graphs
            m = src._host_meta() or src.__dict__["_meta"]
            if arena.is_cuda:
                b.__dict__["_mdq_ptrs"] = (d["_ptr32"], d["_eptr32"], m[2], m[3], m[4])
            b.__dict__["_meta"] = (d["_ptr32"], d["_eptr32"], m[2], m[3], m[4])
            batches[name] = b
        p = parts[""]
        out = ReplayBatch(batches["states"], p["actions"], batches["next_states"], p["next_slot"], p["rewards"], p["owner"])
        out._arena, out._layout = arena, self._layout
        return out

    def pin_memory(self):
        """Pack every tensor of the minibatch into one pinned host arena (256-byte aligned sections): `to(device)` is
        then a single cudaMemcpyAsync instead of ~18 small ones, which is what the host side of the e2e loop costs."""
        named = self._named_tensors()
        layout, off = [], 0
        for slot, t in named:
            layout.append((slot, off, t.dtype, tuple(t.shape)))
            off += (t.numel() * t.element_size() + 255) // 256 * 256
        arena = torch.empty(max(off, 256), dtype=torch.uint8).pin_memory()
        for (slot, o, dtype, shape), (_, t) in zip(layout, named):
            n = t.numel() * t.element_size()
            if n:
                arena[o:o + n].view(dtype).view(shape).copy_(t)
        self._layout = layout
        return self._from_arena(arena)

    def to(self, device, non_blocking=True):
        if self._arena is not None:
            return self._from_arena(self._arena.to(device, non_blocking=non_blocking))
        return ReplayBatch(self.states.to(device, non_blocking=non_blocking), self.actions.to(device, non_blocking=non_blocking),
                           None if self.next_states is None else self.next_states.to(device, non_blocking=non_blocking),
                           self.next_slot.to(device, non_blocking=non_blocking),
                           self.rewards.to(device, non_blocking=non_blocking), self.owner.to(device, non_blocking=non_blocking))

    def tensors(self):
        if self._arena is not None:
            return [self._arena]
        out = [self.actions, self.next_slot, self.rewards, self.owner]
        for b in (self.states, self.next_states):
            if b is None:
                continue
            out += [v for v in b.__dict__.values() if torch.is_tensor(v)]
            out += [v for v in (b.__dict__.get("_mdq_ptrs") or ()) if torch.is_tensor(v)]
        return out

    def h2d_bytes(self):
        if self._arena is not None:
            return int(self._arena.numel())
        n = 0
        for b in (self.states, self.next_states):
            if b is None:
                continue
            n += b.x.numel() * b.x.element_size() + b.edge_index.numel() * 8 + b.ptr.numel() * 8 + b.eptr.numel() * 8 + \
                b.batch.numel() * 8
        return n + self.actions.numel() * 4 + self.next_slot.numel() * 4 + self.rewards.numel() * 4 + self.owner.numel() * 4


class DevicePrefetcher:
    """Host -> device staging of replay minibatches on a side stream, one batch ahead of the training step
    (the collated ``ReplayBatch`` must be pinned).  ``submit(host_batch)`` starts the copies, ``take()`` makes the
    current stream wait for them and hands the device batch over; with a ~10 MB batch the PCIe copy of step k+1
    overlaps the kernels of step k instead of preceding them."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def submit(self, host_batch: "ReplayBatch"):
        with torch.cuda.stream(self.stream):
            dev = host_batch.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def take(self) -> "ReplayBatch":
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev.tensors():          # allocated on the side stream, consumed on the current one
            t.record_stream(cur)
        return dev


def multistep_lr(base_lr, step, milestones=(500000, 1000000, 1500000), gamma=0.1):
    """optim.lr_scheduler.MultiStepLR as used at airfoil_dqn.py:175-176."""
    return base_lr * gamma ** sum(1 for m in milestones if step >= m)


class ReplayTrainer:
    def __init__(self, policy_net_1, policy_net_2, lr=1e-5, weight_decay=1e-6, gamma=1.0, target_update=50,
                 betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.nets = (policy_net_1, policy_net_2)
        self.lr, self.wd, self.gamma = float(lr), float(weight_decay), float(gamma)
        self.betas, self.eps = betas, float(eps)
        self.target_update = int(target_update)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.select = True
        self.num_grads = 0
        self._state = [None, None]
        self.timers = None  # optional {name: [(start_event, end_event), ...]} for per-kernel timing

    # -- helpers ----------------------------------------------------------------------------
    def _adam_state(self, i):
        if self._state[i] is None:
            net = self.nets[i]
            net._ensure_packed()
            z = torch.zeros_like(net._flat)
            self._state[i] = dict(m=z, v=z.clone(), g=torch.zeros_like(net._flat), step=0)
        return self._state[i]

    def _timed(self, name):
        tr = self

        class _T:
            def __enter__(self_inner):
                if tr.timers is not None:
                    self_inner.e0 = torch.cuda.Event(enable_timing=True)
                    self_inner.e1 = torch.cuda.Event(enable_timing=True)
                    self_inner.e0.record()

            def __exit__(self_inner, *a):
                if tr.timers is not None:
                    self_inner.e1.record()
                    tr.timers.setdefault(name, []).append((self_inner.e0, self_inner.e1))

        return _T()

    # -- one replay step ----------------------------------------------------------------------
    @torch.no_grad()
    def step(self, batch: ReplayBatch, fused: bool = True):
        """Returns the Huber loss (device scalar tensor); parameters of the selected net are updated.

        fused=True (default): forward of the NON-selected net only; the selected net's backward kernel
        recomputes its own forward and evaluates the Huber term in place (mdq_qnet_replay_backward):
        launches = forward + memset + backward + 2 weight-gradient + loss + Adam.
        fused=False: the reference's literal sequence forward(Q1), forward(Q2), Huber, backward (used by tests).
        """
        net1, net2 = self.nets
        dev = batch.states.x.device
        L = _lib.lib()
        p = _lib.ptr
        B = int(batch.actions.shape[0])
        s_args = net1._prep(batch.states)
        n_args = net2._prep(batch.next_states) if batch.next_states is not None else None
        n_next = int(n_args[4]) if n_args is not None else 0
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        sel = 0 if self.select else 1
        net = self.nets[sel]
        st = self._adam_state(sel)
        A = net._net.out_dim
        if fused and (self.select or n_args is not None):
            if self.select:
                q_other = None
                if n_args is not None:
                    with self._timed("qnet_fwd"):
                        q_other, _, _ = net2._launch_forward(*n_args, False, False)
                args, mode, index = s_args, 1, batch.next_slot
            else:
                with self._timed("qnet_fwd"):
                    q_other, _, _ = net1._launch_forward(*s_args, False, False)
                args, mode = n_args, 2
                index = batch.owner
            scalar = torch.empty(int(args[4]), dtype=torch.float32, device=dev)
            with self._timed("qnet_bwd+wgrad"):
                net._launch_replay_backward(*args, mode, batch.actions, batch.rewards, index, batch.next_slot, q_other, B,
                                            self.gamma, scalar, loss, st["g"])
        else:
            with self._timed("qnet_fwd"):
                q1, _, _ = net1._launch_forward(*s_args, False, False)
            q2 = None
            if n_args is not None:
                with self._timed("qnet_fwd"):
                    q2, _, _ = net2._launch_forward(*n_args, False, False)
            gq = torch.empty((B if self.select else max(n_next, 1), A), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev), self._timed("huber"):
                rc = L.mdq_huber_replay(p(q1), p(q2), p(batch.actions), p(batch.rewards), p(batch.next_slot), B, n_next, A,
                                        self.gamma, 1 if self.select else 0, p(loss), p(gq) if self.select else None,
                                        None if self.select else p(gq), _lib.stream_ptr())
            _lib.check(rc, "mdq_huber_replay")
            if self.select or n_args is not None:
                args = s_args if self.select else n_args
                with self._timed("qnet_bwd+wgrad"):
                    net._launch_backward(*args, gq, st["g"])
            else:
                st["g"].zero_()
        n_used = net._n_used
        if self.world > 1:
            with self._timed("allreduce"):
                torch.distributed.all_reduce(st["g"][:n_used], group=self.pg)
        st["step"] += 1
        lr = multistep_lr(self.lr, self.num_grads)
        with torch.cuda.device(dev), self._timed("adam"):
            rc = L.mdq_adam_step(p(net._flat), p(st["g"]), p(st["m"]), p(st["v"]), n_used, lr, self.betas[0], self.betas[1],
                                 self.eps, self.wd, 1.0 / self.world, st["step"], _lib.stream_ptr())
        _lib.check(rc, "mdq_adam_step")
        self.num_grads += 1
        if self.num_grads % self.target_update == 0:
            self.select = not self.select
        return loss
