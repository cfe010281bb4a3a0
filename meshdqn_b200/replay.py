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

import os

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
    static = False     # True: these device buffers are refilled in place / reused step after step, so a ReplayTrainer with
                       # graphs=True may capture its launches on them (set by mark_static / DevicePrefetcher(static=True))
    _arena = None      # uint8 tensor holding every tensor of the batch (pinned host, or its device copy)
    _layout = None     # [(slot, offset, dtype, shape)], slot = ("states", "x") / ("", "actions") / ...

    def _named_tensors(self, slim=False):
        out = [(("", k), getattr(self, k)) for k in ("actions", "next_slot", "rewards", "owner")]
        for name in ("states", "next_states"):
            b = getattr(self, name)
            if b is None:
                continue
            m = b._host_meta()
            if m is None:
                raise RuntimeError("ReplayBatch.pin_memory: collate on the host first (Batch.from_data_list)")
            if slim:     # what the kernels read: features, 32-bit edges, 32-bit graph offsets
                out += [((name, "x"), b.x), ((name, "edge_index"), b.edge_index.to(torch.int32))]
            else:
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
            b.batch, b.ptr, b.eptr, b.num_graphs = d.get("batch"), d.get("ptr"), d.get("eptr"), src.num_graphs
            m = src._host_meta() or src.__dict__["_meta"]
            if arena.is_cuda:
                b.__dict__["_mdq_ptrs"] = (d["_ptr32"], d["_eptr32"], m[2], m[3], m[4])
            b.__dict__["_meta"] = (d["_ptr32"], d["_eptr32"], m[2], m[3], m[4])
            batches[name] = b
        p = parts[""]
        out = ReplayBatch(batches["states"], p["actions"], batches["next_states"], p["next_slot"], p["rewards"], p["owner"])
        out._arena, out._layout = arena, self._layout
        return out

    def mark_static(self):
        """Promise that this device batch's buffers stay where they are and are only ever refilled in place."""
        self.static = True
        return self

    def pin_memory(self, slim=False):
        """Pack every tensor of the minibatch into one pinned host arena (256-byte aligned sections): `to(device)` is
        then a single cudaMemcpyAsync instead of ~18 small ones, which is what the host side of the e2e loop costs.
        ``slim``: only what the kernels read travels -- ``edge_index`` as int32 and no ``batch`` / int64 offset
        vectors (-25 % bytes over PCIe); the device batch then has ``batch = ptr = eptr = None`` and 32-bit edges."""
        named = self._named_tensors(slim)
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

    def __init__(self, device, static=False):
        """``static``: the copies land in two persistent device arenas used alternately, so consecutive minibatches of
        the same layout come back at (two) fixed addresses -- what a captured CUDA graph of the training step needs
        (``ReplayTrainer(graphs=True)``) -- and nothing is allocated per step."""
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None
        self.static = bool(static)
        self._slots = []          # static mode: [device arena, ReplayBatch view, "consumer done" event]
        self._turn = 0

    def submit(self, host_batch: "ReplayBatch"):
        if self.static:
            if host_batch._arena is None:
                raise RuntimeError("DevicePrefetcher(static=True) needs an arena batch (ReplayBatch.pin_memory())")
            if len(self._slots) < 2:
                arena = torch.empty_like(host_batch._arena, device=self.device)
                self._slots.append([arena, host_batch._from_arena(arena).mark_static(), None])
            slot = self._slots[self._turn % 2]
            self._turn += 1
            if slot[0].numel() != host_batch._arena.numel() or slot[1]._layout is not host_batch._layout:
                raise RuntimeError("static prefetcher: minibatch layout changed (use DevicePrefetcher(static=False))")
            with torch.cuda.stream(self.stream):
                if slot[2] is not None:
                    self.stream.wait_event(slot[2])      # the step that read this arena two turns ago has finished
                slot[0].copy_(host_batch._arena, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._pending = (slot[1], ev, slot)
            return
        with torch.cuda.stream(self.stream):
            dev = host_batch.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev, None)

    def take(self) -> "ReplayBatch":
        dev, ev, slot = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        if slot is None:
            for t in dev.tensors():          # allocated on the side stream, consumed on the current one
                t.record_stream(cur)
        else:
            self._last = slot
        return dev

    def release(self):
        """Static mode: call after the step that consumed the batch from ``take()`` has been enqueued."""
        slot = getattr(self, "_last", None)
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            slot[2] = ev
            self._last = None


def multistep_lr(base_lr, step, milestones=(500000, 1000000, 1500000), gamma=0.1):
    """optim.lr_scheduler.MultiStepLR as used at airfoil_dqn.py:175-176."""
    return base_lr * gamma ** sum(1 for m in milestones if step >= m)


class _NoTimer:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_TIMER = _NoTimer()


class _Timed:
    def __init__(self, tr, name):
        self.tr, self.name = tr, name

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *a):
        self.e1.record()
        self.tr.timers.setdefault(self.name, []).append((self.e0, self.e1))
        return False


class ReplayTrainer:
    def __init__(self, policy_net_1, policy_net_2, lr=1e-5, weight_decay=1e-6, gamma=1.0, target_update=50,
                 betas=(0.9, 0.999), eps=1e-8, process_group=None, graphs=False, fused_allreduce=True):
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
        self.overlap = True  # segments on three streams (see "one replay step" below); False: everything on the current one
        self._side = None
        self._upd = None
        # early tail launch (opt-in, MDQ_EARLY_TAIL=1 or trainer.early_tail = True): the selected net's tail kernel joins
        # segment H and waits ON THE DEVICE for Q_other (a post enqueued behind the other net's forward) instead of the whole
        # segment T waiting for segment A.  Built to take the tail's forward half (~25 us) off the critical path; measured,
        # it buys nothing -- tools/step_timeline.py shows the two tail kernels slowing each other down when they co-run
        # (23 -> 37 us and 56 -> 70 us): the step is bound by the machine's total issue capacity, not by the dependency
        # chain.  Never on under tools that serialise kernels (ncu, compute-sanitizer: the post could not run
        # while the tail waits; they inject themselves through CUDA_INJECTION64_PATH).
        self.early_tail = os.environ.get("MDQ_EARLY_TAIL") == "1" and not any(
            k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR"))
        self._sync = None
        self.merged_graph = os.environ.get("MDQ_SPLIT_GRAPHS") != "1"
        # graphs=True: the launches of a step (everything but the NCCL all-reduce) are captured once per (select branch,
        # minibatch buffers) and replayed -- the ~13 launches cost ~200 us of host time per step otherwise, more than
        # the kernels.  Needs minibatches at fixed device addresses (the same ReplayBatch, or DevicePrefetcher(static=True)).
        self.graphs = bool(graphs)
        self._graphs = {}
        self._seen = set()
        # world > 1: gradient all-reduce + Adam as ONE kernel over NVLink peer memory (mdq_allreduce_adam) when torch's
        # symmetric memory can map the stage buffers (NCCL all-reduce + Adam otherwise, or with fused_allreduce=False)
        self.fused_allreduce = False
        if self.world > 1 and fused_allreduce and next(policy_net_1.parameters()).is_cuda:
            self._setup_symmetric()

    # -- helpers ----------------------------------------------------------------------------
    def _adam_state(self, i):
        if self._state[i] is None:
            net = self.nets[i]
            net._ensure_packed()
            z = torch.zeros_like(net._flat)
            self._state[i] = dict(m=z, v=z.clone(), g=torch.zeros_like(net._flat), step=0,
                                  step_dev=torch.zeros(2, dtype=torch.int32, device=net._flat.device))
        return self._state[i]

    def _setup_symmetric(self):
        """Map one stage buffer per net on every peer (collective: the same order on every rank)."""
        import ctypes
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else dist.group.WORLD
            rank = dist.get_rank(group)
            for i, net in enumerate(self.nets):
                st = self._adam_state(i)
                n = net._n_used
                stage = symm.empty(int(_lib.lib().mdq_allreduce_stage_floats(n, self.world)), dtype=torch.float32,
                                   device=net._flat.device)
                stage.zero_()
                hdl = symm.rendezvous(stage, group)
                ptrs = (ctypes.c_uint64 * self.world)(*[int(x) for x in hdl.buffer_ptrs])
                st.update(stage=stage, hdl=hdl, peer_ptrs=ptrs, rank=rank,
                          counter=torch.zeros(2, dtype=torch.int32, device=net._flat.device))
            torch.cuda.synchronize()
            dist.barrier(group)                 # every rank's flags are zero before anyone raises one
            self.fused_allreduce = True
        except Exception as err:                # no peer mapping on this box: the NCCL path is the same arithmetic
            import warnings
            warnings.warn(f"fused all-reduce + Adam unavailable ({err}); using NCCL all_reduce + mdq_adam_step_dev")
            self.fused_allreduce = False

    def _timed(self, name):
        return _Timed(self, name) if self.timers is not None else _NO_TIMER

    # -- one replay step ----------------------------------------------------------------------
    # -- one replay step ----------------------------------------------------------------------
    #
    # Segments of a (fused) step and the streams they run on:
    #   H  selected net, stages 0 / 1 (do not read Q_other)             side stream   | beside A
    #   A  forward of the other net -> Q_other                          main stream   |
    #   T  selected net: tail backward, backward 1, weight gradients, loss           main stream, after H and A
    #   U  [all-reduce] + Adam + refresh of the selected net's weight tiles           update stream, after T
    # U of step k is NOT waited for by step k+1's A (the other net is not the one being updated), only by its H: the
    # all-reduce and the optimizer run under the next step's forward.  Every net remembers the event of its last
    # update (`_pending`); its launches wait for it, ReplayTrainer.flush() joins everything to the current stream.

    def _streams(self, dev):
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
            self._upd = torch.cuda.Stream(dev)
        return self._side, self._upd

    def flush(self):
        """Make the current stream wait for every enqueued update (call before reading weights with plain torch ops:
        ``state_dict()``, checkpoints, comparisons)."""
        for n in self.nets:
            ev = n.__dict__.get("_pending")
            if ev is not None:
                torch.cuda.current_stream(n._flat.device).wait_event(ev)

    def _roles(self, batch):
        net1, net2 = self.nets
        s_args = net1._prep(batch.states)
        n_args = net2._prep(batch.next_states) if batch.next_states is not None else None
        sel = 0 if self.select else 1
        if self.select:
            return sel, s_args, 1, batch.next_slot, net2, n_args
        return sel, n_args, 2, batch.owner, net1, s_args

    def _seg_H(self, batch, r, scalar, loss, q_other=None, sync=None):
        """Stages 0 / 1 of the selected net; with ``sync`` also its tail kernel, which waits for ``q_other`` on the device."""
        sel, args, mode, index, other, o_args = r
        st = self._adam_state(sel)
        self.nets[sel]._launch_replay_backward(*args, mode, batch.actions, batch.rewards, index, batch.next_slot, q_other,
                                               int(batch.actions.shape[0]), self.gamma, scalar, loss, st["g"],
                                               phase=1 if sync is None else 3, sync=sync)

    def _seg_A(self, r, out=None, sync=None):
        other, o_args = r[4], r[5]
        if o_args is None:
            return None
        q = other._launch_forward(*o_args, False, False, out=out)[0]
        if sync is not None:
            with torch.cuda.device(q.device):
                _lib.check(_lib.lib().mdq_stream_post(_lib.ptr(sync), _lib.stream_ptr()), "mdq_stream_post")
        return q

    def _early(self, r, dev):
        """The tail may be launched early when both halves exist: a forward of the other net to wait for, and a selected
        net on the staged path (the fused kernel has no split)."""
        sel, args, mode, index, other, o_args = r
        if not self.early_tail or o_args is None:
            return None
        self.nets[sel]._ensure_packed()
        other._ensure_packed()
        if not self.nets[sel]._use_staged(int(args[5]), int(args[6])):
            return None
        if self._sync is None or self._sync.device != dev:
            self._sync = torch.zeros(4, dtype=torch.int32, device=dev)
        return self._sync

    def _seg_T(self, batch, r, q_other, scalar, loss, phase):
        sel, args, mode, index, other, o_args = r
        st = self._adam_state(sel)
        self.nets[sel]._launch_replay_backward(*args, mode, batch.actions, batch.rewards, index, batch.next_slot, q_other,
                                               int(batch.actions.shape[0]), self.gamma, scalar, loss, st["g"], phase=phase)

    def _seg_U(self, sel, refresh=True):
        net, st = self.nets[sel], self._adam_state(sel)
        n_used = net._n_used
        L, p = _lib.lib(), _lib.ptr
        lr = multistep_lr(self.lr, self.num_grads)
        if self.world > 1 and self.fused_allreduce:
            with self._timed("allreduce+adam"):
                rc = L.mdq_allreduce_adam(p(net._flat), p(st["g"]), p(st["m"]), p(st["v"]), n_used, lr, self.betas[0],
                                          self.betas[1], self.eps, self.wd, p(st["step_dev"]), st["peer_ptrs"], st["rank"],
                                          self.world, p(st["counter"]), _lib.stream_ptr())
            _lib.check(rc, "mdq_allreduce_adam")
            if refresh:
                net._staged_refresh(force=True)
            return
        if self.world > 1:
            with self._timed("allreduce"):
                torch.distributed.all_reduce(st["g"][:n_used], group=self.pg)
        # the step count lives on the device (st["step_dev"], advanced by the call) so that the same launch arguments
        # serve every step -- eager and graph-replayed steps share one Adam kernel and one counter
        with self._timed("adam"):
            rc = L.mdq_adam_step_dev(p(net._flat), p(st["g"]), p(st["m"]), p(st["v"]), n_used, lr, self.betas[0],
                                     self.betas[1], self.eps, self.wd, 1.0 / self.world, p(st["step_dev"]),
                                     _lib.stream_ptr())
        _lib.check(rc, "mdq_adam_step_dev")
        if refresh:
            net._staged_refresh(force=True)      # the tiles follow the weights on the update stream, off the next step's path

    def _after_step(self, net, sel, fresh):
        net._bump_weights()               # raw-pointer write: derived weight copies (TF32 hi/lo tiles) are stale now ...
        if fresh:
            net._staged_mark_fresh()      # ... unless segment U rebuilt them right behind the optimizer
        self._state[sel]["step"] += 1
        self.num_grads += 1
        if self.num_grads % self.target_update == 0:
            self.select = not self.select

    @torch.no_grad()
    def step(self, batch: ReplayBatch, fused: bool = True):
        """Returns the Huber loss (device scalar tensor); the selected net's update is enqueued (see ``flush``).

        fused=True (default): forward of the NON-selected net only; the selected net's backward launches recompute its
        own forward and evaluate the Huber term in place.  With ``graphs=True`` and a batch marked static the segments
        are replayed from captured CUDA graphs.
        fused=False: the reference's literal sequence forward(Q1), forward(Q2), Huber, backward (used by tests)."""
        if not fused or self.timers is not None or not self.overlap:
            return self._step_serial(batch, fused)
        static = getattr(batch, "static", False)
        cache = batch.__dict__.setdefault("_step_cache", {}) if static else None
        ck = (id(self), self.select)
        hit = cache.get(ck) if cache is not None else None
        if hit is None:
            r = self._roles(batch)
            hit = (r, self._graph_key(batch, lr=False) if static else None)
            if cache is not None:
                cache[ck] = hit             # a static batch's tensors stay where they are: argument lists and key parts too
        r, key_part = hit
        sel, args, mode, index, other, o_args = r
        if args is None:                    # every transition terminal and the next-state net selected: nothing to train on
            return self._step_serial(batch, fused)
        dev = batch.states.x.device
        net = self.nets[sel]
        main = torch.cuda.current_stream(dev)
        side, upd = self._streams(dev)
        use_graph = self.graphs and static
        entry = None
        if use_graph:
            key = (multistep_lr(self.lr, self.num_grads),) + key_part
            entry = self._graphs.get(key)
            if entry is None:
                if key in self._seen:
                    entry = self._capture(batch, key, r)
                else:
                    if len(self._seen) >= 64:
                        self._seen.clear()
                    self._seen.add(key)
        pend_net, pend_other = net.__dict__.get("_pending"), other.__dict__.get("_pending")
        sync = q_buf = None
        if entry is None:
            scalar = torch.empty(int(args[4]), dtype=torch.float32, device=dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            sync = self._early(r, dev)
            if sync is not None:
                q_buf = torch.empty((int(o_args[4]), other._net.out_dim), dtype=torch.float32, device=dev)
        if entry is not None:
            # graph replay: ONE launch for H || A -> T (the fork to the side stream and the join are inside the captured
            # graph), one for U -- the host issues two launches per step instead of four, which matters because four graph
            # launches cost about as much host time as the step's kernels take
            if pend_net is not None:
                main.wait_event(pend_net)
            if pend_other is not None:
                main.wait_event(pend_other)
            if self.merged_graph:
                entry["HAT"].replay()
            else:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    entry["H"].replay()
                entry["A"].replay()
                main.wait_stream(side)
                entry["T"].replay()
            loss = entry["loss"]
        else:
            inputs_ready = torch.cuda.Event()
            inputs_ready.record(main)
            # kernel by kernel.  A is ENQUEUED first (main stream; only a freshly de-selected net still has an update in
            # flight): an early-launched tail (in H) waits on the device for A's post, and with A already in the queue
            # nothing the host does afterwards -- a first-use cudaMalloc that synchronises the device, an exception -- can
            # keep that post from arriving.
            if pend_other is not None:
                main.wait_event(pend_other)
            q_other = self._seg_A(r, q_buf, sync)
            # H on the side stream: after the inputs and the selected net's previous update, beside A
            side.wait_event(inputs_ready)
            if pend_net is not None:
                side.wait_event(pend_net)
            with torch.cuda.stream(side):
                self._seg_H(batch, r, scalar, loss, q_buf, sync)
                if q_buf is not None:
                    q_buf.record_stream(side)
            main.wait_stream(side)
            self._seg_T(batch, r, q_other, scalar, loss, 2 if sync is None else 4)
        # U on the update stream
        ev = torch.cuda.Event()
        ev.record(main)
        upd.wait_event(ev)
        with torch.cuda.stream(upd):
            if entry is not None:
                entry["U"].replay()
            else:
                self._seg_U(sel)
            done = torch.cuda.Event()
            done.record(upd)
        net._pending = done
        if entry is not None:
            _lib.lib().mdq_launch_count_add(entry["n"])
        self._after_step(net, sel, fresh=True)
        return loss

    def _graph_key(self, batch, lr=True):
        t = [batch.states.x, batch.states.edge_index, batch.actions, batch.rewards, batch.next_slot, batch.owner]
        t += list(graph_ptrs(batch.states)[:2])
        if batch.next_states is not None:
            t += [batch.next_states.x, batch.next_states.edge_index] + list(graph_ptrs(batch.next_states)[:2])
        head = (self.select, multistep_lr(self.lr, self.num_grads)) if lr else (self.select,)
        sizes = tuple(graph_ptrs(batch.states)[2:])           # graph count and size bounds are launch arguments
        if batch.next_states is not None:
            sizes += tuple(graph_ptrs(batch.next_states)[2:])
        return head + sizes + tuple((x.data_ptr(), tuple(x.shape)) for x in t)

    def _capture(self, batch, key, r):
        """Capture the four segments on these minibatch buffers (the second time the buffers show up: minibatches that
        are freshly allocated every step never get here).  Host-side state the kernels would bake into their arguments
        stays out of the graphs: Adam's step count lives on the device, the learning rate and the select branch are part
        of the key.  Capturing records work without running it; the caller replays the segments right away."""
        sel, args, mode, index, other, o_args = r
        dev = batch.states.x.device
        net = self.nets[sel]
        torch.cuda.synchronize(dev)
        if len(self._graphs) >= 128:
            self._graphs.pop(next(iter(self._graphs)))
        other._staged_refresh()
        net._staged_refresh()
        scalar = torch.empty(int(args[4]), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        entry = {"batch": batch, "loss": loss, "scalar": scalar}
        sync = self._early(r, dev)
        q_buf = None
        if sync is not None:
            q_buf = torch.empty((int(o_args[4]), other._net.out_dim), dtype=torch.float32, device=dev)
        entry["q_other"] = q_buf
        L = _lib.lib()
        n0 = int(L.mdq_launch_count())
        pool = torch.cuda.graph_pool_handle()
        side, _ = self._streams(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            cap = torch.cuda.current_stream(dev)        # the capture stream: fork to the side stream, join back
            side.wait_stream(cap)
            with torch.cuda.stream(side):
                self._seg_H(batch, r, scalar, loss, q_buf, sync)
            entry["q_other"] = self._seg_A(r, q_buf, sync)
            cap.wait_stream(side)
            self._seg_T(batch, r, entry["q_other"], scalar, loss, 2 if sync is None else 4)
        entry["HAT"] = g
        if not self.merged_graph:            # the three segments as separate launches (A/B of the host cost)
            for name in ("H", "A", "T"):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    if name == "H":
                        self._seg_H(batch, r, scalar, loss, q_buf, sync)
                    elif name == "A":
                        self._seg_A(r, entry["q_other"], sync)
                    else:
                        self._seg_T(batch, r, entry["q_other"], scalar, loss, 2 if sync is None else 4)
                entry[name] = g
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            self._seg_U(sel)
        entry["U"] = g
        entry["n"] = int(L.mdq_launch_count()) - n0
        L.mdq_launch_count_add(-entry["n"])             # recorded, not run
        net._staged_mark_fresh()                        # the forced refresh inside U was only recorded
        self._graphs[key] = entry
        return entry

    @torch.no_grad()
    def _step_serial(self, batch: ReplayBatch, fused: bool = True):
        """One step, launch by launch on the current stream (timers, overlap off, the unfused reference sequence)."""
        self.flush()
        net1, net2 = self.nets
        dev = batch.states.x.device
        L = _lib.lib()
        p = _lib.ptr
        B = int(batch.actions.shape[0])
        sel, args, mode, index, other, o_args = self._roles(batch)
        s_args = args if self.select else o_args
        n_args = o_args if self.select else args
        n_next = int(n_args[4]) if n_args is not None else 0
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        net = self.nets[sel]
        st = self._adam_state(sel)
        A = net._net.out_dim
        if fused and args is not None:
            scalar = torch.empty(int(args[4]), dtype=torch.float32, device=dev)
            q_other = None
            if o_args is not None:
                with self._timed("qnet_fwd"):
                    q_other = other._launch_forward(*o_args, False, False)[0]
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
                bargs = s_args if self.select else n_args
                with self._timed("qnet_bwd+wgrad"):
                    net._launch_backward(*bargs, gq, st["g"])
            else:
                st["g"].zero_()
        self._seg_U(sel, refresh=False)
        net._pending = None
        self._after_step(net, sel, fresh=False)
        return loss


class DeviceReplayMemory:
    """``ReplayMemory`` (/root/reference/airfoil_dqn.py:48-67: ``push(*Transition)``, ``sample(batch_size)``,
    ``__len__``) with the transitions resident on the device and the minibatch collated by one gather launch.

    The reference keeps a host ``deque`` of PyG ``Data`` objects and every training step runs
    ``DataLoader(batch)`` collation on the host plus a host->device copy of the whole minibatch
    (:240-257).  Here a transition is stored once, into fixed-size device slots (``mdq_replay_store``; the
    states arrive from ``Env2DAirfoil.get_state`` already on the device), and ``sample`` assembles the
    PyG-collated ``ReplayBatch`` with ``mdq_replay_gather``.  The host mirrors only the slots' sizes and
    terminal flags, so it knows every offset / launch parameter of the minibatch without synchronising; what
    crosses PCIe per step is one ~12 KB metadata block instead of the ~10 MB minibatch.

    Ring buffer of ``capacity`` transitions (the oldest is overwritten, as ``deque(maxlen=capacity)`` does);
    ``sample`` draws without replacement like ``random.sample`` (:63-64).  The collated tensors are
    identical to ``ReplayBatch.from_transitions`` of the same transitions in the sampled order.
    """

    def __init__(self, capacity, n_max, e_max, n_features, device):
        self.capacity, self.n_max, self.e_max, self.F = int(capacity), int(n_max), int(e_max), int(n_features)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceReplayMemory lives on a CUDA device (meshdqn_b200 has no CPU path)")
        d = self.device
        f32, i32 = dict(dtype=torch.float32, device=d), dict(dtype=torch.int32, device=d)
        self.x = [torch.zeros((self.capacity, self.n_max, self.F), **f32) for _ in range(2)]      # state, next state
        self.ei = [torch.zeros((self.capacity, 2, max(self.e_max, 1)), **i32) for _ in range(2)]
        self.nn = [torch.zeros(self.capacity, **i32) for _ in range(2)]
        self.ne = [torch.zeros(self.capacity, **i32) for _ in range(2)]
        self.actions = torch.zeros(self.capacity, **i32)
        self.rewards = torch.zeros(self.capacity, **f32)
        import numpy as np
        self._np = np
        self.h_nn = np.zeros((2, self.capacity), dtype=np.int64)     # host mirrors: sizes and terminal flags
        self.h_ne = np.zeros((2, self.capacity), dtype=np.int64)
        self.h_has_next = np.zeros(self.capacity, dtype=bool)
        self.size, self.head = 0, 0
        self._ring, self._ring_pos = [], 0
        self.max_nodes_seen, self.max_edges_seen = 0, 0     # over everything ever stored (upper bounds for static sampling)

    def __len__(self):
        return self.size

    def _store(self, side, slot, data):
        x = data.x
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        x = x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        ei = data.edge_index
        if ei.device != self.device:
            ei = ei.to(self.device, non_blocking=True)
        if ei.dtype != torch.int64 or not ei.is_contiguous():
            ei = ei.to(torch.int64).contiguous()
        n, E = int(x.shape[0]), int(ei.shape[1])
        if x.shape[1] != self.F:
            raise ValueError(f"state has {x.shape[1]} features, the memory was built for {self.F}")
        if n > self.n_max or E > self.e_max:
            raise ValueError(f"graph with {n} nodes / {E} edges exceeds the slot size ({self.n_max} / {self.e_max})")
        L, p = _lib.lib(), _lib.ptr
        with torch.cuda.device(self.device):
            rc = L.mdq_replay_store(p(x), int(x.stride(0)), n, self.F, p(ei) if E else None, E, p(self.x[side]),
                                    p(self.ei[side]), p(self.nn[side]), p(self.ne[side]), slot, self.n_max,
                                    max(self.e_max, 1), _lib.stream_ptr())
        _lib.check(rc, "mdq_replay_store")
        self.h_nn[side, slot], self.h_ne[side, slot] = n, E
        self.max_nodes_seen, self.max_edges_seen = max(self.max_nodes_seen, n), max(self.max_edges_seen, E)

    def push(self, state, action, next_state, reward):
        """One ``Transition(state, action, next_state, reward)`` (airfoil_dqn.py:46-47,58-60); ``next_state`` None = terminal."""
        slot = self.head
        self._store(0, slot, state)
        self.h_has_next[slot] = next_state is not None
        if next_state is not None:
            self._store(1, slot, next_state)
        self.actions[slot] = int(action)
        self.rewards[slot] = float(reward)
        self.head = (self.head + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    # -- minibatch assembly ------------------------------------------------------------------------
    _META_RING = 8          # pinned metadata blocks in flight (a block is reused after its copy's event completed)

    def _meta_block(self, nbytes):
        """[pinned uint8 buffer, event of its last H2D copy] -- a small ring, so a block is only rewritten after the
        copy that read it has completed."""
        if len(self._ring) < self._META_RING:
            self._ring.append([torch.empty(max(nbytes, 4096), dtype=torch.uint8).pin_memory(), None])
            slot = self._ring[-1]
        else:
            slot = self._ring[self._ring_pos % self._META_RING]
            if slot[1] is not None:
                slot[1].synchronize()
            if slot[0].numel() < nbytes:
                slot[0] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self._ring_pos += 1
        return slot

    def sample(self, batch_size, rng=None, idx=None) -> "ReplayBatch":
        """A collated minibatch on the device.  ``idx`` (host ints) fixes the transitions (tests); otherwise
        ``rng`` (a ``numpy.random.RandomState`` / ``Generator``) or numpy's global generator draws them without
        replacement.  Host work: the index draw, five cumulative sums over B integers and one pinned block of
        int64 metadata -> one H2D copy -> one gather launch (which also gathers actions and rewards)."""
        from .data import Batch
        np = self._np
        B = int(batch_size)
        if idx is None:
            if B > self.size:
                raise ValueError(f"sample of {B} from a memory of {self.size}")
            idx = (rng if rng is not None else np.random).choice(self.size, B, replace=False)
        idx = np.asarray(idx, dtype=np.int64)
        B = len(idx)
        has_next = self.h_has_next[idx]
        n_next = int(has_next.sum())
        G2 = max(n_next, 1)
        # metadata block, all int64 words then the int32 arrays: [idx B | ptr B+1 | eptr B+1 | nptr G2+1 | neptr G2+1]
        # then int32 [ptr | eptr | nptr | neptr | next_slot B | owner G2]
        n64 = B + 2 * (B + 1) + 2 * (G2 + 1)
        n32 = 2 * (B + 1) + 2 * (G2 + 1) + B + G2
        nbytes = 8 * n64 + 4 * n32
        slot = self._meta_block(nbytes)
        hb = slot[0].numpy()
        w64 = hb[:8 * n64].view(np.int64)
        w32 = hb[8 * n64:8 * n64 + 4 * n32].view(np.int32)
        o = [0, B, 2 * B + 1, 3 * B + 2, 3 * B + 2 + G2 + 1, n64]
        w64[:B] = idx
        ptr, eptr, nptr, neptr = (w64[o[i]:o[i + 1]] for i in range(1, 5))
        ptr[0] = eptr[0] = nptr[0] = neptr[0] = 0
        np.cumsum(self.h_nn[0, idx], out=ptr[1:])
        np.cumsum(self.h_ne[0, idx], out=eptr[1:])
        nidx = idx[has_next]
        nptr[1:] = 0
        neptr[1:] = 0
        if n_next:
            np.cumsum(self.h_nn[1, nidx], out=nptr[1:])
            np.cumsum(self.h_ne[1, nidx], out=neptr[1:])
        p = [0, B + 1, 2 * B + 2, 2 * B + 2 + G2 + 1, 2 * B + 2 + 2 * (G2 + 1), 2 * B + 2 + 2 * (G2 + 1) + B, n32]
        w32[p[0]:p[1]] = ptr
        w32[p[1]:p[2]] = eptr
        w32[p[2]:p[3]] = nptr
        w32[p[3]:p[4]] = neptr
        w32[p[4]:p[5]] = np.where(has_next, np.cumsum(has_next) - 1, -1)
        w32[p[5]:p[6]] = np.nonzero(has_next)[0] if n_next else 0
        N, E, Nn, En = int(ptr[-1]), int(eptr[-1]), int(nptr[-1]), int(neptr[-1])
        mx = (int(np.diff(ptr).max()), int(np.diff(eptr).max()),
              int(np.diff(nptr).max()) if n_next else 0, int(np.diff(neptr).max()) if n_next else 0)
        d = self.device
        meta = torch.empty(nbytes, dtype=torch.uint8, device=d)
        meta.copy_(slot[0][:nbytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev
        m64 = meta[:8 * n64].view(torch.int64)
        m32 = meta[8 * n64:].view(torch.int32)
        v64 = [m64[o[i]:o[i + 1]] for i in range(5)]
        v32 = [m32[p[i]:p[i + 1]] for i in range(6)]
        x = torch.empty((N, self.F), dtype=torch.float32, device=d)
        ei = torch.empty((2, E), dtype=torch.int64, device=d)
        bvec = torch.empty(N, dtype=torch.int64, device=d)
        act = torch.empty(B, dtype=torch.int32, device=d)
        rew = torch.empty(B, dtype=torch.float32, device=d)
        with_next = n_next > 0
        xn = torch.empty((Nn, self.F), dtype=torch.float32, device=d) if with_next else None
        ein = torch.empty((2, En), dtype=torch.int64, device=d) if with_next else None
        bn = torch.empty(Nn, dtype=torch.int64, device=d) if with_next else None
        L, q = _lib.lib(), _lib.ptr
        with torch.cuda.device(d):
            rc = L.mdq_replay_gather(q(self.x[0]), q(self.ei[0]), q(self.nn[0]), q(self.ne[0]), q(self.x[1]), q(self.ei[1]),
                                     q(self.nn[1]), q(self.ne[1]), self.n_max, max(self.e_max, 1), self.F, q(v64[0]), B,
                                     q(v32[0]), q(v32[1]), E, q(x), q(ei), q(bvec), q(v32[4]), q(v32[2]), q(v32[3]), En,
                                     q(xn), q(ein), q(bn), q(self.actions), q(self.rewards), q(act), q(rew),
                                     _lib.stream_ptr())
        _lib.check(rc, "mdq_replay_gather")

        def mk(xx, ee, bb, p64, e64, p32, e32, G, mn, me):
            b = Batch(x=xx, edge_index=ee)
            b.batch, b.ptr, b.eptr, b.num_graphs = bb, p64, e64, G
            b.__dict__["_mdq_ptrs"] = (p32, e32, G, mn, me)
            b.__dict__["_meta"] = b.__dict__["_mdq_ptrs"]
            return b
        states = mk(x, ei, bvec, v64[1], v64[2], v32[0], v32[1], B, mx[0], mx[1])
        nexts = mk(xn, ein, bn, v64[3][:n_next + 1], v64[4][:n_next + 1], v32[2][:n_next + 1], v32[3][:n_next + 1],
                   n_next, mx[2], mx[3]) if with_next else None
        out = ReplayBatch(states, act, nexts, v32[4], rew, v32[5])
        out._keep = meta                  # the metadata block backs the offset views
        return out

    def static_sampler(self, batch_size):
        """A sampler whose minibatches live in FIXED device buffers (see ``StaticSampler``)."""
        return StaticSampler(self, batch_size)


class StaticSampler:
    """Minibatches of a ``DeviceReplayMemory`` at fixed device addresses and fixed tensor shapes, so that
    ``ReplayTrainer(graphs=True)`` replays its captured step on them: the step's launch arguments then depend only on the
    number of non-terminal transitions in the draw (the next-state batch's graph count), which becomes part of the graph
    key -- a few dozen values at B = 256, each captured on its second sighting.

    Buffers are sized for the worst case (B graphs of ``n_max`` nodes / ``e_max`` edges; edge rows ``B * e_max`` apart);
    the kernels read sizes from the offset vectors, never from tensor shapes.  ``sample`` overwrites the previous
    minibatch in stream order: enqueue it on the stream the trainer steps on, after the step that used the last one.
    Per draw the host does what ``DeviceReplayMemory.sample`` does (index draw, cumulative sums, one pinned block, one
    H2D copy of ~12 KB, one gather launch) but allocates nothing."""

    def __init__(self, mem: "DeviceReplayMemory", batch_size):
        self.mem, self.B = mem, int(batch_size)
        B, d = self.B, mem.device
        N, E = B * mem.n_max, B * max(mem.e_max, 1)
        f32, i64, i32 = dict(dtype=torch.float32, device=d), dict(dtype=torch.int64, device=d), dict(dtype=torch.int32, device=d)
        self.x = [torch.zeros((N, mem.F), **f32) for _ in range(2)]
        self.ei = [torch.zeros((2, E), **i64) for _ in range(2)]
        self.bvec = [torch.zeros(N, **i64) for _ in range(2)]
        self.act, self.rew = torch.zeros(B, **i32), torch.zeros(B, **f32)
        self.n64 = B + 4 * (B + 1)
        self.n32 = 4 * (B + 1) + 2 * B
        self.meta = torch.zeros(8 * self.n64 + 4 * self.n32, dtype=torch.uint8, device=d)
        self.E = E

    def sample(self, rng=None, idx=None) -> "ReplayBatch":
        from .data import Batch
        mem, B = self.mem, self.B
        np = mem._np
        if idx is None:
            if B > mem.size:
                raise ValueError(f"sample of {B} from a memory of {mem.size}")
            idx = (rng if rng is not None else np.random).choice(mem.size, B, replace=False)
        idx = np.asarray(idx, dtype=np.int64)
        if len(idx) != B:
            raise ValueError(f"this sampler draws minibatches of {B}")
        has_next = mem.h_has_next[idx]
        n_next = int(has_next.sum())
        n64, n32 = self.n64, self.n32
        nbytes = 8 * n64 + 4 * n32
        slot = mem._meta_block(nbytes)
        hb = slot[0].numpy()
        w64 = hb[:8 * n64].view(np.int64)
        w32 = hb[8 * n64:nbytes].view(np.int32)
        G = B + 1                                     # fixed section sizes: every offset vector has room for B graphs
        w64[:B] = idx
        ptr, eptr, nptr, neptr = (w64[B + i * G:B + (i + 1) * G] for i in range(4))
        ptr[0] = eptr[0] = 0
        nptr[:] = 0
        neptr[:] = 0
        np.cumsum(mem.h_nn[0, idx], out=ptr[1:])
        np.cumsum(mem.h_ne[0, idx], out=eptr[1:])
        if n_next:
            nidx = idx[has_next]
            np.cumsum(mem.h_nn[1, nidx], out=nptr[1:n_next + 1])
            np.cumsum(mem.h_ne[1, nidx], out=neptr[1:n_next + 1])
            nptr[n_next + 1:] = nptr[n_next]
            neptr[n_next + 1:] = neptr[n_next]
        for i, v in enumerate((ptr, eptr, nptr, neptr)):
            w32[i * G:(i + 1) * G] = v
        w32[4 * G:4 * G + B] = np.where(has_next, np.cumsum(has_next) - 1, -1)      # next_slot
        w32[4 * G + B:4 * G + 2 * B] = 0
        if n_next:
            w32[4 * G + B:4 * G + B + n_next] = np.nonzero(has_next)[0]              # owner
        self.meta.copy_(slot[0][:nbytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev
        m64 = self.meta[:8 * n64].view(torch.int64)
        m32 = self.meta[8 * n64:].view(torch.int32)
        v64 = [m64[:B]] + [m64[B + i * G:B + (i + 1) * G] for i in range(4)]
        v32 = [m32[i * G:(i + 1) * G] for i in range(4)] + [m32[4 * G:4 * G + B], m32[4 * G + B:4 * G + 2 * B]]
        with_next = n_next > 0
        L, q = _lib.lib(), _lib.ptr
        with torch.cuda.device(mem.device):
            rc = L.mdq_replay_gather(q(mem.x[0]), q(mem.ei[0]), q(mem.nn[0]), q(mem.ne[0]), q(mem.x[1]), q(mem.ei[1]),
                                     q(mem.nn[1]), q(mem.ne[1]), mem.n_max, max(mem.e_max, 1), mem.F, q(v64[0]), B,
                                     q(v32[0]), q(v32[1]), self.E, q(self.x[0]), q(self.ei[0]), q(self.bvec[0]), q(v32[4]),
                                     q(v32[2]), q(v32[3]), self.E, q(self.x[1]) if with_next else None,
                                     q(self.ei[1]) if with_next else None, q(self.bvec[1]) if with_next else None,
                                     q(mem.actions), q(mem.rewards), q(self.act), q(self.rew), _lib.stream_ptr())
        _lib.check(rc, "mdq_replay_gather")
        mn, me = max(mem.max_nodes_seen, 1), max(mem.max_edges_seen, 1)

        def mk(side, p64, e64, p32, e32, Gn):
            b = Batch(x=self.x[side], edge_index=self.ei[side])
            b.batch, b.ptr, b.eptr, b.num_graphs = self.bvec[side], p64, e64, Gn
            b.__dict__["_mdq_ptrs"] = (p32, e32, Gn, mn, me)
            b.__dict__["_meta"] = b.__dict__["_mdq_ptrs"]
            return b
        states = mk(0, v64[1], v64[2], v32[0], v32[1], B)
        nexts = mk(1, v64[3][:n_next + 1], v64[4][:n_next + 1], v32[2][:n_next + 1], v32[3][:n_next + 1], n_next) if with_next else None
        out = ReplayBatch(states, self.act, nexts, v32[4], self.rew, v32[5][:max(n_next, 1)])
        out.static = True
        return out
