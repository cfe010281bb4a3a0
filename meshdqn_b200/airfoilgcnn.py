"""Host-side mirror of the reference's ``airfoilgcnn.py`` on the fused sm_100a kernels.

Same classes, constructor arguments, method names and ``state_dict`` keys as
/root/reference/airfoilgcnn.py:24-209:

* ``NodeRemovalNet(output_dim, conv_width=64, topk=0.5, initial_num_nodes=None)`` with
  ``reset()``, ``set_num_nodes(n)``, ``set_removable(r)``, ``forward(data, embedding=False)``
* ``AirfoilGCNN(conv_width=64)`` with ``forward(data)``

plus the parameter-server helpers the reference's callers expect but never define
(``get_weights/set_weights/get_gradients/set_gradients``, airfoil_dqn.py:194-206,291-310).

All parameters alias ONE flat fp32 device buffer laid out for the kernels (dense weights
transposed, see include/meshdqn_b200.h); the ``nn.Parameter`` objects are strided views into
it, so ``state_dict()`` / ``load_state_dict()`` / optimizers see the usual PyG-shaped tensors.
``forward`` is one kernel launch (mdq_qnet_forward); autograd backward is
mdq_qnet_backward (recompute + deterministic split-K weight gradients).  There is no CPU
path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn

from . import _lib
from ._lib import MDQ_BLOCK_GCN, MDQ_BLOCK_SAGE, mdq_net_t

if torch.cuda.is_available():
    device = torch.device("cuda:0")
else:  # the reference prints and falls back to CPU; here CPU tensors are only legal for state_dict I/O
    device = torch.device("cpu")


class _on_device:
    """``torch.cuda.device(dev)`` only when `dev` is not already current (the context manager costs ~10 us of host time)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


class _Lin(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            self.bias = nn.Parameter(torch.empty(o))
            bound = 1 / math.sqrt(i) if i > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)
        else:
            self.register_parameter("bias", None)


class SAGEConv(nn.Module):
    """Parameter holder with PyG's SAGEConv names (``lin_l`` with bias, ``lin_r`` without)."""

    def __init__(self, i, o):
        super().__init__()
        self.in_channels, self.out_channels = i, o
        self.lin_l = _Lin(i, o, bias=True)
        self.lin_r = _Lin(i, o, bias=False)


class GCNConv(nn.Module):
    """Parameter holder with PyG's GCNConv names (``lin.weight``, ``bias``)."""

    def __init__(self, i, o):
        super().__init__()
        self.in_channels, self.out_channels = i, o
        self.lin = _Lin(i, o, bias=False)
        self.bias = nn.Parameter(torch.zeros(o))


class TopKPooling(nn.Module):
    def __init__(self, c, ratio):
        super().__init__()
        self.ratio = ratio
        self.weight = nn.Parameter(torch.empty(1, c))
        bound = 1.0 / math.sqrt(c)
        nn.init.uniform_(self.weight, -bound, bound)


def graph_ptrs(data):
    """(node_ptr i32 [B+1], edge_ptr i32 [B+1], B, max_n, max_e) on data.x's device.

    ``Batch.from_data_list`` precomputes these; for a bare ``Data`` (batch None) they are trivial;
    for a user-made ``batch`` vector they are derived with torch ops.
    """
    cached = getattr(data, "_mdq_ptrs", None)
    dev = data.x.device
    if cached is not None and cached[0].device == dev:
        return cached
    N = int(data.x.shape[0])
    E = int(data.edge_index.shape[1]) if data.edge_index is not None else 0
    batch = getattr(data, "batch", None)
    if batch is None:
        out = (torch.tensor([0, N], dtype=torch.int32, device=dev), torch.tensor([0, E], dtype=torch.int32, device=dev),
               1, N, E)
    elif getattr(data, "ptr", None) is not None and getattr(data, "eptr", None) is not None:
        ptr, eptr = data.ptr.cpu(), data.eptr.cpu()
        out = (ptr.to(torch.int32).to(dev), eptr.to(torch.int32).to(dev), len(ptr) - 1,
               int((ptr[1:] - ptr[:-1]).max()), int((eptr[1:] - eptr[:-1]).max()) if len(eptr) > 1 else 0)
    else:
        B = int(batch.max().item()) + 1 if N else 0
        ncount = torch.bincount(batch, minlength=B)
        ecount = torch.bincount(batch[data.edge_index[0]], minlength=B) if E else torch.zeros(B, dtype=torch.long, device=dev)
        ptr = torch.zeros(B + 1, dtype=torch.long, device=dev)
        eptr = torch.zeros(B + 1, dtype=torch.long, device=dev)
        ptr[1:] = ncount.cumsum(0)
        eptr[1:] = ecount.cumsum(0)
        out = (ptr.to(torch.int32), eptr.to(torch.int32), B, int(ncount.max().item()), int(ecount.max().item()))
    try:
        data._mdq_ptrs = out
    except Exception:
        pass
    return out


class _QNetFunction(torch.autograd.Function):
    """out = Q(data); backward = mdq_qnet_backward.  The parameter tensors are inputs only so that
    autograd routes gradients to them; the kernels read/write the flat buffers."""

    @staticmethod
    def forward(ctx, module, x, edge_index, nptr, eptr, B, max_n, max_e, embedding, *params):
        out, emb, _ = module._launch_forward(x, edge_index, nptr, eptr, B, max_n, max_e, embedding, False)
        ctx.module = module
        ctx.meta = (B, max_n, max_e)
        ctx.embedding = embedding
        ctx.save_for_backward(x, edge_index, nptr, eptr)
        return emb if embedding else out

    @staticmethod
    def backward(ctx, gout):
        if ctx.embedding:
            raise NotImplementedError("backward through embedding=True is not supported")
        m = ctx.module
        x, edge_index, nptr, eptr = ctx.saved_tensors
        B, max_n, max_e = ctx.meta
        flat_grad = torch.empty_like(m._flat)
        m._launch_backward(x, edge_index, nptr, eptr, B, max_n, max_e, gout.contiguous().float(), flat_grad)
        grads = tuple(m._grad_view(flat_grad, name) if (p.requires_grad and name not in m._unused) else None
                      for name, p in m._named_flat_params())
        return (None,) * 9 + grads


class _FusedQNet(nn.Module):
    """Shared machinery: flat parameter buffer + kernel launches."""

    _blocks_used = ()
    _softmax = 1
    _in_col0 = 0

    # -- flat buffer ------------------------------------------------------------------------
    def _layout(self):
        """[(name, param, offset, rows(K), cols(C), transposed)] and the mdq_net_t descriptor."""
        W = self.conv_width
        entries = []
        off = 0

        def take(n):
            nonlocal off
            at = off
            off += (n + 3) // 4 * 4
            return at

        net = mdq_net_t()
        used = list(self._blocks_used)
        net.n_blocks = len(used)
        net.width = W
        net.ratio = float(self.topk)
        net.softmax = self._softmax
        net.in_col0 = self._in_col0
        unused = [k for k in range(1, 7) if k not in used]

        def add_block(bi, k):
            conv, pool = getattr(self, f"conv{k}"), getattr(self, f"pool{k}")
            kin = conv.in_channels
            if isinstance(conv, SAGEConv):
                w_off = take(2 * kin * W)
                entries.append((f"conv{k}.lin_l.weight", conv.lin_l.weight, w_off, kin, W, True))
                entries.append((f"conv{k}.lin_r.weight", conv.lin_r.weight, w_off + kin * W, kin, W, True))
                b_off = take(W)
                entries.append((f"conv{k}.lin_l.bias", conv.lin_l.bias, b_off, 1, W, False))
                typ = MDQ_BLOCK_SAGE
            else:
                w_off = take(kin * W)
                entries.append((f"conv{k}.lin.weight", conv.lin.weight, w_off, kin, W, True))
                b_off = take(W)
                entries.append((f"conv{k}.bias", conv.bias, b_off, 1, W, False))
                typ = MDQ_BLOCK_GCN
            p_off = take(W)
            entries.append((f"pool{k}.weight", pool.weight, p_off, 1, W, False))
            if bi is not None:
                blk = net.blk[bi]
                blk.type, blk.kin, blk.w_off, blk.b_off, blk.pool_off = typ, kin, w_off, b_off, p_off
                if bi == 0:
                    net.in_dim = kin

        for bi, k in enumerate(used):
            add_block(bi, k)
        for i, lin in enumerate((self.lin1, self.lin2, self.lin3)):
            o, k = lin.weight.shape
            w_off = take(k * o)
            entries.append((f"lin{i + 1}.weight", lin.weight, w_off, k, o, True))
            b_off = take(o)
            entries.append((f"lin{i + 1}.bias", lin.bias, b_off, 1, o, False))
            net.lin_off[i], net.lin_boff[i], net.lin_in[i], net.lin_out[i] = w_off, b_off, k, o
        self._n_used = off  # floats the forward reads; unused blocks (kept for state_dict parity) follow
        for k in unused:
            add_block(None, k)
        net.out_dim = self.lin3.weight.shape[0]
        net.n_params = off
        return entries, net

    def _pack(self):
        entries, net = self._layout()
        dev = entries[0][1].device
        flat = torch.zeros(net.n_params, dtype=torch.float32, device=dev)
        views = {}
        with torch.no_grad():
            for name, p, off, K, C, tr in entries:
                v = flat[off:off + K * C].view(K, C)
                v = v.t() if tr else v.view(p.shape)
                v.copy_(p.data.to(torch.float32))
                p.data = v
                views[name] = (off, K, C, tr, tuple(p.shape))
        self._flat = flat
        self._views = views
        self._entries = [(n, p) for n, p, *_ in entries]
        used = set(self._blocks_used)
        # blocks the forward never touches get no gradient (torch leaves .grad = None for them)
        self._unused = {n for n, _ in self._entries
                        if n[:4] in ("conv", "pool") and int(n[4]) not in used}
        self._net = net
        self._packed_ptr = flat.data_ptr()
        self._ws = None
        self._wsplit = None          # derived weight copies are rebuilt from the new flat buffer
        self._wgen = getattr(self, "_wgen", 0) + 1

    def _ensure_packed(self):
        flat = getattr(self, "_flat", None)
        if flat is not None:
            name, p = self._entries[0]
            if p.data_ptr() == flat.data_ptr() + 4 * self._views[name][0] and p.device == flat.device:
                return
        self._pack()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._flat = None
        return out

    def _named_flat_params(self):
        return self._entries

    def _weights_version(self):
        """Changes whenever the flat buffer may hold different numbers: in-place torch writes bump the parameters'
        own version counters (``load_state_dict``, ``set_weights``, torch optimizers -- the flat tensor's counter does
        NOT move, the parameters are separate views), raw-pointer writers (``mdq_adam_step``) call
        ``_bump_weights()``, a re-pack allocates a new buffer."""
        return (self._flat.data_ptr(), self._wgen, sum(p._version for _, p in self._entries))

    def _bump_weights(self):
        self._wgen = getattr(self, "_wgen", 0) + 1

    def _grad_view(self, flat_grad, name):
        off, K, C, tr, shape = self._views[name]
        v = flat_grad[off:off + K * C].view(K, C)
        return v.t() if tr else v.view(shape)

    # -- reference helper API (airfoil_dqn.py:194-206,291-310) ----------------------------------
    def get_weights(self):
        return {k: v.detach().cpu() for k, v in self.state_dict().items()}

    def set_weights(self, weights):
        self.load_state_dict(weights)

    def get_gradients(self):
        return [None if p.grad is None else p.grad.detach().cpu().numpy() for p in self.parameters()]

    def set_gradients(self, gradients):
        for g, p in zip(gradients, self.parameters()):
            if g is not None:
                p.grad = torch.as_tensor(g, dtype=p.dtype, device=p.device).reshape(p.shape).clone()

    # -- launches ---------------------------------------------------------------------------
    def _prep(self, data):
        x = data.x
        if not x.is_cuda:
            raise RuntimeError("meshdqn_b200 has no CPU path: move the data to a CUDA device (data.to('cuda'))")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        ei = data.edge_index
        if ei.dtype == torch.int32 and ei.is_contiguous():
            pass        # 32-bit edges (ReplayBatch arenas): read as they are by the staged kernels, widened for the others
        elif ei.dtype != torch.int64 or not ei.is_contiguous():
            ei = ei.to(torch.int64).contiguous()
        nptr, eptr, B, max_n, max_e = graph_ptrs(data)
        return x, ei, nptr, eptr, B, max_n, max_e

    @staticmethod
    def _edge_ptrs(ei):
        """(row 0 pointer, row 1 pointer, is_int32) of a contiguous [2, E] edge_index."""
        E = int(ei.shape[1])
        i32 = ei.dtype == torch.int32
        base = ei.data_ptr()
        return _lib.c_void_p(base), _lib.c_void_p(base + (4 if i32 else 8) * E), 1 if i32 else 0

    # -- staged tensor-core path (csrc/gnn_staged.cuh): tcgen05 3xTF32 GEMMs per stage instead of one CTA per graph ----
    qpath = os.environ.get("MDQ_QPATH", "auto")     # "auto": staged when the net / graph sizes allow it | "fused" | "staged"

    def _use_staged(self, max_n, max_e):
        if self.qpath == "fused":
            return False
        cache = self.__dict__.setdefault("_stg_ok", {})
        key = (int(max_n), int(max_e), id(self._net))
        ok = cache.get(key)
        if ok is None:
            ok = cache[key] = bool(_lib.lib().mdq_qnet_staged_supported(self._net, int(max_n), int(max_e)))
        if not ok and self.qpath == "staged":
            raise RuntimeError("qpath='staged': this network / graph size is not served by the staged kernels")
        return ok

    def _staged_wsplit(self):
        """hi / lo TF32 halves of every weight matrix in the staged kernels' tile layout; rebuilt (one launch) whenever
        the weights changed."""
        ver = self._weights_version()
        if getattr(self, "_stg_w", None) is None or self._stg_w.device != self._flat.device or self._stg_wver != ver:
            L = _lib.lib()
            n = int(L.mdq_qnet_staged_wsplit_floats(self._net))
            if getattr(self, "_stg_w", None) is None or self._stg_w.numel() != n or self._stg_w.device != self._flat.device:
                self._stg_w = torch.empty(n, dtype=torch.float32, device=self._flat.device)
            with _on_device(self._flat.device):
                rc = L.mdq_qnet_staged_wsplit(self._net, _lib.ptr(self._flat), _lib.ptr(self._stg_w), _lib.stream_ptr())
            _lib.check(rc, "mdq_qnet_staged_wsplit")
            self._stg_wver = ver
        return self._stg_w

    def _staged_refresh(self, force=False):
        """Bring the staged path's derived weights up to date now (no-op when they are, or when the net never ran staged).
        ``force``: launch the refresh whatever the host's bookkeeping says (inside a captured graph)."""
        if getattr(self, "_stg_w", None) is not None and getattr(self, "_flat", None) is not None:
            if force:
                self._stg_wver = None
            self._staged_wsplit()

    def _staged_mark_fresh(self):
        if getattr(self, "_stg_w", None) is not None:
            self._stg_wver = self._weights_version()

    def state_dict(self, *args, **kwargs):
        flat = getattr(self, "_flat", None)
        if flat is not None and flat.is_cuda:
            self._wait_pending(flat.device)      # a trainer's update of these weights may still be in flight
        return super().state_dict(*args, **kwargs)

    def _wait_pending(self, dev):
        """A trainer may still have this net's update in flight on its update stream (ReplayTrainer, segment U)."""
        ev = self.__dict__.get("_pending")
        if ev is not None:
            if ev.query():                       # already finished: nothing to order against (and a capture in
                self.__dict__["_pending"] = None  # progress could not wait on uncaptured work anyway)
            else:
                torch.cuda.current_stream(dev).wait_event(ev)

    def _staged_ws(self, B, max_n, max_e, backward, dev, shared=False):
        """Workspace of the staged launches, one per (stream, direction): replicas share a net across streams.
        ``shared``: one workspace whatever the stream (the two phases of the replay backward run on two streams)."""
        sizes = self.__dict__.setdefault("_stg_need", {})
        skey = (B, max_n, max_e, bool(backward))
        need = sizes.get(skey)
        if need is None:
            need = sizes[skey] = int(_lib.lib().mdq_qnet_staged_workspace_floats(self._net, B, max_n, max_e, 1 if backward else 0))
        if need < 0:
            raise RuntimeError("mdq_qnet_staged_workspace_floats failed")
        if getattr(self, "_stg_wss", None) is None:
            self._stg_wss = {}
        key = (0 if shared else torch._C._cuda_getCurrentRawStream(dev.index), bool(backward), dev.index)
        ws = self._stg_wss.get(key)
        if ws is None or ws.numel() < need:
            ws = self._stg_wss[key] = torch.empty(need, dtype=torch.float32, device=dev)
        return ws

    def _launch_forward(self, x, ei, nptr, eptr, B, max_n, max_e, embedding, want_argmax, out=None):
        self._ensure_packed()
        if self._flat.device != x.device:
            raise RuntimeError(f"network on {self._flat.device}, data on {x.device}")
        net = self._net
        net.x_stride = int(x.shape[1])
        if net.in_col0 + net.in_dim > net.x_stride:
            raise ValueError(f"data.x has {net.x_stride} columns, network expects >= {net.in_col0 + net.in_dim}")
        A = net.out_dim
        if out is None:
            out = torch.empty((B, A), dtype=torch.float32, device=x.device)
        emb = torch.empty((B, 2 * net.width), dtype=torch.float32, device=x.device) if embedding else None
        am = torch.empty((B,), dtype=torch.int32, device=x.device) if want_argmax else None
        E = int(ei.shape[1])
        L = _lib.lib()
        if self._use_staged(max_n, max_e):
            wsp = self._staged_wsplit()
            ws = self._staged_ws(B, max_n, max_e, False, x.device)
            e0, e1, i32 = self._edge_ptrs(ei)
            with _on_device(x.device):
                rc = L.mdq_qnet_staged_forward(net, _lib.ptr(self._flat), _lib.ptr(wsp), _lib.ptr(x), e0, e1, i32,
                                               _lib.ptr(nptr), _lib.ptr(eptr), B, max_n, max_e, _lib.ptr(out),
                                               _lib.ptr(emb), _lib.ptr(am), _lib.ptr(ws), _lib.stream_ptr())
            _lib.check(rc, "mdq_qnet_staged_forward")
            return out, emb, am
        if ei.dtype != torch.int64:
            ei = ei.long()
        with torch.cuda.device(x.device):
            rc = L.mdq_qnet_forward(net, _lib.ptr(self._flat), _lib.ptr(x), _lib.c_void_p(ei.data_ptr()),
                                    _lib.c_void_p(ei.data_ptr() + 8 * E), _lib.ptr(nptr), _lib.ptr(eptr), B, max_n, max_e,
                                    _lib.ptr(out), _lib.ptr(emb), _lib.ptr(am), _lib.stream_ptr())
        _lib.check(rc, "mdq_qnet_forward")
        return out, emb, am

    def _launch_backward(self, x, ei, nptr, eptr, B, max_n, max_e, gout, flat_grad):
        self._ensure_packed()
        net = self._net
        net.x_stride = int(x.shape[1])
        L = _lib.lib()
        if self._use_staged(max_n, max_e):
            wsp = self._staged_wsplit()
            ws = self._staged_ws(B, max_n, max_e, True, x.device)
            e0, e1, i32 = self._edge_ptrs(ei)
            with _on_device(x.device):
                rc = L.mdq_qnet_staged_backward(net, _lib.ptr(self._flat), _lib.ptr(wsp), _lib.ptr(x), e0, e1, i32,
                                                _lib.ptr(nptr), _lib.ptr(eptr), B, max_n, max_e, _lib.ptr(gout),
                                                _lib.ptr(flat_grad), _lib.ptr(ws), _lib.stream_ptr())
            _lib.check(rc, "mdq_qnet_staged_backward")
            return
        if ei.dtype != torch.int64:
            ei = ei.long()
        need = int(L.mdq_qnet_bwd_workspace_floats(net, B, max_n))
        if need < 0:
            raise RuntimeError("mdq_qnet_bwd_workspace_floats failed")
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.float32, device=x.device)
        E = int(ei.shape[1])
        with torch.cuda.device(x.device):
            rc = L.mdq_qnet_backward(net, _lib.ptr(self._flat), _lib.ptr(x), _lib.c_void_p(ei.data_ptr()),
                                     _lib.c_void_p(ei.data_ptr() + 8 * E), _lib.ptr(nptr), _lib.ptr(eptr), B, max_n,
                                     max_e, _lib.ptr(gout), _lib.ptr(flat_grad), _lib.ptr(self._ws), _lib.stream_ptr())
        _lib.check(rc, "mdq_qnet_backward")

    def _launch_replay_backward(self, x, ei, nptr, eptr, B, max_n, max_e, mode, action, reward, index, next_slot, q_other,
                                batch, gamma, scalar, loss, flat_grad, phase=0, sync=None):
        self._ensure_packed()
        net = self._net
        net.x_stride = int(x.shape[1])
        L = _lib.lib()
        if self._use_staged(max_n, max_e):
            wsp = self._staged_wsplit()
            ws = self._staged_ws(B, max_n, max_e, True, x.device, shared=True)
            e0, e1, i32 = self._edge_ptrs(ei)
            with _on_device(x.device):
                rc = L.mdq_qnet_staged_replay_backward(net, _lib.ptr(self._flat), _lib.ptr(wsp), _lib.ptr(x), e0, e1, i32,
                                                       _lib.ptr(nptr), _lib.ptr(eptr), B, max_n, max_e, int(mode),
                                                       _lib.ptr(action), _lib.ptr(reward), _lib.ptr(index),
                                                       _lib.ptr(next_slot), _lib.ptr(q_other), int(batch), float(gamma),
                                                       _lib.ptr(scalar), _lib.ptr(loss), _lib.ptr(flat_grad), _lib.ptr(ws),
                                                       int(phase), _lib.ptr(sync), _lib.stream_ptr())
            _lib.check(rc, "mdq_qnet_staged_replay_backward")
            return
        if phase in (3, 4) or sync is not None:
            raise RuntimeError("phases 3 / 4 (early tail launch) exist on the staged path only")
        if phase == 1:
            return      # the fused kernel has no split: everything happens in the finishing call
        if ei.dtype != torch.int64:
            ei = ei.long()

        need = int(L.mdq_qnet_bwd_workspace_floats(net, B, max_n))
        if need < 0:
            raise RuntimeError("mdq_qnet_bwd_workspace_floats failed")
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.float32, device=x.device)
        E = int(ei.shape[1])
        with torch.cuda.device(x.device):
            rc = L.mdq_qnet_replay_backward(net, _lib.ptr(self._flat), _lib.ptr(x), _lib.c_void_p(ei.data_ptr()),
                                            _lib.c_void_p(ei.data_ptr() + 8 * E), _lib.ptr(nptr), _lib.ptr(eptr), B, max_n,
                                            max_e, int(mode), _lib.ptr(action), _lib.ptr(reward), _lib.ptr(index),
                                            _lib.ptr(next_slot), _lib.ptr(q_other), int(batch), float(gamma),
                                            _lib.ptr(scalar), _lib.ptr(loss), _lib.ptr(flat_grad), _lib.ptr(self._ws),
                                            _lib.stream_ptr())
        _lib.check(rc, "mdq_qnet_replay_backward")

    # -- layered path: one large graph (does not fit a CTA's shared memory) -----------------------------------
    FUSED_MAX_NODES = 2048          # above this a single graph goes layer by layer (gnn_layered.cu)
    layered_gemm = "tf32x3"         # "tf32x3": tcgen05 tensor cores (3xTF32, ~fp32 accuracy) | "fp32": FFMA

    def _pack_wsplit(self):
        """hi/lo TF32 split of the conv weights, tiled as K-major UMMA core matrices (include/meshdqn_b200.h)."""
        net, flat = self._net, self._flat
        ws = torch.zeros(3 * net.n_params, dtype=torch.float32, device=flat.device)
        W = net.width
        for bi in range(net.n_blocks):
            b = net.blk[bi]
            if b.type == MDQ_BLOCK_SAGE:
                # A rows are [x (Fp) | mean (Fp)]: lin_r^T goes to rows 0..F-1, lin_l^T to rows Fp..Fp+F-1
                F = b.kin
                Fp = (F + 3) // 4 * 4
                kpad = (2 * Fp + 7) // 8 * 8
                w = torch.zeros(kpad, W, dtype=torch.float32, device=flat.device)
                wl = flat[b.w_off:b.w_off + 2 * F * W].view(2 * F, W)
                w[:F] = wl[F:]
                w[Fp:Fp + F] = wl[:F]
            else:
                K = b.kin
                kpad = (K + 7) // 8 * 8
                w = torch.zeros(kpad, W, dtype=torch.float32, device=flat.device)
                w[:K] = flat[b.w_off:b.w_off + K * W].view(K, W)
            hi = (w.view(torch.int32) & -8192).view(torch.float32)          # 0xffffe000
            lo = w - hi

            def tile(m):   # [kpad][W] -> [kpad/4][W/8][8][4], element [c][g][r][kk] = m[4c+kk][8g+r]
                return m.view(kpad // 4, 4, W // 8, 8).permute(0, 2, 3, 1).contiguous().view(-1)
            o = 3 * b.w_off
            ws[o:o + kpad * W] = tile(hi)
            ws[o + kpad * W:o + 2 * kpad * W] = tile(lo)
        self._wsplit = ws
        self._wsplit_version = self._weights_version()

    def _launch_forward_layered(self, x, ei, embedding, want_argmax):
        self._ensure_packed()
        if ei.dtype != torch.int64:
            ei = ei.long()
        net = self._net
        net.x_stride = int(x.shape[1])
        N, E = int(x.shape[0]), int(ei.shape[1])
        mode = 1 if (self.layered_gemm == "tf32x3" and net.width == 128) else 0
        if mode == 1 and (getattr(self, "_wsplit", None) is None or self._wsplit.device != x.device
                          or self._wsplit_version != self._weights_version()):
            self._pack_wsplit()
        L = _lib.lib()
        need = int(L.mdq_qnet_layered_workspace_bytes(net, N, E))
        if need < 0:
            raise RuntimeError("mdq_qnet_layered_workspace_bytes failed")
        lw = getattr(self, "_layered_ws", None)
        if lw is None or lw.numel() < need or lw.device != x.device:
            self._layered_ws = lw = torch.empty(need, dtype=torch.uint8, device=x.device)
        out = torch.empty((1, net.out_dim), dtype=torch.float32, device=x.device)
        emb = torch.empty((1, 2 * net.width), dtype=torch.float32, device=x.device) if embedding else None
        am = torch.empty((1,), dtype=torch.int32, device=x.device) if want_argmax else None
        with torch.cuda.device(x.device):
            rc = L.mdq_qnet_forward_layered(net, _lib.ptr(self._flat), _lib.ptr(self._wsplit if mode == 1 else None),
                                            _lib.ptr(x), _lib.c_void_p(ei.data_ptr()), _lib.c_void_p(ei.data_ptr() + 8 * E),
                                            N, E, mode, _lib.ptr(out), _lib.ptr(emb), _lib.ptr(am), _lib.ptr(lw), need,
                                            _lib.stream_ptr())
        _lib.check(rc, "mdq_qnet_forward_layered")
        return out, emb, am

    def _use_layered(self, B, max_n, max_e=0):
        """One graph that does not fit the fused kernel's shared-memory tiles (mdq_qnet_smem_bytes, host-callable)
        goes layer by layer; batches of such graphs are not supported (loud MDQ_ESMEM from the fused launch)."""
        if B != 1:
            return False
        if max_n > self.FUSED_MAX_NODES:
            return True
        need = int(_lib.lib().mdq_qnet_smem_bytes(self._net, int(max_n), int(max_e), 1, 0))
        return need < 0 or need > 227 * 1024

    def _forward_impl(self, data, embedding=False):
        x, ei, nptr, eptr, B, max_n, max_e = self._prep(data)
        self._ensure_packed()
        self._wait_pending(x.device)
        if self._use_layered(B, max_n, max_e):
            if torch.is_grad_enabled() and any(p.requires_grad for _, p in self._entries):
                raise NotImplementedError("the layered large-graph path is forward-only (Q-evaluation); wrap the call in "
                                          "torch.no_grad()")
            out, emb, _ = self._launch_forward_layered(x, ei, embedding, False)
            return emb if embedding else out
        params = [p for _, p in self._entries]
        if not embedding:
            # through the dispatcher: torch.ops.meshdqn_b200.qnet_forward (ops.py) -- the same launch, with its autograd
            # formula (mdq_qnet_backward) and a fake implementation for tracing
            from . import ops
            return torch.ops.meshdqn_b200.qnet_forward(params, x, ei, nptr, eptr, ops.net_handle(self), B, max_n, max_e)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _QNetFunction.apply(self, x, ei, nptr, eptr, B, max_n, max_e, embedding, *params)
        out, emb, _ = self._launch_forward(x, ei, nptr, eptr, B, max_n, max_e, embedding, False)
        return emb if embedding else out

    @torch.no_grad()
    def select_action(self, data):
        """Fused softmax + argmax (airfoil_dqn.py:208-209): returns (action i32 [B], q [B, A])."""
        x, ei, nptr, eptr, B, max_n, max_e = self._prep(data)
        self._ensure_packed()
        self._wait_pending(x.device)
        if self._use_layered(B, max_n, max_e):
            out, _, am = self._launch_forward_layered(x, ei, False, True)
            return am, out
        from . import ops
        return torch.ops.meshdqn_b200.qnet_select_action(x, ei, nptr, eptr, ops.net_handle(self), B, max_n, max_e)


class NodeRemovalNet(_FusedQNet):
    """/root/reference/airfoilgcnn.py:24-145.  conv3/pool3/conv6/pool6 are constructed (their
    parameters exist in the state_dict) but unused by forward, exactly as in the reference."""

    _blocks_used = (1, 2, 4, 5)
    _softmax = 1

    def __init__(self, output_dim, conv_width=64, topk=0.5, initial_num_nodes=None):
        super().__init__()
        self.conv_width = conv_width
        self.topk = topk
        self.initial_num_nodes = initial_num_nodes
        w = conv_width
        self.conv1 = SAGEConv(2, w)
        self.pool1 = TopKPooling(w, ratio=topk)
        self.conv2 = SAGEConv(w, w)
        self.pool2 = TopKPooling(w, ratio=topk)
        self.conv3 = SAGEConv(w, w)
        self.pool3 = TopKPooling(w, ratio=topk)
        self.conv4 = GCNConv(w, w)
        self.pool4 = TopKPooling(w, ratio=topk)
        self.conv5 = GCNConv(w, w)
        self.pool5 = TopKPooling(w, ratio=topk)
        self.conv6 = GCNConv(w, w)
        self.pool6 = TopKPooling(w, ratio=topk)
        self.lin1 = torch.nn.Linear(2 * w, 128)
        self.lin2 = torch.nn.Linear(128, 64)
        self.lin3 = torch.nn.Linear(64, output_dim)
        self._flat = None
        torch.manual_seed(0)
        self.reset()

    def reset(self):
        with torch.no_grad():
            for c in (self.conv1, self.conv2, self.conv3):
                nn.init.xavier_normal_(c.lin_l.weight, gain=0.9)
                nn.init.normal_(c.lin_l.bias)
                nn.init.xavier_normal_(c.lin_r.weight, gain=0.9)
            for c in (self.conv4, self.conv5, self.conv6):
                nn.init.xavier_normal_(c.lin.weight, gain=0.9)
            for l in (self.lin1, self.lin2, self.lin3):
                nn.init.xavier_normal_(l.weight, gain=0.9)
                nn.init.normal_(l.bias)

    def set_num_nodes(self, initial_num_nodes):
        self.initial_num_nodes = initial_num_nodes
        dev = self.lin1.weight.device
        self.conv1 = SAGEConv(self.initial_num_nodes, self.conv_width).to(dev)
        self._flat = None

    def set_removable(self, removable):
        self.removable = removable

    def forward(self, data, embedding=False):
        return self._forward_impl(data, embedding)


class AirfoilGCNN(_FusedQNet):
    """/root/reference/airfoilgcnn.py:148-209: six blocks, TopK 0.5, features x[:, [2, 3]], scalar output."""

    _blocks_used = (1, 2, 3, 4, 5, 6)
    _softmax = 0
    _in_col0 = 2

    def __init__(self, conv_width=64):
        super().__init__()
        self.conv_width = conv_width
        self.topk = 0.5
        w, topk = conv_width, 0.5
        self.conv1 = SAGEConv(2, w)
        self.pool1 = TopKPooling(w, ratio=topk)
        self.conv2 = SAGEConv(w, w)
        self.pool2 = TopKPooling(w, ratio=topk)
        self.conv3 = SAGEConv(w, w)
        self.pool3 = TopKPooling(w, ratio=topk)
        self.conv4 = GCNConv(w, w)
        self.pool4 = TopKPooling(w, ratio=topk)
        self.conv5 = GCNConv(w, w)
        self.pool5 = TopKPooling(w, ratio=topk)
        self.conv6 = GCNConv(w, w)
        self.pool6 = TopKPooling(w, ratio=topk)
        self.lin1 = torch.nn.Linear(2 * w, 128)
        self.lin2 = torch.nn.Linear(128, 64)
        self.lin3 = torch.nn.Linear(64, 1)
        self._flat = None

    def forward(self, data):
        return self._forward_impl(data, False)
