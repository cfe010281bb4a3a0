"""``torch.library`` registration of the hot-path entry points: namespace ``meshdqn_b200``.

The north-star asks for "a thin C-ABI torch custom-op layer".  The C ABI is ``include/meshdqn_b200.h``; this module puts
the entry points a PyTorch program composes under ``torch.ops.meshdqn_b200.*`` -- visible to the dispatcher, with fake
(meta) implementations so that ``torch.compile`` / FakeTensor tracing treat them as opaque ops of known output shape, and
with an autograd formula for the Q-network -- on top of the same launchers the module classes use.  CUDA only: there is
no CPU kernel behind any of them (``device_types="cuda"``; a CPU tensor raises NotImplementedError from the dispatcher).

    qnet_forward(params[], x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e) -> Q [n_graphs, A]
        differentiable w.r.t. ``params`` (mdq_qnet_backward / the staged backward); airfoilgcnn.py:85-145
    qnet_select_action(x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e) -> (action i32 [n], Q [n, A])
        fused softmax + argmax, airfoil_dqn.py:208-209
    mesh_smooth(coords, nbr_ptr, nbr_idx, vc_ptr, vc_idx, cells, on_boundary, iters) -> coords'     flow_solver.py:67,237
    polygon_distance(coords, idx, ring) -> dist f64 [len(idx)]                                      Env2DAirfoil.py:232-241
    drag_lift(coords, cells, cell_edges, tags, edge_cell, U, P, mu, n_edges) -> [2, T] f64           probes.py:23-50

``net`` is an integer handle of a network descriptor + flat parameter buffer (``net_handle(module)``): the descriptor is a
C struct, not a tensor.
"""
from __future__ import annotations

import itertools
import weakref
from typing import Sequence, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd

from . import _lib

_NETS = weakref.WeakValueDictionary()
_counter = itertools.count(1)


def net_handle(module) -> int:
    """Integer handle under which the ops find ``module`` (a NodeRemovalNet / AirfoilGCNN); stable for its lifetime."""
    h = module.__dict__.get("_op_handle")
    if h is None:
        h = module.__dict__["_op_handle"] = next(_counter)
        _NETS[h] = module
    return h


def _net(h):
    m = _NETS.get(int(h))
    if m is None:
        raise RuntimeError(f"meshdqn_b200 ops: unknown network handle {h} (the module was freed?)")
    return m


# ------------------------------------------------------------------------------------------------ Q-network
@custom_op("meshdqn_b200::qnet_forward", mutates_args=(), device_types="cuda")
def qnet_forward(params: Sequence[Tensor], x: Tensor, edge_index: Tensor, node_ptr: Tensor, edge_ptr: Tensor, net: int,
                 n_graphs: int, max_n: int, max_e: int) -> Tensor:
    m = _net(net)
    out, _, _ = m._launch_forward(x, edge_index, node_ptr, edge_ptr, n_graphs, max_n, max_e, False, False)
    return out


@qnet_forward.register_fake
def _(params, x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e):
    return x.new_empty((n_graphs, _net(net).lin3.out_features), dtype=torch.float32)


@custom_op("meshdqn_b200::qnet_backward", mutates_args=(), device_types="cuda")
def qnet_backward(grad_out: Tensor, x: Tensor, edge_index: Tensor, node_ptr: Tensor, edge_ptr: Tensor, net: int,
                  n_graphs: int, max_n: int, max_e: int) -> Tensor:
    m = _net(net)
    flat_grad = torch.empty_like(m._flat)
    m._launch_backward(x, edge_index, node_ptr, edge_ptr, n_graphs, max_n, max_e, grad_out.contiguous().float(), flat_grad)
    return flat_grad


@qnet_backward.register_fake
def _(grad_out, x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e):
    return x.new_empty((_net(net)._flat.numel(),), dtype=torch.float32)


def _qnet_setup(ctx, inputs, output):
    params, x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e = inputs
    ctx.save_for_backward(x, edge_index, node_ptr, edge_ptr)
    ctx.meta = (net, n_graphs, max_n, max_e)
    ctx.needs = [bool(p.requires_grad) for p in params]


def _qnet_bwd(ctx, grad_out):
    x, edge_index, node_ptr, edge_ptr = ctx.saved_tensors
    net, n_graphs, max_n, max_e = ctx.meta
    m = _net(net)
    flat_grad = torch.ops.meshdqn_b200.qnet_backward(grad_out, x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e)
    grads = [m._grad_view(flat_grad, name) if (need and name not in m._unused) else None
             for need, (name, _) in zip(ctx.needs, m._named_flat_params())]
    return grads, None, None, None, None, None, None, None, None


register_autograd("meshdqn_b200::qnet_forward", _qnet_bwd, setup_context=_qnet_setup)


@custom_op("meshdqn_b200::qnet_select_action", mutates_args=(), device_types="cuda")
def qnet_select_action(x: Tensor, edge_index: Tensor, node_ptr: Tensor, edge_ptr: Tensor, net: int, n_graphs: int,
                       max_n: int, max_e: int) -> Tuple[Tensor, Tensor]:
    m = _net(net)
    out, _, am = m._launch_forward(x, edge_index, node_ptr, edge_ptr, n_graphs, max_n, max_e, False, True)
    return am, out


@qnet_select_action.register_fake
def _(x, edge_index, node_ptr, edge_ptr, net, n_graphs, max_n, max_e):
    return (x.new_empty((n_graphs,), dtype=torch.int32), x.new_empty((n_graphs, _net(net).lin3.out_features), dtype=torch.float32))


# ------------------------------------------------------------------------------------------------ mesh services
@custom_op("meshdqn_b200::mesh_smooth", mutates_args=(), device_types="cuda")
def mesh_smooth(coords: Tensor, nbr_ptr: Tensor, nbr_idx: Tensor, vc_ptr: Tensor, vc_idx: Tensor, cells: Tensor,
                on_boundary: Tensor, iters: int) -> Tensor:
    out = coords.clone()
    status = torch.zeros(1, dtype=torch.int32, device=coords.device)
    p = _lib.ptr
    with torch.cuda.device(coords.device):
        rc = _lib.lib().mdq_mesh_smooth(p(out), int(coords.shape[0]), int(cells.shape[0]), p(nbr_ptr), p(nbr_idx), p(vc_ptr),
                                        p(vc_idx), p(cells), p(on_boundary), int(iters), p(status), _lib.stream_ptr())
    _lib.check(rc, "mdq_mesh_smooth")
    return out


@mesh_smooth.register_fake
def _(coords, nbr_ptr, nbr_idx, vc_ptr, vc_idx, cells, on_boundary, iters):
    return torch.empty_like(coords)


@custom_op("meshdqn_b200::polygon_distance", mutates_args=(), device_types="cuda")
def polygon_distance(coords: Tensor, idx: Tensor, ring: Tensor) -> Tensor:
    n = int(idx.shape[0])
    out = torch.empty(n, dtype=torch.float64, device=coords.device)
    p = _lib.ptr
    with torch.cuda.device(coords.device):
        rc = _lib.lib().mdq_polygon_distance(p(coords), p(idx), n, p(ring), int(ring.shape[0]), p(out), _lib.stream_ptr())
    _lib.check(rc, "mdq_polygon_distance")
    return out


@polygon_distance.register_fake
def _(coords, idx, ring):
    return coords.new_empty((idx.shape[0],), dtype=torch.float64)


@custom_op("meshdqn_b200::drag_lift", mutates_args=(), device_types="cuda")
def drag_lift(coords: Tensor, cells: Tensor, cell_edges: Tensor, tags: Tensor, edge_cell: Tensor, U: Tensor, P: Tensor,
              mu: float, n_edges: int) -> Tensor:
    T = int(U.shape[0])
    out = torch.empty((2, T), dtype=torch.float64, device=coords.device)
    p = _lib.ptr
    with torch.cuda.device(coords.device):
        rc = _lib.lib().mdq_drag_lift(p(coords), p(cells), p(cell_edges), int(coords.shape[0]), int(n_edges), p(tags),
                                      p(edge_cell), T, p(U), p(P), float(mu), p(out), _lib.stream_ptr())
    _lib.check(rc, "mdq_drag_lift")
    return out


@drag_lift.register_fake
def _(coords, cells, cell_edges, tags, edge_cell, U, P, mu, n_edges):
    return coords.new_empty((2, U.shape[0]), dtype=torch.float64)
