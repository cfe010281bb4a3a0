"""ctypes loader / builder for libmeshdqn_b200.so (the C-ABI in include/meshdqn_b200.h).

There is NO CPU fallback: if the shared library is missing or a CUDA device is
absent, every op raises.  `build()` cross-compiles the kernels for sm_100a with
nvcc (works without a GPU); the built .so lives in-tree so it travels to the GPU
box with the repo snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libmeshdqn_b200.so")
_OBJ = os.path.join(_HERE, "csrc", "_obj")

# (source, extra flags).  geom.cu holds float64 geometry whose results must equal the CPU
# oracle's operation for operation, so FMA contraction is off there.
_SOURCES = (
    ("common.cu", ()),
    ("gnn_fused.cu", ()),
    ("gnn_layered.cu", ()),
    ("geom.cu", ("-fmad=false",)),
    ("replay_buffer.cu", ()),
)
_NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-I" + os.path.join(_ROOT, "include"), "-I" + _CSRC)

MDQ_MAX_BLOCKS = 6
MDQ_BLOCK_SAGE = 0
MDQ_BLOCK_GCN = 1


class mdq_block_t(Structure):
    _fields_ = [("type", c_int32), ("kin", c_int32), ("w_off", c_int32), ("b_off", c_int32), ("pool_off", c_int32)]


class mdq_net_t(Structure):
    _fields_ = [("n_blocks", c_int32), ("width", c_int32), ("in_dim", c_int32), ("in_col0", c_int32),
                ("x_stride", c_int32), ("ratio", c_float), ("softmax", c_int32), ("out_dim", c_int32),
                ("lin_off", c_int32 * 3), ("lin_boff", c_int32 * 3), ("lin_in", c_int32 * 3), ("lin_out", c_int32 * 3),
                ("n_params", c_int32), ("blk", mdq_block_t * MDQ_MAX_BLOCKS)]


class mdq_tile_index_t(Structure):
    _fields_ = [("n_leaves", c_int32), ("depth", c_int32), ("T", c_int32), ("max_nv", c_int32), ("max_np2", c_int32),
                ("max_nc", c_int32), ("max_nbin", c_int32), ("max_nent", c_int32), ("u_stride", c_int64),
                ("p_stride", c_int64), ("tree", c_void_p), ("leaf_info", c_void_p), ("leaf_rect", c_void_p),
                ("coordsL", c_void_p), ("UL", c_void_p), ("PL", c_void_p), ("gidL", c_void_p), ("cvL", c_void_p),
                ("binptrL", c_void_p), ("binsL", c_void_p), ("leaf_base", c_void_p), ("total_cap", c_int64),
                ("smem_bytes", c_int32), ("reserved", c_int32)]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libmeshdqn_b200.so (in-tree)."""
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(_OBJ, exist_ok=True)
    hdrs = [os.path.join(_ROOT, "include", "meshdqn_b200.h")] + \
        [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    objs = []
    relink = force or not os.path.exists(_SO)
    for name, extra in _SOURCES:
        src = os.path.join(_CSRC, name)
        obj = os.path.join(_OBJ, name.replace(".cu", ".o"))
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            cmd = [nvcc, *_NVCC_FLAGS, *extra, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            relink = True
        objs.append(obj)
    if relink:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", _SO, *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return _SO


_lib = None

_P = c_void_p
_SIGS = {
    "mdq_last_error": (ctypes.c_char_p, []),
    "mdq_version": (c_int, []),
    "mdq_launch_count": (c_int64, []),
    "mdq_launch_count_add": (None, [c_int64]),
    "mdq_qnet_set_trace": (None, [_P]),
    "mdq_qnet_smem_bytes": (c_int64, [POINTER(mdq_net_t), c_int, c_int, c_int, c_int]),
    "mdq_qnet_occupancy": (c_int, [POINTER(mdq_net_t), c_int, c_int, c_int]),
    "mdq_qnet_pick_gpc": (c_int, [POINTER(mdq_net_t), c_int, c_int, c_int]),
    "mdq_qnet_forward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "mdq_qnet_bwd_workspace_floats": (c_int64, [POINTER(mdq_net_t), c_int, c_int]),
    "mdq_qnet_backward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "mdq_qnet_replay_backward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                                         _P, c_int, c_float, _P, _P, _P, _P, _P]),
    "mdq_qnet_staged_supported": (c_int, [POINTER(mdq_net_t), c_int, c_int]),
    "mdq_qnet_staged_wsplit_floats": (c_int64, [POINTER(mdq_net_t)]),
    "mdq_qnet_staged_wsplit": (c_int, [POINTER(mdq_net_t), _P, _P, _P]),
    "mdq_qnet_staged_workspace_floats": (c_int64, [POINTER(mdq_net_t), c_int, c_int, c_int, c_int]),
    "mdq_qnet_staged_forward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mdq_qnet_staged_backward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "mdq_qnet_staged_replay_backward": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int,
                                                _P, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P, c_int, _P, _P]),
    "mdq_stream_post": (c_int, [_P, _P]),
    "mdq_qnet_layered_workspace_bytes": (c_int64, [POINTER(mdq_net_t), c_int, c_int]),
    "mdq_qnet_forward_layered": (c_int, [POINTER(mdq_net_t), _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_int64, _P]),
    "mdq_csr_build": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P, _P]),
    "mdq_csr_build_scratch_words": (c_int64, [c_int, c_int]),
    "mdq_sage_aggregate": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, _P]),
    "mdq_node_gemm": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "mdq_huber_replay": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P]),
    "mdq_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, c_int, _P]),
    "mdq_adam_step_dev": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, _P, _P]),
    "mdq_allreduce_stage_floats": (c_int64, [c_int64, c_int]),
    "mdq_debug_smooth_trace": (c_int, [_P]),
    "mdq_debug_fast_math_check": (c_int, [c_uint64, c_int64, c_int, _P, _P]),
    "mdq_allreduce_adam": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, _P, _P, c_int, c_int, _P, _P]),
    "mdq_scan_i32": (c_int, [_P, _P, c_int, _P]),
    "mdq_mesh_topology": (c_int, [_P, c_int, c_int] + [_P] * 14),
    "mdq_mesh_smooth": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P]),
    "mdq_mesh_tags_removable": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, _P, _P, _P]),
    "mdq_polygon_distance": (c_int, [_P, _P, c_int, _P, c_int, _P, _P]),
    "mdq_grid_count": (c_int, [_P, _P, c_int, POINTER(c_double), _P, _P]),
    "mdq_grid_fill": (c_int, [_P, _P, c_int, POINTER(c_double), _P, _P, _P, _P]),
    "mdq_interpolate": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, POINTER(c_double), _P, _P,
                                c_double, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mdq_interpolate_drag_lift": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, POINTER(c_double), _P, _P,
                                c_double, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_double, _P, _P, _P]),
    "mdq_interp_miss_distance": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "mdq_interp_tiled_counter_words": (c_int64, [POINTER(mdq_tile_index_t)]),
    "mdq_interp_tiled_scratch_words": (c_int64, [POINTER(mdq_tile_index_t), c_int]),
    "mdq_interp_tiled_smem_bytes": (c_int64, [POINTER(mdq_tile_index_t)]),
    "mdq_interpolate_tiled": (c_int, [_P, c_int, _P, c_int, POINTER(mdq_tile_index_t), _P, _P, _P, c_int, c_int, c_int,
                                      _P, _P, c_double, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mdq_drag_lift": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, c_double, _P, _P]),
    "mdq_replay_store": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "mdq_replay_gather": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, c_int64, _P, _P, _P,
                                  _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mdq_build_state": (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P,
                                _P, _P, c_int, _P, _P]),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Load the shared library (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(meshdqn_b200 has no CPU fallback)")
        L = ctypes.CDLL(_SO)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mdq_last_error().decode()
        raise RuntimeError(f"meshdqn_b200 {what} failed (rc={rc}): {msg}")


def stream_ptr():
    import torch
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())
