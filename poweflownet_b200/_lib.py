"""ctypes binding of libpfn_b200.so (the C ABI declared in include/pfn_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing callers get a
RuntimeError that says how to build it (`python -m poweflownet_b200.build`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libpfn_b200.so")

c_f32p = C.c_void_p  # device pointers travel as plain addresses
c_i64 = C.c_int64
c_u64 = C.c_uint64
c_sz = C.c_size_t


class GraphLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "meta", "rowptr_t", "nbr_t", "eid_t", "ea_t", "rowptr_s", "nbr_s", "eid_s", "ea_s", "deg", "dis", "cursor",
        "total_bytes", "e_cap")]


class MpnDesc(C.Structure):
    _fields_ = [("nfeature_dim", C.c_int32), ("efeature_dim", C.c_int32), ("output_dim", C.c_int32),
                ("hidden_dim", C.c_int32), ("n_gnn_layers", C.c_int32), ("K", C.c_int32),
                ("dropout_rate", C.c_float), ("reserved", C.c_int32)]


class DatasetCase(C.Structure):
    _fields_ = [("node_features", C.c_void_p), ("edge_features", C.c_void_p), ("n_samples", C.c_int64),
                ("n_nodes", C.c_int64), ("n_edges", C.c_int64)]


# name -> (restype, argtypes); mirrors include/pfn_b200.h one to one
SIGNATURES = {
    "pfn_version": (C.c_char_p, []),
    "pfn_last_error": (C.c_char_p, []),
    "pfn_launch_count": (c_u64, []),
    "pfn_profile_enable": (C.c_int, [C.c_int]),
    "pfn_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(c_i64)]),
    "pfn_graph_layout_get": (C.c_int, [c_i64, c_i64, C.POINTER(GraphLayout)]),
    "pfn_graph_prep": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p, C.c_void_p]),
    "pfn_graph_prep_tiled": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, c_i64, C.c_void_p, C.c_void_p]),
    "pfn_graph_prep_tiled_supported": (C.c_int, [c_i64, c_i64, c_i64]),
    "pfn_graph_meta": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "pfn_graph_export": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfn_ea_fwd": (C.c_int, [c_f32p, c_f32p, c_i64, C.c_void_p, c_i64, c_i64, c_f32p, c_i64, c_f32p, c_i64, c_i64,
                             C.c_void_p]),
    "pfn_ea_bwd_scratch_bytes": (c_sz, [c_i64]),
    "pfn_ea_bwd": (C.c_int, [c_f32p, c_i64, c_f32p, c_f32p, c_i64, C.c_void_p, c_i64, c_i64, c_f32p, c_i64, c_f32p,
                             c_f32p, c_i64, c_f32p, c_i64, C.c_void_p, c_i64, C.c_void_p]),
    "pfn_spmm_hop": (C.c_int, [c_f32p, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, c_f32p, c_i64, c_f32p, c_i64,
                               C.c_float, c_f32p, c_i64, c_i64, C.c_void_p]),
    "pfn_linear_fwd": (C.c_int, [c_f32p, c_i64, c_f32p, c_i64, c_f32p, c_f32p, c_f32p, c_i64, c_f32p, c_i64, c_i64,
                                 c_i64, c_i64, C.c_int, C.c_float, c_u64, c_f32p, c_i64, C.c_void_p]),
    "pfn_linear_dgrad": (C.c_int, [c_f32p, c_i64, c_f32p, c_i64, c_f32p, c_i64, C.c_float, c_f32p, c_i64, c_i64,
                                   c_i64, c_i64, C.c_void_p]),
    "pfn_linear_wgrad_scratch_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pfn_linear_wgrad": (C.c_int, [c_f32p, c_i64, c_f32p, c_i64, c_f32p, c_f32p, c_i64, c_f32p, c_i64, c_i64, c_i64,
                                   C.c_void_p, C.c_void_p]),
    "pfn_mpn_num_params": (C.c_int, [C.POINTER(MpnDesc)]),
    "pfn_mpn_workspace": (C.c_int, [C.POINTER(MpnDesc), c_i64, c_i64, C.POINTER(c_sz), C.POINTER(c_sz)]),
    "pfn_mpn_forward": (C.c_int, [C.POINTER(MpnDesc), C.c_void_p, c_f32p, C.c_void_p, c_i64, c_i64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, c_u64, C.c_void_p, C.c_void_p, c_f32p, C.c_void_p]),
    "pfn_mpn_fused_supported": (C.c_int, [C.POINTER(MpnDesc), c_i64]),
    "pfn_mpn_forward_tiled": (C.c_int, [C.POINTER(MpnDesc), C.c_void_p, c_f32p, C.c_void_p, c_i64, c_i64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, c_u64, C.c_void_p, C.c_void_p, c_f32p, c_i64,
                                        C.c_void_p, c_i64, C.c_void_p]),
    "pfn_graph_tile_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "pfn_mpn_backward": (C.c_int, [C.POINTER(MpnDesc), C.c_void_p, C.c_void_p, c_f32p, c_i64, c_i64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pfn_mpn_backward_tiled": (C.c_int, [C.POINTER(MpnDesc), C.c_void_p, C.c_void_p, c_f32p, c_i64, c_i64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int, c_i64, c_i64, C.c_void_p]),
    "pfn_mse_scratch_bytes": (c_sz, [c_i64]),
    "pfn_mse_fwd_bwd": (C.c_int, [c_f32p, c_f32p, c_i64, C.c_float, c_f32p, c_f32p, C.c_void_p, C.c_void_p]),
    "pfn_masked_l2_scratch_bytes": (c_sz, [c_i64]),
    "pfn_masked_l2_fwd_bwd": (C.c_int, [c_f32p, c_f32p, C.c_void_p, c_i64, C.c_int, C.c_float, c_f32p, c_f32p, c_f32p,
                                        C.c_void_p, C.c_void_p]),
    "pfn_power_imbalance_scratch_bytes": (c_sz, [c_i64]),
    "pfn_power_imbalance_fwd_bwd": (C.c_int, [c_f32p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, c_f32p, c_f32p, c_i64,
                                              C.c_void_p, C.c_void_p]),
    "pfn_batch_assemble_scratch_bytes": (c_sz, [c_i64]),
    "pfn_batch_assemble": (C.c_int, [C.POINTER(DatasetCase), C.c_int, C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, c_u64,
                                     c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, c_f32p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pfn_batch_assemble_status": (C.c_int, [C.c_void_p, c_i64, C.POINTER(C.c_int32), C.c_void_p]),
    "pfn_allreduce_peer": (C.c_int, [c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_i64, C.c_int, C.c_int,
                                     C.c_void_p]),
    "pfn_allreduce_debug_stamps": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "pfn_adamw_step": (C.c_int, [c_i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_double, c_i64, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -m poweflownet_b200.build` (needs nvcc). There is no CPU fallback for this path.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover - depends on the box
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}. There is no CPU fallback for this path.") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pfn_last_error().decode(errors="replace")
        raise RuntimeError(f"libpfn_b200 {what} failed (code {rc}): {msg}")
