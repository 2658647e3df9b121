"""ctypes binding of the C-ABI in ``include/gnntf_b200.h`` (``libgnntf_b200.so``).

This is the only place Python touches the native library.  There is NO CPU or PyTorch
fallback: if the shared library is missing, every entry point raises.  PyTorch is used for
device memory and streams only (``tensor.data_ptr()``, ``torch.cuda.current_stream()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# GNNTF_B200_LIB points the binding at another build of the same ABI (A/B measurements of whole
# library versions); unset = the in-tree build.
LIB_PATH = os.environ.get("GNNTF_B200_LIB") or os.path.join(_HERE, "_lib", "libgnntf_b200.so")

GNNTF_OK = 0
GNNTF_E_SHAPE = -6   # a specialised entry point declined the shape (nothing was enqueued)
NORM = {"symmetric": 0, "bipartite": 1, "none": 2}
EYE = {"none": 0, "before": 1, "after": 2}
ACT_IDENTITY, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2

# every symbol include/gnntf_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "gnntf_abi_version", "gnntf_status_str", "gnntf_csr_build_ws_bytes", "gnntf_csr_build",
    "gnntf_normalize_f32", "gnntf_arrange_sweep_f32", "gnntf_spmm_plan_count", "gnntf_spmm_plan_fill", "gnntf_spmm_f32", "gnntf_spmm_acc_f32",
    "gnntf_appnp_step_f32", "gnntf_appnp_propagate_f32", "gnntf_appnp_propagate_multi_f32",
    "gnntf_appnp_propagate_bwd_f32", "gnntf_appnp_propagate_host_f32", "gnntf_appnp_propagate_host_batched_f32", "gnntf_appnp_propagate_cluster_f32",
    "gnntf_halo_pack_f32", "gnntf_halo_push_f32", "gnntf_ipc_alloc", "gnntf_ipc_open", "gnntf_ipc_close", "gnntf_ipc_free",
    "gnntf_halo_push_signal_f32", "gnntf_step_push_f32", "gnntf_flags_wait", "gnntf_flags_signal", "gnntf_peer_copy_signal",
    "gnntf_bias_act_dropout_f32", "gnntf_bias_act_dropout_bwd_f32", "gnntf_node_xent_f32", "gnntf_node_xent_bwd_f32",
]


class CsrStruct(Structure):
    """``gnntf_csr_t``."""
    _fields_ = [
        ("n_rows", c_int64), ("nnz", c_int64),
        ("row_ptr", c_void_p), ("col_idx", c_void_p), ("val", c_void_p), ("row_map", c_void_p),
        ("long_threshold", c_int32), ("chunk", c_int32), ("n_long", c_int32), ("n_chunks", c_int32),
        ("long_row", c_void_p), ("long_first_chunk", c_void_p), ("long_n_chunks", c_void_p),
        ("chunk_row", c_void_p), ("chunk_begin", c_void_p), ("partials", c_void_p),
    ]


_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the native library; raise loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"gnntf_b200 native library not found at {LIB_PATH}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C gnn-tf_b200/csrc`). "
            f"There is no CPU fallback for the propagation path.")
    L = ctypes.CDLL(LIB_PATH)
    L.gnntf_abi_version.restype = c_int
    L.gnntf_status_str.restype = c_char_p
    L.gnntf_status_str.argtypes = [c_int]
    L.gnntf_csr_build_ws_bytes.argtypes = [c_int64, c_int64, c_int, c_int, POINTER(c_size_t)]
    L.gnntf_csr_build.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]
    L.gnntf_normalize_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                      c_int, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]
    L.gnntf_arrange_sweep_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]
    L.gnntf_spmm_plan_count.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]
    L.gnntf_spmm_plan_fill.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]
    L.gnntf_spmm_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p]
    L.gnntf_spmm_acc_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_int64, c_void_p, c_int64, c_int64, c_double, c_void_p]
    L.gnntf_appnp_step_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                       c_double, c_void_p, c_float, c_int, c_void_p]
    L.gnntf_appnp_propagate_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_void_p, c_int64,
                                            c_int64, c_double, c_int, c_void_p]
    L.gnntf_appnp_propagate_multi_f32.argtypes = [POINTER(CsrStruct), c_int, c_void_p, c_void_p, c_void_p,
                                                  c_int64, c_int64, c_double, c_void_p]
    L.gnntf_appnp_propagate_bwd_f32.argtypes = [POINTER(CsrStruct), c_int, c_void_p, c_void_p, c_void_p,
                                                c_int64, c_int64, c_double, c_void_p]
    L.gnntf_appnp_propagate_host_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_int64, c_int64, c_double, c_int, c_void_p]
    L.gnntf_appnp_propagate_cluster_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_int64, c_int64, c_double,
                                                    c_int, c_int, c_int, c_void_p]
    L.gnntf_appnp_propagate_host_batched_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_int, c_void_p,
                                                         c_int64, c_int64, c_double, c_int, c_void_p]
    L.gnntf_halo_pack_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p]
    L.gnntf_halo_push_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                      c_int64, c_int64, c_void_p]
    L.gnntf_peer_copy_signal.argtypes = [c_void_p, c_void_p, ctypes.c_size_t, c_void_p, c_void_p, c_void_p]
    L.gnntf_halo_push_signal_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                             c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int32, c_void_p]
    L.gnntf_step_push_f32.argtypes = [POINTER(CsrStruct), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                      c_void_p, c_int32, c_void_p]
    L.gnntf_flags_wait.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int32, c_void_p]
    L.gnntf_flags_signal.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int32, c_void_p]
    L.gnntf_bias_act_dropout_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_int, c_float, c_void_p,
                                             c_int64, c_int64, c_int64, c_void_p]
    L.gnntf_bias_act_dropout_bwd_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float, c_int, c_float,
                                                 c_void_p, c_int64, c_int64, c_int64, c_void_p]
    L.gnntf_node_xent_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    L.gnntf_node_xent_bwd_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                          c_int64, c_void_p]
    L.gnntf_ipc_alloc.argtypes = [c_size_t, POINTER(c_void_p), c_void_p]
    L.gnntf_ipc_open.argtypes = [c_void_p, POINTER(c_void_p)]
    L.gnntf_ipc_close.argtypes = [c_void_p]
    L.gnntf_ipc_free.argtypes = [c_void_p]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("gnntf_status_str",):
            fn.restype = c_int
    if L.gnntf_abi_version() != 1:
        raise RuntimeError(f"gnntf_b200 ABI version mismatch: library reports {L.gnntf_abi_version()}, binding expects 1")
    _lib = L
    return L


def check(rc: int, what: str = "gnntf_b200"):
    """Translate a status code into the reference's error convention (bare ``Exception``)."""
    if rc == GNNTF_OK:
        return
    msg = lib().gnntf_status_str(rc).decode()
    raise Exception(f"{what}: {msg} (status {rc})" if rc != -3 else msg)


def ptr(t):
    """Device/host pointer of a tensor as ``c_void_p`` (``None`` -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


__all__ = ["lib", "check", "ptr", "stream_ptr", "CsrStruct", "NORM", "EYE", "SYMBOLS", "LIB_PATH",
           "NativeLibraryMissing", "ACT_IDENTITY", "ACT_RELU", "byref"]
