"""The C-ABI library loads on a CPU-only box and exports every symbol include/gnntf_b200.h
declares.  No compute entry point is called here (there is no GPU): only version/status strings
and argument validation that returns before any CUDA call."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gnntf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnntf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from gnntf import _native
    lib = _native.lib()
    names = _declared()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_native.SYMBOLS) == names, "ctypes binding list and header disagree"


def test_version_and_status_strings():
    from gnntf import _native
    lib = _native.lib()
    assert lib.gnntf_abi_version() == 1
    assert lib.gnntf_status_str(0) == b"ok"
    assert lib.gnntf_status_str(-3) == b"Invalid matrix normalization"      # gnn.py:46-47 wording
    assert b"NULL" in lib.gnntf_status_str(-1)
    assert b"workspace" in lib.gnntf_status_str(-4)


def test_argument_errors_return_codes_not_exceptions():
    from gnntf import _native
    lib = _native.lib()
    null = ctypes.c_void_p(0)
    assert lib.gnntf_spmm_f32(None, null, 0, null, 0, 4, null) == -1
    assert lib.gnntf_appnp_propagate_f32(None, null, null, null, 4, 4, 0.1, 10, null) == -1
    csr = _native.CsrStruct()
    csr.n_rows, csr.nnz = 5, -1
    assert lib.gnntf_spmm_f32(ctypes.byref(csr), null, 4, ctypes.c_void_p(16), 4, 4, null) == -2
    csr.nnz = 2 ** 31
    assert lib.gnntf_spmm_f32(ctypes.byref(csr), null, 4, ctypes.c_void_p(16), 4, 4, null) == -2   # nnz >= 2^31
    csr.nnz = 3
    assert lib.gnntf_spmm_f32(ctypes.byref(csr), null, 4, ctypes.c_void_p(16), 4, 4, null) == -1   # row_ptr NULL
    assert lib.gnntf_normalize_f32(null, null, null, null, 5, 3, 3, 0, null, 1.0, 7, 0, null, null, null, null,
                                   null, null) == -3
    assert lib.gnntf_csr_build_ws_bytes(-1, 0, 0, 0, ctypes.byref(ctypes.c_size_t())) == -2
    assert lib.gnntf_csr_build_ws_bytes(10, 2 ** 31, 0, 0, ctypes.byref(ctypes.c_size_t())) == -2
    assert lib.gnntf_halo_pack_f32(null, 2, null, 3, null, 4, 4, null) == -2                       # ld < F
    # entry points added in round 2: argument errors come back before any CUDA call
    assert lib.gnntf_appnp_propagate_host_batched_f32(None, null, null, 2, null, 4, 4, 0.1, 10, null) == -1
    csr.n_rows, csr.nnz = 5, 3
    assert lib.gnntf_appnp_propagate_host_batched_f32(ctypes.byref(csr), null, null, -1, null, 4, 4, 0.1, 10, null) == -2
    assert lib.gnntf_appnp_propagate_host_batched_f32(ctypes.byref(csr), null, null, 0, null, 4, 4, 0.1, 10, null) == 0   # nothing to do
    assert lib.gnntf_appnp_propagate_host_batched_f32(ctypes.byref(csr), null, null, 2, null, 4, 4, 0.1, 10, null) == -1
    assert lib.gnntf_appnp_propagate_cluster_f32(None, null, null, 4, 4, 0.1, 10, 0, 0, null) == -1
    assert lib.gnntf_peer_copy_signal(null, null, 16, null, null, null) == -1
    assert lib.gnntf_peer_copy_signal(null, null, 0, null, null, null) == 0                        # no rows, no flag: nothing enqueued
    assert b"shape" in lib.gnntf_status_str(-6)
    with pytest.raises(Exception, match="Invalid matrix normalization"):
        _native.check(-3)


def test_product_path_fails_loudly_without_gpu_or_library(monkeypatch):
    import torch
    from gnntf import _native
    import gnntf
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            gnntf.edges2adj([[0, 1]], None, 2)
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libgnntf_b200.so")
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnn-tf_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(base, f)).read()
                assert "gnntf_oracle" not in src and "oracle_c" not in src, f"{f} references the oracle"
