import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-tf_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle of the C oracle (built on demand)."""
    import ctypes
    import __graft_entry__ as entry
    lib = ctypes.CDLL(entry.build_oracle())
    return lib


def pytest_sessionfinish(session, exitstatus):
    """Record the measured worst error of every parity comparison of this session
    (oracle.assert_close appends to PARITY_LOG): gpurun_out/parity_report.json, summarised in
    DESIGN.md §3."""
    import json
    try:
        import gnntf_oracle
    except Exception:
        return
    log = gnntf_oracle.PARITY_LOG
    if not log:
        return
    import torch
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    name = "parity_report_gpu.json" if torch.cuda.is_available() else "parity_report_cpu.json"
    worst = {}
    for rec in log:
        key = rec["what"] or "(unnamed)"
        cur = worst.get(key)
        if cur is None or rec["max_err_over_bound"] > cur["max_err_over_bound"]:
            worst[key] = rec
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump({"comparisons": len(log), "worst_by_name": worst}, f, indent=1, sort_keys=True)
