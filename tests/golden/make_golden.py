"""Writes tests/golden/kat.json.

The reference cannot run here (TensorFlow is not installable: no network, un-pinned dependency,
setup.py:26-28) and ships no golden vectors, so these fixtures are NOT reference outputs: the
first two cases are the known-answer tests SURVEY.md §8c derives by hand from the cited reference
lines (their numbers are typed in below, not computed), the rest are produced by the fp64 oracle
(oracle/gnntf_oracle.py) on small seeded graphs and serve as regression anchors for the fp32
oracle and the CUDA path.  Parity therefore stays "unpinned" in the sense of DESIGN.md.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gnntf_oracle as oracle  # noqa: E402

KAT_IDX = [[0, 3], [1, 2], [1, 0], [2, 1], [3, 3], [3, 0], [2, 1], [0, 1], [1, 2], [3, 3]]
KAT_VAL = [2.5, 1, 1, 1, 1, 2.5, 1, 1, 1, 1]
KAT_H0 = [[(3 * i + j) / 7 - 1 for j in range(3)] for i in range(5)]

cases = [
    dict(name="KAT-1 (SURVEY §8c, hand-derived)", n=5, indices=KAT_IDX, values=KAT_VAL, keep=None, rate=0.0,
         normalized="symmetric", add_eye="none", alpha=0.1, K=10, H0=KAT_H0,
         norm_values=[0.6299408, 0.40824828, 0.30860668, 0.40824828, 0.22222225] * 2,
         H_K=[[-0.386717034, -0.242003374, -0.097289714], [-0.37323383, -0.233458343, -0.093682856],
              [-0.301202958, -0.181881658, -0.062560359], [-0.311411903, -0.151985331, 0.007441241],
              [0.071428571, 0.085714286, 0.1]]),
]
# KAT-2: normalised values typed in from the survey; H_K from the fp64 oracle
keep2 = [1, 0, 1, 1, 0, 1, 1, 0, 1, 1]
idx = np.array(KAT_IDX, np.int64)
mv = oracle.sparse_dropout(np.array(KAT_VAL, np.float32), 0.5, keep2)
h = oracle.appnp_propagate(idx, mv, 5, np.array(KAT_H0), 0.1, 10, dtype=np.float64)[-1]
cases.append(dict(name="KAT-2 (SURVEY §8c, hand-derived mask case)", n=5, indices=KAT_IDX, values=KAT_VAL, keep=keep2,
                  rate=0.5, normalized="symmetric", add_eye="none", alpha=0.1, K=10, H0=KAT_H0,
                  norm_values=[0.7142858, 0, 0.3779645, 0.70710677, 0, 0.7142858, 0.70710677, 0, 0.70710677, 0.28571433],
                  H_K=h.tolist()))
rng = np.random.default_rng(7)
for name, n, e, F, mode, eye in [("rand-sym", 23, 60, 4, "symmetric", "none"), ("rand-bip-eye", 17, 40, 3, "bipartite", "after"),
                                 ("rand-sym-eye-before", 19, 50, 2, "symmetric", "before")]:
    edges = rng.integers(0, n, (e, 2))
    w = np.round(rng.random(e) + 0.5, 3).astype(np.float32)
    ci, cv, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(ci, cv, n, mode, eye, dtype=np.float64)
    H0 = np.round(rng.standard_normal((n, F)), 4)
    hk = oracle.appnp_propagate(ci, cv, n, H0, 0.1, 10, dtype=np.float64)[-1]
    cases.append(dict(name=name, n=n, indices=ci.tolist(), values=cv.tolist(), keep=None, rate=0.0, normalized=mode,
                      add_eye=eye, alpha=0.1, K=10, H0=H0.tolist(), norm_values=nv.tolist(), H_K=hk.tolist()))

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(dict(source="see make_golden.py docstring", cases=cases), f)
print("wrote", len(cases), "cases")
