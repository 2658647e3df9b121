import sys, json, numpy as np, torch
sys.path.insert(0, "gnn-tf_b200")
import gnntf, synthetic
from gnntf import ops
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
n, edges = synthetic.rmat_edges(23, 100_000_000, seed=0, device="cuda")
adj = gnntf.edges2adj(edges, None, n); A = adj.normalized("symmetric"); del edges
for F in (256, 500):
    B = synthetic.features(n, F, 1, "cuda"); C = torch.empty_like(B); s = A.struct(F)
    print("rmat23 F", F, round(t(lambda: ops.spmm_raw(s, n, B, out=C)), 3)); del B, C
n, e, w, c = synthetic.SHAPES["pubmed"]; adj = gnntf.graph2adj(synthetic.citation_graph(n, e, 0)); A = adj.normalized("symmetric")
fl = torch.empty(64*1024*1024, device="cuda")
for F in (500, 1433):
    B = synthetic.features(n, F, 1, "cuda"); C = torch.empty((n, F), device="cuda"); s = A.struct(F)
    def run(): fl.fill_(1.0); ops.spmm_raw(s, n, B, out=C)
    tf = t(lambda: fl.fill_(1.0)); print("pubmed F", F, round(t(run) - tf, 4))
