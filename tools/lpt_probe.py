"""How much of a batched step is tail? Run the EPS batch in id order, then again with the subproblems ordered by
decreasing sweep count of the first run (longest-processing-time-first: a near-optimal schedule for the dynamic claim
queue). python tools/lpt_probe.py [n_subproblems]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W, sharding

L.device_init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
net = W.config4_base()
table = L.Table(net.records, net.nvars)
s = L.Store(values=net.store)
L.fixpoint(table, s)
root = s.read()
dec, obj = W.eps_decisions(net.records, root, n=24)
dec = dec[:16]
ids = sharding.strong_shard_ids(0, 65536 // n, 65536) if n < 65536 else np.arange(n, dtype=np.int64)
e = L.Eps(table, n, survivor_cap=8192)
for mode_name, mode in (("dense", L.MODE_SWEEP), ("auto", L.MODE_AUTO)):
    def run(order):
        e.upload(root, dec, ids=ids[order])
        best = None
        for _ in range(4):
            r = e.run(objective_var=obj, mode=mode)
            best = r if best is None or r.device_ms < best.device_ms else best
        return best
    ident = np.arange(n)
    r0 = run(ident)
    sw = e.sweeps()
    lpt = np.argsort(-sw, kind="stable")
    r1 = run(lpt)
    r2 = run(lpt[::-1])
    work = float(sw.sum())
    print(f"{mode_name} n={n}: id order {r0.device_ms:.3f} ms | longest first {r1.device_ms:.3f} ms | shortest first {r2.device_ms:.3f} ms | "
          f"max sweeps {sw.max()}, mean {sw.mean():.2f}, stores with > 8 sweeps {(sw > 8).sum()}, their share of the sweeps {sw[sw > 8].sum() / work:.3f}")
