"""Threads per block of the search kernel on a small model (bench.py's backtracking workload), dense nodes.
LPC_SEARCH_TPB=<n> python tools/search_tpb_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

L.device_init(0)
net = W.pir_network(200, 500, seed=77, width=6)
table = L.Table(net.records, net.nvars)
s = L.Store(values=net.store)
L.fixpoint(table, s)
root = s.read()
dec, obj = W.eps_decisions(net.records, root, n=12, min_degree=2)
stores = W.eps_stores(root, dec, 0, 1 << len(dec))
width = root[:, 1].astype(np.int64) - root[:, 0]
bv = [int(v) for v in np.argsort(-width, kind="stable")]
batch = L.Batch(table, len(stores))
for cd in (False, True):
    best = None
    for _ in range(3):
        batch.write(stores)
        r, _per = batch.search(bv, objective_var=obj, max_nodes=2048, max_depth=96, change_driven=cd, want_per_store=False)
        best = r if best is None or r.device_ms < best.device_ms else best
    print("TPB %s %s: %.2f ms, %d nodes, %.1f M nodes/s" % (os.environ.get("LPC_SEARCH_TPB", "default"), "cd   " if cd else "dense", best.device_ms,
                                                          best.n_nodes, best.n_nodes / best.device_ms / 1e3), flush=True)
