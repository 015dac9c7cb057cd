"""Per-sweep cost of a single-store fixpoint: time of the fixpoint cut off after k sweeps (opts.max_sweeps), k = 1 .. n,
best of a few runs each; the differences are the sweeps' costs.  python tools/sweep_profile.py [c2|c1] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

L.device_init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = W.config2() if which == "c2" else W.config1()
table = L.Table(net.records, net.nvars)
for name, mode in (("sweep", L.MODE_SWEEP), ("auto", L.MODE_AUTO), ("worklist", L.MODE_WORKLIST)):
    full = L.fixpoint(table, L.Store(values=net.store), mode=mode)
    prev, rows = 0.0, []
    for k in range(1, full.sweeps + 1):
        best = None
        for _ in range(reps):
            s = L.Store(values=net.store)
            r = L.fixpoint(table, s, mode=mode, max_sweeps=k)
            best = r if best is None or r.device_ms < best.device_ms else best
            s.close()
        rows.append((k, best.device_ms * 1e3, (best.device_ms - prev) * 1e3, best.deductions))
        prev = best.device_ms
    print(which, name, "total %.1f us, %d sweeps (%d dense), %d deductions" % (full.device_ms * 1e3, full.sweeps, full.dense_sweeps, full.deductions))
    print("   k: cumulative us / this sweep us / cumulative deductions")
    print("   " + "  ".join("%d:%.0f/%.1f/%.1fM" % (k, c, d, n / 1e6) for k, c, d, n in rows), flush=True)
