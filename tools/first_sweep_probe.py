"""Where does the first sweep of config 2 go? Device time of a fixpoint cut off after 1 and after 2 sweeps, for the whole
table and for each operator's records alone (dense schedule; best of 5; L2 warm).  python tools/first_sweep_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

L.device_init(0)
net = W.config2()
ops = {"all": None, "add": 2, "mul": 4, "leq": 48}
for name, sig in ops.items():
    recs = net.records if sig is None else net.records[net.records[:, 0] == sig]
    table = L.Table(recs, net.nvars)
    row = []
    for k in (1, 2, 3):
        best = None
        for _ in range(5):
            s = L.Store(values=net.store)
            r = L.fixpoint(table, s, mode=L.MODE_SWEEP, max_sweeps=k)
            best = r.device_ms if best is None else min(best, r.device_ms)
            s.close()
        row.append(best * 1e3)
    print("%-4s %8d records: 1 sweep %.1f us, 2 sweeps %.1f us (2nd: %.1f), 3 sweeps %.1f us (3rd: %.1f)"
          % (name, len(recs), row[0], row[1], row[1] - row[0], row[2], row[2] - row[1]), flush=True)
    table.close()
