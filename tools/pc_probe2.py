"""GPU probe: fixed launch cost vs per-sweep cost of the PC kernel (max_sweeps = 1..k)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
L.device_init(0)
for name, net in (("c3", W.config3()), ("c5", W.config5()), ("c3/20", W.config3(0.05))):
    t = L.PcTable(net.props, net.terms, net.nvars)
    row = []
    for k in (1, 2, 3, 4, 5, 6):
        best = 1e9
        for _ in range(5):
            s = L.Store(values=net.store)
            r = t.fixpoint(s, max_sweeps=k)
            best = min(best, r.device_ms)
        row.append((k, r.sweeps, round(best * 1e3, 1)))
    print(name, row, flush=True)
    # a store already at its fixpoint: one quiescent sweep, no atomics
    s = L.Store(values=net.store)
    t.fixpoint(s)
    best = min(t.fixpoint(s).device_ms for _ in range(5))
    print(name, "quiescent sweep us", round(best * 1e3, 1), flush=True)
