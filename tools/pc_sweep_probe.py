"""Where does a PC fixpoint's time go? Device time of a launch capped at k sweeps (k = 1, 2, ...) from the initial store, and
of one confirming sweep on the fixpoint, for both schedules. python tools/pc_sweep_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

for name, net in (("c3", W.config3()), ("c5", W.config5())):
    t = L.PcTable(net.props, net.terms, net.nvars)
    for mode, mname in ((L.MODE_SWEEP, "dense"), (L.MODE_AUTO, "auto")):
        s = L.Store(values=net.store)
        row = []
        for k in (1, 2, 3, 4, 5, 6, 8, 0):
            best = None
            for _ in range(5):
                s.write(net.store)
                r = t.fixpoint(s, mode=mode, max_sweeps=k)
                best = r.device_ms if best is None else min(best, r.device_ms)
            row.append("k=%d: %.1f us (%d sw, %d ded)" % (k, best * 1e3, r.sweeps, r.deductions))
        best = min(t.fixpoint(s, mode=mode).device_ms for _ in range(5))
        row.append("at the fixpoint: %.1f us" % (best * 1e3))
        print(name, mname, " | ".join(row), flush=True)
