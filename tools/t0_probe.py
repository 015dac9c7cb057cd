"""Fixed cost of a single-store fixpoint launch (launch + memset + prologue scan + barriers), and the first sweep of config 2
with and without bound movements.  python tools/t0_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

L.device_init(0)
net = W.config2()


def best_of(table, store, mode, k, n=8):
    best = None
    s = L.Store(values=store)
    for _ in range(n):
        s.write(store)
        r = L.fixpoint(table, s, mode=mode, max_sweeps=k)
        best = r.device_ms if best is None else min(best, r.device_ms)
    return best * 1e3, r


for label, recs, nvars, store in (("8 records, 1M-var store", net.records[:8], net.nvars, net.store),
                                  ("8 records, 4k-var store", net.records[:8] % np.array([1 << 30, 4096, 4096, 4096]), 4096, net.store[:4096]),
                                  ("64k records, 1M-var store", net.records[:65536], net.nvars, net.store)):
    table = L.Table(np.ascontiguousarray(recs, dtype=np.int32), nvars)
    for mode, mn in ((L.MODE_SWEEP, "dense"), (L.MODE_AUTO, "auto")):
        print("%-34s %-5s 1 sweep: %.1f us" % (label, mn, best_of(table, store, mode, 1)[0]), flush=True)
table = L.Table(net.records, net.nvars)
s = L.Store(values=net.store)
L.fixpoint(table, s)
fix = s.read()
for mode, mn in ((L.MODE_SWEEP, "dense"), (L.MODE_AUTO, "auto")):
    a, ra = best_of(table, net.store, mode, 1)
    b, rb = best_of(table, fix, mode, 1)
    print("config 2 %-5s first sweep from the initial store %.1f us, one sweep on the fixpoint (nothing moves) %.1f us" % (mn, a, b), flush=True)
