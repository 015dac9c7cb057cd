"""Hand-over point of the change-driven kernel on config 2 (lpc_fixpoint_opts.reserved = divisor d: flagged sweeps start once
fewer than n_groups / d groups changed in a sweep).  python tools/switch_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

L.device_init(0)
net = W.config2()
table = L.Table(net.records, net.nvars)
for d in (0, 2, 3, 4, 6, 8, 12, 16, 32, 64):
    best = None
    for _ in range(5):
        s = L.Store(values=net.store)
        r = L.fixpoint(table, s, mode=L.MODE_AUTO, switch_div=d)
        best = r if best is None or r.device_ms < best.device_ms else best
        s.close()
    print("d = %2d: %.1f us, %d iterations (%d dense), %d deductions" % (d, best.device_ms * 1e3, best.sweeps, best.dense_sweeps, best.deductions), flush=True)
