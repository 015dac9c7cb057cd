"""GPU tuning aid: config 1 (10 k variables / 50 k propagators) time to fixpoint, grid kernels vs the cluster kernel
(LPC_CLUSTER is read per call)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
from oracle import oracle as O
L.device_init(0)
net = W.config1()
want, st = O.pir_fixpoint(net.store, net.records)
t = L.Table(net.records, net.nvars)
for env in ("0", "1", "0", "1"):
    os.environ["LPC_CLUSTER"] = env
    for mode, name in ((L.MODE_SWEEP, "sweep"), (L.MODE_AUTO, "auto")):
        ms = []
        for i in range(8):
            s = L.Store(values=net.store)
            r = L.fixpoint(t, s, mode=mode)
            if i >= 3: ms.append(r.device_ms)
        ok = np.array_equal(s.read(), want)
        print(f"LPC_CLUSTER={env} {name}: {np.mean(ms)*1e3:.1f} us, {r.sweeps} sweeps, {np.mean(ms)*1e3/r.sweeps:.2f} us/sweep, parity={ok}", flush=True)
