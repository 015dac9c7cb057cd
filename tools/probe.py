"""GPU probe: quick timings of the hot path on the BASELINE.json shapes (development aid, not the bench)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W

L.device_init(0)
out = {}


def time_fix(net, label, reps=5):
    t = L.Table(net.records, net.nvars)
    for name, mode in (("sweep", L.MODE_SWEEP), ("auto", L.MODE_AUTO), ("worklist", L.MODE_WORKLIST)):
        best = None
        for _ in range(reps):
            s = L.Store(values=net.store)
            r = L.fixpoint(t, s, mode=mode)
            if best is None or r.device_ms < best.device_ms:
                best = r
        d = best.as_dict()
        d["gded_per_s"] = d["deductions"] / d["device_ms"] / 1e6
        d["us_per_sweep"] = d["device_ms"] * 1e3 / max(1, d["sweeps"])
        out[f"{label}.{name}"] = d
        print(label, name, json.dumps(d), flush=True)


t0 = time.time()
time_fix(W.config1(), "c1")
net2 = W.config2()
print("gen c2", time.time() - t0, flush=True)
time_fix(net2, "c2", reps=3)

# batched
from oracle import oracle as O
net = W.config4_base()
root, st = O.pir_fixpoint(net.store, net.records)
dec, obj = W.eps_decisions(net.records, root)
t = L.Table(net.records, net.nvars)
for n in (4096, 65536):
    b = L.Batch(t, n)
    best = None
    for _ in range(3):
        b.init_split(root, dec, 0)
        r = b.fixpoint(objective_var=obj)
        if best is None or r.device_ms < best.device_ms:
            best = r
    d = best.as_dict()
    d["gded_per_s"] = d["deductions"] / d["device_ms"] / 1e6
    out[f"c4.{n}"] = d
    print("c4", n, json.dumps(d), flush=True)
    b.close()
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
