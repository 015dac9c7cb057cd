"""GPU tuning aid: per-operator sweep cost of the dense kernel. Config 2's records split by operator (same store),
each subset run to its own fixpoint; prints us/sweep and ns per 1000 records."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
import torch
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
L.device_init(0)
net = W.config2()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
names = {W.ADD: "add", W.MUL: "mul", W.LEQ: "leq"}
out = {}
for op in (None, W.ADD, W.MUL, W.LEQ):
    recs = net.records if op is None else net.records[net.records[:, 0] == op]
    t = L.Table(np.ascontiguousarray(recs), net.nvars)
    ms = []
    for i in range(5):
        s = L.Store(values=net.store)
        flush.zero_()
        r = L.fixpoint(t, s, mode=L.MODE_SWEEP)
        if i >= 2: ms.append(r.device_ms)
    m = float(np.mean(ms))
    out["all" if op is None else names[op]] = dict(n=len(recs), ms=round(m, 4), sweeps=r.sweeps,
        us_per_sweep=round(1e3 * m / r.sweeps, 2), ps_per_record=round(1e9 * m / r.sweeps / len(recs), 2))
print(json.dumps(out))
