"""GPU tuning aid: does the ORDER of the records inside an operator segment matter? The fixpoint does not depend on it
(DESIGN.md 2), the gathers' locality does. Times the dense fixpoint of config 2 with the reference's (op, y, x, z) sort
and with tiled orders (op, y / T, x / T, z / T, y, x, z)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
import torch
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
L.device_init(0)
net = W.config2()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
r = net.records.astype(np.int64)
orders = {"ref (op,y,x,z)": np.lexsort((r[:, 3], r[:, 1], r[:, 2], r[:, 0]))}
for T in (256, 1024, 4096):
    orders[f"tile {T}: (op,y/T,x/T,z/T)"] = np.lexsort((r[:, 2], r[:, 3] // T, r[:, 1] // T, r[:, 2] // T, r[:, 0]))
    orders[f"tile {T}: (op,y/T,z/T,x/T)"] = np.lexsort((r[:, 2], r[:, 1] // T, r[:, 3] // T, r[:, 2] // T, r[:, 0]))
out = {}
for name, perm in orders.items():
    recs = np.ascontiguousarray(net.records[perm])
    t = L.Table(recs, net.nvars)
    ms = []
    for i in range(6):
        s = L.Store(values=net.store)
        flush.zero_()
        res = L.fixpoint(t, s, mode=L.MODE_SWEEP)
        if i >= 2: ms.append(res.device_ms)
    out[name] = dict(ms=round(float(np.mean(ms)), 4), sweeps=res.sweeps)
    print(name, out[name], flush=True)
