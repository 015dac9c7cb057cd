import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
use_torch = len(sys.argv) > 1 and sys.argv[1] == "torch"
if use_torch:
    import torch
    torch.cuda.set_device(0)
    x = torch.zeros(10, device="cuda")
L.device_init(0)
net = W.config4_base()
t = L.Table(net.records, net.nvars)
s = L.Store(values=net.store); L.fixpoint(t, s); root = s.read()
for n_dec in (16, 24):
    dec, obj = W.eps_decisions(net.records, root, n=n_dec)
    dec = dec[:16]
    b = L.Batch(t, 65536)
    ms = []
    for i in range(5):
        b.init_split(root, dec, 0)
        r = b.fixpoint(objective_var=obj)
        ms.append(round(r.device_ms, 2))
    print("torch" if use_torch else "plain", "n_dec", n_dec, "ms", ms, "bot", r.n_bot, "sweeps", r.sweeps_total, "max", r.max_sweeps_seen, flush=True)
    b.close()
