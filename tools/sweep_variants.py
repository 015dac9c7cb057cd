"""GPU tuning aid: time the dense fixpoint of config 2 for each built (RPT, MINB) kernel variant (one process each,
the variant is chosen once per process from LPC_RPT / LPC_MINB)."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import numpy as np
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    L.device_init(0)
    net = W.config2()
    t = L.Table(net.records, net.nvars)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, mode in (("sweep", L.MODE_SWEEP), ("auto", L.MODE_AUTO)):
        ms = []
        for i in range(6):
            s = L.Store(values=net.store)
            flush.zero_()
            r = L.fixpoint(t, s, mode=mode)
            if i >= 2: ms.append(r.device_ms)
        out[name] = dict(ms=float(np.mean(ms)), sweeps=r.sweeps, gded=r.deductions / np.mean(ms) / 1e6)
    print(json.dumps(out))
else:
    for rpt, minb in ((2, 3), (1, 4), (11, 4), (11, 3), (12, 3), (12, 2)):
        env = dict(os.environ, LPC_RPT=str(rpt), LPC_MINB=str(minb))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(rpt, minb, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
