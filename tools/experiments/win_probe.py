"""GPU tuning aid: dense fixpoint of configs 1 and 2 with the shared-memory window kernel on / off (LPC_WINDOW)."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import numpy as np
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    L.device_init(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, net in (("c1", W.config1()), ("c2", W.config2())):
        t = L.Table(net.records, net.nvars)
        ms = []
        for i in range(8):
            s = L.Store(values=net.store)
            flush.zero_()
            r = L.fixpoint(t, s, mode=L.MODE_SWEEP)
            if i >= 3: ms.append(r.device_ms)
        out[name] = dict(ms=round(float(np.mean(ms)), 4), sweeps=r.sweeps, gded=round(r.deductions / np.mean(ms) / 1e6, 1),
                         us_per_sweep=round(float(np.mean(ms)) * 1e3 / r.sweeps, 2))
        for div in (0, 4, 16, 64):
            ms = []
            for i in range(6):
                s = L.Store(values=net.store)
                flush.zero_()
                r = L.fixpoint(t, s, mode=L.MODE_AUTO, switch_div=div)
                if i >= 2: ms.append(r.device_ms)
            out[name + ".auto%d" % div] = dict(ms=round(float(np.mean(ms)), 4), sweeps=r.sweeps, dense=r.dense_sweeps,
                                               mded=round(r.deductions / 1e6, 2))
    print(json.dumps(out))
else:
    for win in sys.argv[1:] or ("0", "1"):
        env = dict(os.environ, LPC_WINDOW=win)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("LPC_WINDOW=" + win, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
