"""GPU tuning aid: A/B of environment switches on config 2 (one process per setting; the switches are read once per
process / table). Usage: python tools/ab_env.py "LPC_SMORDER=0 LPC_CARVE=0" "LPC_SMORDER=1 LPC_CARVE=0" ...
Each setting is run twice, interleaved, so box drift shows up as the spread between the two runs of one setting."""
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
        for i in range(8):
            s = L.Store(values=net.store)
            flush.zero_()
            r = L.fixpoint(t, s, mode=mode)
            if i >= 3: ms.append(r.device_ms)
        out[name] = dict(ms=round(float(np.mean(ms)), 4), sweeps=r.sweeps)
    print(json.dumps(out))
    sys.stdout.flush()
    os._exit(0)
else:
    settings = sys.argv[1:] or [""]
    for rep in range(2):
        for sset in settings:
            env = dict(os.environ)
            for kv in sset.split():
                k, v = kv.split("=")
                env[k] = v
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print(repr(sset), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
