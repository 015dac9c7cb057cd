"""GPU probe of the PC kernels on BASELINE.json configs 3 and 5 (development aid and the command profiled by ncu)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W

L.device_init(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
out = {}
for name, net in (("c3", W.config3()), ("c5", W.config5())):
    t = L.PcTable(net.props, net.terms, net.nvars)
    for bitset in ((False, True) if name == "c5" else (False,)):
        cells = L.nbit_from_intervals(net.store) if bitset else None
        best = None
        for _ in range(reps):
            s = L.Store(values=net.store)
            if bitset:
                s.write_bits(cells)
            r = t.fixpoint(s, bitset=bitset)
            best = r if best is None or r.device_ms < best.device_ms else best
        d = best.as_dict()
        d["us_per_sweep"] = d["device_ms"] * 1e3 / max(1, d["sweeps"])
        d["gded_per_s"] = d["deductions"] / d["device_ms"] / 1e6
        bytes_per_sweep = 16 * len(net.props) + 16 * len(net.terms)
        d["algorithmic_gbs"] = bytes_per_sweep * d["sweeps"] / d["device_ms"] / 1e6
        out[name + (".bitset" if bitset else ".interval")] = d
        print(name, "bitset" if bitset else "interval", json.dumps(d), flush=True)
json.dump(out, open("gpurun_out/pc_probe.json", "w"), indent=1)
