"""One workload, a few warm-up launches and ONE final launch whose result record goes to a JSON file: the process ncu
wraps in tools/profile_r02.sh (`-k regex:<kernel> -s <warm-ups> -c 1` captures exactly that final launch), so that the
counters of the capture can be divided by the deductions of the very launch they belong to.

  python tools/prof_one.py <eps_dense|eps_auto|resident_dense|c2_dense|c2_auto|c1_dense|pc_c3[_dense]|pc_c5[_dense]|pc_c5_bits[_dense]> [warmups] [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402


def main():
    what = sys.argv[1]
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_%s.json" % what)
    L.device_init(0)
    rec = {"what": what, "warmups": warm}
    if what.startswith("eps") or what.startswith("resident"):
        net = W.config4_base()
        table = L.Table(net.records, net.nvars)
        s = L.Store(values=net.store)
        L.fixpoint(table, s)
        root = s.read()
        dec, obj = W.eps_decisions(net.records, root, n=24)   # bench.py's decision list
        dec = dec[:16]
        mode = L.MODE_AUTO if what.endswith("auto") else L.MODE_SWEEP
        if what.startswith("eps"):
            e = L.Eps(table, 65536, survivor_cap=8192)
            e.upload(root, dec, first_id=0, n=65536)
            for _ in range(warm + 1):
                r = e.run(objective_var=obj, mode=mode)
        else:
            b = L.Batch(table, 65536)
            for _ in range(warm + 1):
                b.init_split(root, dec, 0)
                r = b.fixpoint(objective_var=obj, mode=mode)
        rec.update(r.as_dict())
    elif what.startswith("c2") or what.startswith("c1"):
        net = W.config2() if what.startswith("c2") else W.config1()
        table = L.Table(net.records, net.nvars)
        mode = L.MODE_AUTO if what.endswith("auto") else L.MODE_SWEEP
        ms = 1 if what.endswith("first") else 0      # c2_first: the first sweep alone
        for _ in range(warm + 1):
            s = L.Store(values=net.store)
            r = L.fixpoint(table, s, mode=mode, max_sweeps=ms)
        rec.update(r.as_dict())
    else:
        dense = what.endswith("_dense")
        what = what[:-6] if dense else what
        net = W.config3() if what == "pc_c3" else W.config5()
        bits = what.endswith("bits")
        t = L.PcTable(net.props, net.terms, net.nvars)
        for _ in range(warm + 1):
            s = L.Store(values=net.store)
            if bits:
                s.write_bits(L.nbit_from_intervals(net.store))
            r = t.fixpoint(s, bitset=bits, mode=L.MODE_SWEEP if dense else L.MODE_AUTO)
        rec.update(r.as_dict())
        rec["propagators"], rec["terms"] = len(net.props), len(net.terms)
    with open(out, "w") as f:
        json.dump(rec, f)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
