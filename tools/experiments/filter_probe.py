"""How many deduce() evaluations of a config-4 subproblem are needed at all? Sequential Gauss-Seidel over the propagators
that are not entailed on the root, evaluating a record only if one of its operands changed since the record's last
evaluation (a per-variable "changed in sweep k" stamp) - against evaluating every record in every sweep.
CPU only (the checker's per-record deduce).  python tools/experiments/filter_probe.py [n_subproblems]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from lala_pc_b200 import workloads as W  # noqa: E402

n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 48
net = W.config4_base()
root, _ = O.pir_fixpoint(net.store, net.records)
dec, obj = W.eps_decisions(net.records, root, n=24)
dec = dec[:16]
recs = np.ascontiguousarray(net.records, dtype=np.int32)
live = [i for i in range(len(recs)) if not O.pir_ask(root, recs[i])]
print("live records", len(live), "of", len(recs))
rng = np.random.default_rng(1)
ids = rng.integers(0, 1 << 16, n_sub)
stores = W.eps_stores(root, dec, 0, n_sub, ids=ids.astype(np.int64))
tot_dense = tot_filt = 0
hist = {}
for k in range(n_sub):
    s = stores[k].copy()
    stamp = np.zeros(net.nvars, dtype=np.int64)      # sweep in which the variable last changed (0 = never)
    stamp[dec] = 1                                   # the halved decision variables "changed in sweep 0 + 1"
    last_eval = np.zeros(len(recs), dtype=np.int64)  # sweep of the record's last evaluation... (time = sweep * N + position)
    t_changed = np.zeros(net.nvars, dtype=np.int64)
    t_changed[dec] = 1
    t_eval = np.zeros(len(recs), dtype=np.int64)
    bot, sweep, dense, filt, clock = False, 0, 0, 0, 1
    per = []
    while True:
        sweep += 1
        changed_any, n_eval = False, 0
        for i in live:
            clock += 1
            x, y, z = recs[i, 1], recs[i, 2], recs[i, 3]
            if max(t_changed[x], t_changed[y], t_changed[z]) <= t_eval[i]:
                continue
            n_eval += 1
            t_eval[i] = clock
            before = s[[x, y, z]].copy()
            s, ch, bot = O.pir_deduce(s, recs[i], bot)
            if ch:
                changed_any = True
                for v in (x, y, z):
                    if not np.array_equal(s[v], before[[x, y, z].index(v)]):
                        t_changed[v] = clock
            if bot:
                break
        dense += len(live)
        filt += n_eval
        per.append(n_eval)
        if bot or not changed_any:
            break
    tot_dense += dense
    tot_filt += filt
    hist.setdefault(sweep, []).append(per)
    print("subproblem %5d: %2d sweeps, %s, evaluations needed per sweep %s" % (ids[k], sweep, "bot" if bot else "ok ", per[:12]), flush=True)
print("dense evaluations %d, needed %d (%.2f)" % (tot_dense, tot_filt, tot_filt / tot_dense))
