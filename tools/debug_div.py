import sys, ctypes
import numpy as np
sys.path.insert(0, ".")
import lala_pc_b200 as L
from oracle import oracle as O
L.device_init(0)
ops = dict(TDIV=25, FDIV=27, CDIV=29, EDIV=31, EQ=46, LEQ=48)
lo, hi = -5, 5
for name, op in ops.items():
    stats, fix = O.pir_exhaustive(op, lo, hi, name in ("EQ", "LEQ"), want_fixpoints=True)
    n = len(fix)
    vals = [(a, b) for a in range(lo, hi + 1) for b in range(a, hi + 1)]
    v = np.array(vals, dtype=np.int32); m = len(v)
    stores = np.zeros((n, 4, 2), dtype=np.int32)
    idx = np.arange(n)
    stores[:, 0] = v[idx // (m * m)]; stores[:, 1] = v[(idx // m) % m]; stores[:, 2] = v[idx % m]
    if name in ("EQ", "LEQ"):
        stores[:, 0, 0] = np.maximum(stores[:, 0, 0], 0); stores[:, 0, 1] = np.minimum(stores[:, 0, 1], 1)
    t = L.Table(np.array([[op, 0, 1, 2]], dtype=np.int32), 4)
    b = L.Batch(t, n); b.write(stores); res = b.fixpoint(); got = b.read(); flags = b.flags()
    want_bot = fix[:, 6].astype(bool)
    print(name, "bot mismatch", int(((flags & 1).astype(bool) != want_bot).sum()))
    ok = ~want_bot
    g = got[:, :3].reshape(-1, 6)
    bad = np.flatnonzero(ok & (g != fix[:, :6]).any(1))
    print("  store mismatches", len(bad))
    for i in bad[:8]:
        print("   in", stores[i, :3].reshape(-1).tolist(), "gpu", g[i].tolist(), "oracle", fix[i, :6].tolist())
    # same through the single-store kernel
    bad2 = 0
    for i in bad[:50]:
        s = L.Store(values=stores[i]); r = L.fixpoint(t, s, mode=L.MODE_SWEEP)
        if not np.array_equal(s.read()[:3].reshape(-1), fix[i, :6]): bad2 += 1
    print("  of the first 50, also wrong in the single-store kernel:", bad2)
