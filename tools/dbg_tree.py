import sys, numpy as np
sys.path.insert(0, '.')
import lala_pc_b200 as L
from lala_pc_b200 import pcflat
from oracle import oracle as O
L.device_init(0)
forms=[('ne', ('sub', ('cdiv', ('var', 1), ('abs', ('var', 2))), ('sub', ('abs', ('var', 0)), ('var', 0))), ('tdiv', ('max', ('var', 1), ('abs', ('var', 8))), ('fdiv', ('var', 4), ('var', 6)))),
('ne', ('ediv', ('const', 3), ('var', 6)), ('const', 4)),
('and', ('ae', 'eq', 2, 2), ('le', ('min', ('add', ('var', 3), ('var', 8)), ('min', ('var', 7), ('var', 5))), ('const', 0))),
('or', ('gt', ('tdiv', ('cdiv', ('var', 5), ('const', 4)), ('const', -4)), ('mul', ('const', -1), ('var', 3))), ('ne', ('mul', ('const', -1), ('var', 2)), ('sub', ('mul', ('var', 2), ('var', 1)), ('const', -2)))),
('equiv', ('lit', 2), ('le', ('sum', ('mul', ('const', -1), ('var', 7)), ('mul', ('const', -1), ('var', 3)), ('mul', ('const', 2), ('var', 8)), ('mul', ('const', -1), ('var', 5)), ('var', 1)), ('const', 19))),
('equiv', ('lit', 3), ('le', ('mul', ('const', -1), ('var', 2)), ('const', -3))),
('le', ('add', ('mul', ('const', 3), ('var', 8)), ('var', 0)), ('const', 8))]
store=np.array([[0, 5], [5, 11], [-6, 7], [-6, 10], [9, 10], [0, 1], [0, 6], [0, 1], [0, 1]],dtype=np.int32)
props, terms = pcflat.flatten(forms)
m = O.PCModel(forms)
for sub in [forms] + [[f] for f in forms]:
    p, t = pcflat.flatten(sub)
    tab = L.PcTable(p, t, len(store))
    s = L.Store(values=store)
    r = tab.fixpoint(s)
    want, st = O.PCModel(sub).fixpoint(store)
    print("subset", [f[0] for f in sub], "device bot", r.is_bot, "sweeps", r.sweeps, s.read().tolist(), "| oracle bot", st.is_bot, want.tolist())
# step by step on the device, Gauss-Seidel order
tab = L.PcTable(props, terms, len(store))
s = L.Store(values=store)
cur, bot = store.copy(), False
for sweep in range(3):
    for i in range(len(forms)):
        cur, changed, bot = m.deduce(i, cur, bot)
        c = tab.deduce(s, i)
        print(sweep, i, "dev", c, s.read().tolist(), "| oracle", changed, cur.tolist(), bot)
