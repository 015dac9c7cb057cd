"""compute-sanitizer target: one grouped batch launch (k_pir_batch2) on a few thousand stores + one dense / change-driven
fixpoint + one PC fixpoint with tree propagators. Usage: compute-sanitizer --tool memcheck python tools/sanitize_batch.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W, pcflat
L.device_init(0)
net = W.config4_base()
t = L.Table(net.records, net.nvars)
s = L.Store(values=net.store)
L.fixpoint(t, s, mode=L.MODE_SWEEP)
root = s.read()
dec, obj = W.eps_decisions(net.records, root, n=12)
b = L.Batch(t, 4096)
b.init_split(root, dec, 0)
r = b.fixpoint(objective_var=obj)
print("batch", r.n_bot, r.n_solution, r.n_unknown)
s2 = L.Store(values=net.store)
print("auto", L.fixpoint(t, s2, mode=L.MODE_AUTO).sweeps)
forms = [("le", ("var", 0), ("add", ("const", -1), ("var", 1))), ("eq", ("min", ("var", 2), ("var", 3)), ("var", 4)),
         ("le", ("sum", ("var", 0), ("var", 1), ("var", 2)), ("const", 20))]
p, tm = pcflat.flatten(forms)
pt = L.PcTable(p, tm, 5)
ps = L.Store(values=np.array([[0, 10]] * 5, dtype=np.int32))
print("pc", pt.fixpoint(ps).sweeps)
