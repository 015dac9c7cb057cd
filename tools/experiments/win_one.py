import sys
sys.path.insert(0, ".")
import lala_pc_b200 as L
from lala_pc_b200 import workloads as W
L.device_init(0)
net = W.config2()
t = L.Table(net.records, net.nvars)
for i in range(3):
    s = L.Store(values=net.store)
    r = L.fixpoint(t, s, mode=L.MODE_SWEEP)
    print(r.as_dict())
