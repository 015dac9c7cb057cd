#!/bin/bash
# config 2 single fixpoint, both schedules, L2 flushed between runs (the bench's own single_fixpoint section), 3 times
for i in 1 2 3; do
  timeout 300 python - <<'PY'
import json, sys, argparse
sys.path.insert(0, ".")
import bench
import torch
torch.cuda.set_device(0)
a = argparse.Namespace(steps=10, warmup=3, no_cpu_baseline=True, scale=1.0)
r = bench.own_single(a, 0)
if r is None:
    print([n for n in dir(bench) if n.startswith("own")])
else:
    t = r["time_to_fixpoint"]
    print("c2 dense %.1f us  auto %.1f us | c1 dense %.1f auto %.1f" % (t["dense_ms"] * 1e3, t["auto_ms"] * 1e3, t["config1"]["dense_ms"] * 1e3, t["config1"]["auto_ms"] * 1e3))
PY
done
