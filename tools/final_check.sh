#!/bin/bash
# Run under gpurun: the round-end checks in one go - GPU test suite, smoke(), the default bench line and its summary.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/c1_probe.py 2>/dev/null | head -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
r = d["roofline"]
print("config 2:", round(d["value"] / 1e9, 1), "G ded/s,", round(d["ms_per_step"], 4), "ms, frac hbm", round(r["frac"], 3),
      "frac l2", round(r["frac_of_l2_copy_peak"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "ms, auto", round(d["latency"]["auto_ms"], 4))
b = d["batched"]
print("config 4:", round(b["value"] / 1e9, 1), "G ded/s,", round(b["ms_per_step"], 3), "ms, e2e", round(b["e2e"]["ms_per_step"], 2),
      "ms, e2e_split", round(b["e2e_split"]["ms_per_step"], 2), "ms")
print("pc:", {k: round(v["ms_per_fixpoint"], 4) for k, v in d["pc"].items()}, "search", round(d["search"]["ms"], 3), "ms")
print("launches", d["gpu_launches"], "clocks", d["clocks"], "cpu", round(d["cpu_baseline"]["fixpoint_ms"], 1), "ms")
PY
