#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6) 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in eps_dense eps_auto resident_dense; do timeout 300 python tools/prof_one.py $w 3 /tmp/$w.json | tail -1 | cut -c1-330; done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err
python tools/summarize_bench.py gpurun_out/bench2.json
