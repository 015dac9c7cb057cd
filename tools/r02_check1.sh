#!/bin/bash
# First GPU pass of round 2: new kernels first (under their own time limits), then the whole suite, smoke, bench, A/B.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
(time timeout 900 python -m pytest tests/test_gpu_eps.py -m gpu -q --timeout 600 2>&1 | tail -25) 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_eps.py 2>&1 | tail -15) 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -5 gpurun_out/bench1.err
python tools/summarize_bench.py gpurun_out/bench1.json
echo "== A/B old grouped kernel vs new (resident images, dense)"
LPC_BATCH_V2=0 timeout 300 python tools/prof_one.py resident_dense 2 /tmp/a.json | tail -1
timeout 300 python tools/prof_one.py resident_dense 2 /tmp/b.json | tail -1
for w in eps_dense eps_auto c2_dense c2_auto c1_dense; do timeout 300 python tools/prof_one.py $w 3 /tmp/$w.json | tail -1; done
