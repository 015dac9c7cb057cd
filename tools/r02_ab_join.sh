#!/bin/bash
# A/B of the join variants of the grouped kernel (LPC_JOIN: 0 one region with six atomics, 1 region + predicated atomics,
# 2 predicated atomics without a region, 3 plain predicated stores, 4 plain predicated stores inside the region) on the config-4 EPS workload, then parity under 3.
mkdir -p gpurun_out
for j in 0 1 2 3 4; do
  for w in eps_dense eps_auto; do
    echo -n "LPC_JOIN=$j $w: "
    LPC_JOIN=$j timeout 300 python tools/prof_one.py $w 3 /tmp/x.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms, %d sweeps, %.1f G ded/s' % (d['device_ms'], d['sweeps_total'], d['deductions']/d['device_ms']/1e6))"
  done
done 2>&1 | tee gpurun_out/r02_ab_join.txt
LPC_JOIN=${PARITY_JOIN:-3} timeout 900 python -m pytest tests/test_gpu_eps.py -m gpu -q --timeout 600 2>&1 | tail -5
