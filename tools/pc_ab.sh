#!/bin/bash
# PC kernel timing (best of a few fixpoints per config), for A/B between builds.
for w in pc_c3 pc_c3_dense pc_c5 pc_c5_dense pc_c5_bits pc_c5_bits_dense; do
  for i in 1 2 3; do timeout 120 python tools/prof_one.py $w 4 /tmp/x.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', round(d['device_ms']*1e3,1), 'us', d['sweeps'], 'sweeps', d['deductions'], 'deductions')"; done
done
