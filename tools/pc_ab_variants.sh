#!/bin/bash
# PC kernel: tile-loop variants (LPC_PC_VARIANT) under both schedules, best-of-3 per cell.
for v in 0 1 2 3 4; do
  for w in pc_c3 pc_c3_dense pc_c5 pc_c5_bits; do
    best=$(for i in 1 2 3; do LPC_PC_VARIANT=$v timeout 120 python tools/prof_one.py $w 4 /tmp/x.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['device_ms']*1e3,1), d['sweeps'])"; done | sort -n | head -1)
    echo "variant $v $w $best"
  done
done
