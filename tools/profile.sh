#!/bin/bash
# Run under gpurun: bench line, ncu launch list, ncu --set full captures of the hot kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pir_fixpoint -s 3 -c 2 -f -o gpurun_out/prof_fixpoint \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pc > gpurun_out/ncu_fix.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pir_batch -s 2 -c 1 -f -o gpurun_out/prof_batch \
  python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_batch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pc_fixpoint -s 1 -c 3 -f -o gpurun_out/prof_pc \
  python tools/pc_probe.py 2 > gpurun_out/ncu_pc.log 2>&1
ls -la gpurun_out
# summaries on the box (the three .ncu-rep files together exceed gpurun's 64 MiB return limit)
for k in fixpoint batch pc; do
  python tools/ncu_summary.py gpurun_out/prof_$k.ncu-rep > gpurun_out/r_${k}_ncu.txt 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv > /tmp/src_$k.csv 2>/dev/null
  python tools/sass_hot.py /tmp/src_$k.csv > gpurun_out/r_${k}_sass_hot.txt 2>&1
done
rm -f gpurun_out/prof_pc.ncu-rep gpurun_out/prof_fixpoint.ncu-rep   # the batch capture travels back for reference
ls -la gpurun_out
