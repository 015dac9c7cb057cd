#!/bin/bash
# Run under gpurun (one GPU): ncu launch list of the default bench command, one `ncu --set full` capture per hot kernel
# (each of exactly one launch whose deductions are recorded beside it), their summaries and profiles/ncu_counters.json.
mkdir -p gpurun_out
W=2
cap() {  # name kernel-regex
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $W -c 1 -f -o gpurun_out/prof_$1 \
    python tools/prof_one.py $1 $W gpurun_out/prof_$1.json > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1.ncu-rep > gpurun_out/r02_$1_ncu.txt 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
  python tools/sass_hot.py /tmp/src_$1.csv > gpurun_out/r02_$1_sass_hot.txt 2>&1
}
for spec in ${PROFILE_CAPS:-eps_dense:k_pir_group eps_auto:k_pir_group c2_dense:k_pir_fixpoint c2_auto:k_pir_dirty pc_c3:k_pc_fixpoint pc_c3_dense:k_pc_fixpoint}; do
  cap ${spec%%:*} ${spec##*:}
done
python tools/ncu_counters.py gpurun_out > gpurun_out/ncu_counters.json
if [ -z "$PROFILE_NO_LAUNCHES" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
fi
# only the summaries travel back (the .ncu-rep files together exceed gpurun's 64 MiB return limit)
mkdir -p /tmp/reps && mv gpurun_out/*.ncu-rep /tmp/reps/ 2>/dev/null
ls -la gpurun_out
