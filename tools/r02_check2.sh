#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_eps.py tests/test_gpu_pir.py -m gpu -q --timeout 600 2>&1 | tail -8) 2>&1
timeout 300 python tools/sweep_profile.py c2 3 2>&1 | tail -12
PROFILE_NO_LAUNCHES=1 bash tools/profile_r02.sh > gpurun_out/profile.log 2>&1
tail -3 gpurun_out/profile.log
cat gpurun_out/ncu_counters.json | head -120
