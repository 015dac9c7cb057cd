#!/bin/bash
# compute-sanitizer memcheck and synccheck over the GPU tests that finish in minutes under them (run under gpurun, one GPU).
timeout 2000 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 8 python -m pytest tests/test_gpu_pc.py tests/test_search.py \
  tests/test_gpu_eps.py tests/test_gpu_pir.py -x -q -k "not config2_full and not exhaustive and not scale" 2>&1 | tail -8
timeout 2000 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 8 python -m pytest tests/test_gpu_pc.py tests/test_search.py \
  tests/test_gpu_eps.py tests/test_gpu_pir.py -x -q -k "not config2_full and not exhaustive and not scale" 2>&1 | grep -v "Host Frame" | tail -8
