#!/bin/bash
# GPU run A of round 2: north-star parity tests, weight-streaming micro-benchmark, phase traces and ncu captures of
# the brick kernel family (tc_conv_kernel) at the benched batch.
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_north_star.py -q -s > $O/r02a_north_star.log 2>&1
echo "north_star exit $?" >> $O/r02a_north_star.log
timeout 120 tests/ubench_wstream.bin > $O/r02a_wstream.txt 2>&1
SGM_TRACE=1 timeout 300 python tests/prof_forward.py fwd 125 1 > $O/r02a_trace.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 11 \
  -o $O/r02a_tc_full python tests/prof_forward.py fwd 125 1 > $O/r02a_ncu.log 2>&1
ls -la $O
