#!/bin/bash
# GPU run L of round 2 (1 GPU, last minute of budget): ncu launch list of the final code (gpu__time_duration per launch).
cd "$GRAFT_REPO_ROOT"
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-profile > gpurun_out/r02l_ncu.log 2>&1
echo "exit $?" >> gpurun_out/r02l_ncu.log
