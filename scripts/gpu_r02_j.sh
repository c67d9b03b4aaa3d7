#!/bin/bash
# GPU run J of round 2 (1 GPU, last ~2 minutes of budget): the default bench line with the final code.
cd "$GRAFT_REPO_ROOT"
timeout 115 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
echo "exit $?" >> gpurun_out/r02j_bench.err
