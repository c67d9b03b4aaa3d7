#!/bin/bash
# GPU run C of round 2: clean traces / bench lines of the channel-streamed kernel (4 issuers), N / G variants,
# confident-checkpoint recipe sweep (label smoothing) -- everything sequential (nothing shares the GPU with a bench).
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python tests/diag_tc_layers.py cs > $O/r02c_cs_layers.log 2>&1
SGM_TRACE=1 timeout 300 python tests/prof_forward.py fwd 125 1 > $O/r02c_trace.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02c_bench.json 2> $O/r02c_bench.err
SGM_CS_N=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02c_bench_n128.json 2> $O/r02c_bench_n128.err
SGM_CS_G=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02c_bench_g2.json 2> $O/r02c_bench_g2.err
SGM_CS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02c_bench_nocs.json 2> $O/r02c_bench_nocs.err
timeout 900 python tests/explore_confident.py > $O/r02c_confident.log 2>&1
ls -la $O | tail -12
