#!/bin/bash
# GPU run B of round 2: layer parity of the channel-streamed persistent kernel (conv_cs.cu), training-length sweep of
# the confident checkpoint, traces and a per-layer bench profile with the new kernel.
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python tests/diag_tc_layers.py cs > $O/r02b_cs_layers.log 2>&1
timeout 300 python tests/diag_tc_layers.py torch > $O/r02b_torch_layers.log 2>&1
timeout 900 python tests/explore_confident.py > $O/r02b_confident.log 2>&1 &
EXP=$!
SGM_TRACE=1 SGM_DEBUG=1 timeout 300 python tests/prof_forward.py fwd 125 1 > $O/r02b_trace.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02b_bench.json 2> $O/r02b_bench.err
wait $EXP
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_bf16.py -x -q > $O/r02b_pytest.log 2>&1
ls -la $O | tail -8
