#!/bin/bash
# GPU run G of round 2 (1 GPU): north-star tests (3500-step recipe), whole GPU suite, bench configs 1..4.
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_north_star.py -q -s > $O/r02g_north_star.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_north_star.py > $O/r02g_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02g_bench.json 2> $O/r02g_bench.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > $O/r02g_bench_c2.json 2> $O/r02g_bench_c2.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > $O/r02g_bench_c3.json 2> $O/r02g_bench_c3.err
timeout 900 python bench.py --config 4 --steps 2 --warmup 2 > $O/r02g_bench_c4.json 2> $O/r02g_bench_c4.err
ls -la $O | tail -8
