#!/bin/bash
# GPU run F of round 2 (1 GPU): north-star tests with the pinned recipe, seed sweep for margin, whole GPU suite,
# bench lines (configs 1 with the pipelined e2e, 2 with a launch list, 3 with the per-layer profile).
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_north_star.py -q -s > $O/r02f_north_star.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_north_star.py > $O/r02f_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02f_bench.json 2> $O/r02f_bench.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > $O/r02f_bench_c2.json 2> $O/r02f_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_c2_launches.csv \
   python bench.py --config 2 --steps 1 --warmup 3 > $O/r02f_c2_ncu.log 2>&1
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > $O/r02f_bench_c3.json 2> $O/r02f_bench_c3.err
timeout 900 python tests/explore_confident.py > $O/r02f_confident.log 2>&1
ls -la $O | tail -8
