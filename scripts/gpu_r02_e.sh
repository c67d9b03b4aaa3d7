#!/bin/bash
# GPU run E of round 2 (1 GPU): recipe sweep at full size on the device path, the whole GPU test suite, bench lines
# of configs 1 / 2 / 3 (3 on one GPU: chunked deferred blend).
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python tests/explore_confident.py > $O/r02e_confident.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_north_star.py > $O/r02e_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02e_bench.json 2> $O/r02e_bench.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > $O/r02e_bench_c2.json 2> $O/r02e_bench_c2.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > $O/r02e_bench_c3.json 2> $O/r02e_bench_c3.err
ls -la $O | tail -8
