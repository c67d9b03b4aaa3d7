#!/bin/bash
# GPU run H of round 2 (8 GPUs): real multi-GPU tests (world 2/4/8), weak-scaling bench lines at N = 1, 2, 4, 8 (owned
# windows + NVLink peer push), configs[3] strong scaling at N = 1, 2, 4, 8.
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi topo -m > $O/r02h_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q > $O/r02h_multi.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-profile > $O/r02h_bench_n1.json 2> $O/r02h_bench_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-profile > $O/r02h_bench_n$n.json 2> $O/r02h_bench_n$n.err
done
timeout 600 python bench.py --config 3 --steps 2 --warmup 1 --no-profile > $O/r02h_c3_n1.json 2> $O/r02h_c3_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n \
    bench.py --config 3 --gpus $n --steps 3 --warmup 1 --no-profile > $O/r02h_c3_n$n.json 2> $O/r02h_c3_n$n.err
done
ls -la $O | tail -14
