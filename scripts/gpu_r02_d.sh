#!/bin/bash
# GPU run D of round 2 (2 GPUs): multi-GPU tests of the peer-memory seam exchange, 2-GPU bench lines (owned + p2p,
# owned + NCCL, slabs), then on single GPUs: strided 16-channel block on the channel-streamed kernel, recipe sweep.
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi topo -m > $O/r02d_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/r02d_multi.log 2>&1
for mode in owned owned-nccl slab; do
  SGM_MGPU=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/r02d_bench_n2_$mode.json 2> $O/r02d_bench_n2_$mode.err
done
CUDA_VISIBLE_DEVICES=1 timeout 600 python tests/explore_confident.py > $O/r02d_confident.log 2>&1 &
EXP=$!
SGM_CS_S2MIN=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02d_bench_s2min16.json 2> $O/r02d_bench_s2min16.err
SGM_CS_S2MIN=16 timeout 300 python tests/diag_tc_layers.py s2 > $O/r02d_s2_layers.log 2>&1
wait $EXP
ls -la $O | tail -12
