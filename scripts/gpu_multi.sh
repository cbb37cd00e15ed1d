#!/bin/bash
# Multi-GPU pass: NCCL + in-process shard tests, then torchrun bench at N=WORLD (strong scaling).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_multi_$N.log
cat gpurun_out/pytest_multi_$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    fi
    tail -c 1500 gpurun_out/scale_$n.json; tail -5 gpurun_out/scale_$n.err
  fi
done
if [ $N -ge 8 ]; then
  # A/B: same strong-scaling point through ncclAllGather + merge kernel instead of the fused exchange
  NM_DISABLE_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/scale_8_nccl.json 2> gpurun_out/scale_8_nccl.err
  tail -c 600 gpurun_out/scale_8_nccl.json
  # BASELINE config 5: 80M x 768 over 8 GPUs (10M rows per GPU)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus 8 --steps 100 --warmup 10 --scaling weak > gpurun_out/weak_8.json 2> gpurun_out/weak_8.err
  tail -c 1200 gpurun_out/weak_8.json; tail -3 gpurun_out/weak_8.err
fi
