#!/bin/bash
# Sanitizers + ncu of the batched kernels + refreshed bench/ncu of the headline path.
mkdir -p gpurun_out
bash scripts/gpu_sanitize.sh
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-staging > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 \
   -o gpurun_out/scan_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-staging > gpurun_out/ncu_full.log 2>&1
# batched kernels on a 2M x 1536 slice (config 4 shape, smaller N to keep ncu replays short)
cat > /tmp/batch_ncu.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows
idx = DeviceIndex(1536); idx.fill_synthetic(2_000_000, 0x5EED0001)
q = synth_rows(64, 1536, 0x5EED1001)
for _ in range(2): idx.search(q, 100, "euclidean")
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"score_batch|select_batch" -s 2 -c 2 \
   -o gpurun_out/batch_full -f python /tmp/batch_ncu.py > gpurun_out/ncu_batch.log 2>&1
ls -la gpurun_out | head -40
