#!/bin/bash
# Round-2 ncu evidence (1 GPU): launch list of the default bench command, full capture of the
# scan kernel (-> profiles/traffic_r02.json), launch list + full capture of the tensor-core GEMM
# on BASELINE config 4.
mkdir -p gpurun_out
BENCH="python bench.py --steps 6 --warmup 3 --no-staging --no-prefilter --no-configs --no-parity --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench.csv $BENCH > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 2 \
    -o gpurun_out/r02_scan $BENCH > gpurun_out/r02_scan_ncu.log 2>&1
echo "scan capture rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_tc_launches_config4.csv \
    python scripts/gpu_tc_bench.py 10000000 1536 256 100 euclidean 1 > gpurun_out/r02_tc_launches.log 2>&1
echo "tc launch list rc=$?"
# the two bulk GEMM phases of the second call (launch index: prepare + 12 gemm per call)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_filter -s 16 -c 2 \
    -o gpurun_out/r02_tc_gemm python scripts/gpu_tc_bench.py 10000000 1536 256 100 euclidean 1 > gpurun_out/r02_tc_ncu.log 2>&1
echo "tc capture rc=$?"
ls -la gpurun_out/*.ncu-rep
