#!/bin/bash
# One GPU-box pass: pytest -m gpu, smoke, bench (both arms), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
# launch list (cold-cache, serialised): SHARES only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the scan kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 \
   -o gpurun_out/scan_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
cat gpurun_out/pytest_gpu.log | tail -15
cat gpurun_out/smoke.log | tail -3
cat gpurun_out/bench.json gpurun_out/bench_reference.json
tail -3 gpurun_out/bench.err
