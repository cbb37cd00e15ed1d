#!/bin/bash
# Round checkpoint: full GPU test suite + both bench arms, logs into gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench.json; python -c "import json; d=json.load(open(\"gpurun_out/bench.json\")); print(json.dumps(d.get(\"batch_tc_int8\")))"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/bench_ref.json
