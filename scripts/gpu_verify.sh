#!/bin/bash
# Round checkpoint (1 GPU): full GPU test suite, smoke, both bench arms exactly as the driver runs them.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02_bench_reference_n1.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
c = d.get("configs", {})
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "serial", round(d["serial_ms_per_step"], 4),
      "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"],
      "parity", d["parity"]["ok"], "cpu", round(d["cpu_baseline"]["value"], 2), "clocks", d["clocks"])
for k, v in c.items():
    if "error" in v: print(k, v); continue
    if k == "cfg4":
        dp = v["default_path"]
        print(k, dp["path"][:30], "device", round(dp["device_ms_per_batch"], 3), "e2e", round(dp["e2e_ms_per_batch"], 3),
              "roofline", round(dp.get("roofline", {}).get("frac", 0), 3), "identical", dp.get("identical_to_exact"),
              "exact ms", round(v["exact_batched_kernels"]["device_ms_per_batch"], 1), "parity", v["parity"]["ok"])
    else:
        print(k, "us", round(v["device_us_per_query"], 2), "serial", round(v["device_us_per_query_serial"], 2),
              "e2e us", round(v["e2e_us_per_query"], 1), "roofline", round(v["roofline"]["frac"], 3), "parity", v["parity"]["ok"])
print("filtered", {k: (round(v["first_call_vs_unfiltered_e2e"], 3), round(v["cached_vs_unfiltered_e2e"], 3)) for k, v in d.get("filtered", {}).items() if isinstance(v, dict)})
print("prefilter_int8", round(d["prefilter_int8"]["e2e_value"], 1), "batch_tc_int8 device ms", round(d["batch_tc_int8"]["device_ms_per_batch"], 3))
PY
tail -3 gpurun_out/r02_bench_n1.err
