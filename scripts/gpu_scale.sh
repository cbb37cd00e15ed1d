#!/bin/bash
# 8-GPU box: multi-GPU parity tests, then bench.py at N = 1, 2, 4, 8 exactly as the driver launches it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02_pytest_multi_8gpu.log 2>&1
echo "multi pytest rc=$?"; tail -3 gpurun_out/r02_pytest_multi_8gpu.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-prefilter --no-staging > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  echo "n$N rc=$?"; tail -c 300 gpurun_out/r02_scale_n$N.err
done
NM_DISABLE_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-parity > gpurun_out/r02_scale_n8_nccl_ab.json 2> gpurun_out/r02_scale_n8_nccl_ab.err; echo "n8 nccl a/b rc=$?"
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "ERR", e); continue
    base = base or d["value"]
    w = (d.get("configs") or {}).get("cfg5_weak") or {}
    print(f"N={n} value {d['value']:.1f} QPS ({d['ms_per_step']:.4f} ms, serial {d.get('serial_ms_per_step') or 0:.4f}) "
          f"speedup {d['value']/base:.2f} eff {d['value']/base/n:.3f} e2e {d['e2e']['value']:.1f} roofline {d['roofline']['frac']:.3f} "
          f"parity {d.get('parity', {}).get('ok')} | weak {w.get('value')} parity {(w.get('parity') or {}).get('ok')}")
try:
    d = json.loads(open("gpurun_out/r02_scale_n8_nccl_ab.json").read().strip().splitlines()[-1])
    print(f"N=8 NCCL all-gather A/B: value {d['value']:.1f} QPS e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print("ab ERR", e)
PY
