#!/bin/bash
# launch list (device time per kernel) of one call of the tensor-core pre-filter
mkdir -p gpurun_out
N=${1:-4000000}; D=${2:-768}; NQ=${3:-256}; K=${4:-10}; M=${5:-cosine}
timeout 300 python scripts/gpu_tc_bench.py $N $D $NQ $K $M 5 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv \
    python scripts/gpu_tc_bench.py $N $D $NQ $K $M 1 > gpurun_out/tc_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/tc_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
seq = [(r[ki].split('(')[0], float(r[vi].replace(',', ''))) for r in rows[1:]]
agg = collections.OrderedDict()
for name, v in seq: agg.setdefault(name, []).append(v)
for name, vs in agg.items(): print(f"{name[:60]:60s} n={len(vs):3d} total {sum(vs)/1e3:10.1f} us  max {max(vs)/1e3:10.1f} us")
g = [round(v/1e3, 1) for n_, v in seq if 'tc_gemm' in n_]; print("tc_gemm launches (us):", g[len(g)//2:])
g = [round(v/1e3, 1) for n_, v in seq if 'tc_refine' in n_]; print("tc_refine launches (us):", g[len(g)//2:])
PY
