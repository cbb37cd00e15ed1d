"""Config 4 style measurement: batched vs one-scan-per-query, device-resident."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

def run(n, d, nq, k, metric, batching, reps=3):
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001); idx.set_batching(batching)
    q = torch.from_numpy(synth_rows(nq, d, 0x5EED1001)).cuda()
    r = torch.zeros((nq, k), dtype=torch.int64, device="cuda"); s = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    c = torch.zeros(nq, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        idx.search_device(q.data_ptr(), nq, k, metric, r.data_ptr(), s.data_ptr(), c.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            idx.search_device(q.data_ptr(), nq, k, metric, r.data_ptr(), s.data_ptr(), c.data_ptr(), st.cuda_stream)
        b.record(st)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out = (r.cpu().numpy().copy(), s.cpu().numpy().copy())
    idx.close()
    return ms, out

for (n, d, nq, k, metric) in [(10_000_000, 1536, 256, 100, "euclidean"), (10_000_000, 768, 256, 10, "cosine"),
                              (10_000_000, 768, 64, 10, "euclidean"), (10_000_000, 768, 16, 10, "dot")]:
    mb, ob = run(n, d, nq, k, metric, True)
    ms_, os_ = run(n, d, min(nq, 16), k, metric, False, reps=1)
    ms_full = ms_ * nq / min(nq, 16)
    same = np.array_equal(ob[0][:min(nq,16)], os_[0]) and np.array_equal(ob[1][:min(nq,16)].view(np.uint32), os_[1].view(np.uint32))
    gb = n * d * 4 / 1e9
    print(f"{n}x{d} {metric} nq={nq} k={k}: batched {mb:9.2f} ms = {nq/mb*1e3:8.1f} QPS | per-query scans {ms_full:9.2f} ms = {nq/ms_full*1e3:7.1f} QPS | speedup {ms_full/mb:5.2f}x | identical={same} | lane-ops/s {n*d*nq*(3 if metric=='euclidean' else 2)/mb/1e9:.1f} T", flush=True)
