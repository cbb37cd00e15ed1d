"""Probe: fixed per-launch overhead of the scan kernel (small corpora) + kernel time vs k."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

def kernel_ms(idx, q_dev, k, metric, steps=50):
    dev = q_dev.device
    d_rows = torch.zeros((1, k), dtype=torch.int64, device=dev)
    d_scores = torch.zeros((1, k), dtype=torch.float32, device=dev)
    d_counts = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(5):
            idx.search_device(q_dev.data_ptr(), 1, k, metric, d_rows.data_ptr(), d_scores.data_ptr(), d_counts.data_ptr(), st.cuda_stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(steps):
            idx.search_device(q_dev.data_ptr(), 1, k, metric, d_rows.data_ptr(), d_scores.data_ptr(), d_counts.data_ptr(), st.cuda_stream)
        b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

d = 768
q = torch.from_numpy(synth_rows(1, d, 0x5EED1001)).cuda()
for n in (256, 148 * 256, 148 * 256 * 4, 148 * 256 * 33, 1_250_000, 2_500_000, 10_000_000):
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
    out = []
    for k in (1, 10, 100, 1000):
        out.append(f"k={k}: {kernel_ms(idx, q, k, 'cosine') * 1e3:8.1f} us")
    ideal = n * d * 4 / 7.3e12 * 1e6
    print(f"n={n:9d}  ideal@7.3TB/s {ideal:8.1f} us | " + " | ".join(out), flush=True)
    idx.close()
