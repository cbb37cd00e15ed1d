"""Tiny end-to-end pass over every kernel, meant to run under compute-sanitizer."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_ffi as o
from neumann_b200 import DeviceIndex

def same(got, exp):
    return np.array_equal(got[0], exp[0]) and np.array_equal(got[1].view(np.uint32), exp[1].view(np.uint32))

ok = True
for (n, d) in [(3000, 100), (1500, 64), (700, 13)]:
    rows = o.fill_synthetic(n, d, 1)
    idx = DeviceIndex(d)
    idx.load(rows)
    qs = o.fill_synthetic(9, d, 2)
    for m in ("cosine", "euclidean", "dot"):
        for k in (10, 1500):
            (g,) = idx.search(qs[0], k, m)
            ok &= same(g, o.search(rows, qs[0], k, m))
        res = idx.search(qs, 7, m)                       # batched kernels (nq = 9)
        for i in range(9):
            ok &= same(res[i], o.search(rows, qs[i], 7, m))
        mask = (np.arange(n) % 3) == 0
        (g,) = idx.search_masked(qs[1], 5, m, mask)
        sub = np.nonzero(mask)[0]
        er, es = o.search(rows[sub], qs[1], 5, m)
        ok &= np.array_equal(g[0], sub[er.astype(np.int64)].astype(np.uint64))
    # r02: device-evaluated filter over metadata columns, column moves on swap-remove
    from neumann_b200._ffi import NM_C_GE, NM_C_LT, NM_F_AND, NM_F_CMP, NM_F_EXISTS, NM_F_OR, NM_V_FLOAT, NM_V_INT, NmFilterOp
    import struct
    vals = (np.arange(n) * 7 % 11).astype(np.uint64)
    tags = np.where(np.arange(n) % 5 == 0, 0, 3).astype(np.uint8)          # every 5th row: field missing
    idx.column_set(1, 0, tags, vals)
    prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=4),
            NmFilterOp(kind=NM_F_CMP, cmp=NM_C_GE, lit_tag=NM_V_FLOAT, column=1,
                       lit=struct.unpack("<Q", struct.pack("<d", 9.5))[0]),
            NmFilterOp(kind=NM_F_OR), NmFilterOp(kind=NM_F_EXISTS, column=1), NmFilterOp(kind=NM_F_AND)]
    keep = (tags == 3) & ((vals < 4) | (vals >= 10))
    ok &= np.array_equal(idx.filter_mask(prog), keep)
    # a stack 40 deep takes the 64-bit-stack instance of the kernel
    deep = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=9 - (j % 2)) for j in range(40)] \
        + [NmFilterOp(kind=NM_F_AND)] * 39
    ok &= np.array_equal(idx.filter_mask(deep), (tags == 3) & (vals < 8))
    sub = np.nonzero(keep)[0]
    for m in ("cosine", "euclidean"):
        res = idx.search_filtered(qs[:2], 6, m, prog)
        for i in range(2):
            er, es = o.search(rows[sub], qs[i], 6, m)
            ok &= np.array_equal(res[i][0], sub[er.astype(np.int64)].astype(np.uint64))
    # r02: pipelined asynchronous device-resident searches (programmatic dependent launch)
    import torch
    stream = torch.cuda.Stream()
    dq = torch.from_numpy(qs).cuda()
    d_r = torch.zeros((9, 5), dtype=torch.int64, device="cuda")
    d_s = torch.zeros((9, 5), dtype=torch.float32, device="cuda")
    d_c = torch.zeros(9, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    idx.set_pipelining(True)
    for i in range(9):
        idx.search_device(dq[i].data_ptr(), 1, 5, "cosine", d_r[i].data_ptr(), d_s[i].data_ptr(),
                          d_c[i].data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    idx.set_pipelining(False)
    for i in range(9):
        er, es = o.search(rows, qs[i], 5, "cosine")
        ok &= np.array_equal(d_r[i].cpu().numpy().astype(np.uint64), er)
    idx.release_stream(stream.cuda_stream)
    idx.set_prefilter(1)                                   # int8 pre-filter kernels
    for m in ("cosine", "dot", "euclidean"):
        (g,) = idx.search(qs[4], 10, m)
        ok &= same(g, o.search(rows, qs[4], 10, m))
    idx.set_prefilter(0)
    idx.update(5, qs[3]); idx.swap_remove(7); idx.append(rows[:10])
    (g,) = idx.search(qs[3], 3, "cosine")
    ok &= g[0][0] == 5
    idx.close()
print("SANITIZE_DRIVER", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
