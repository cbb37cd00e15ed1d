"""bench.py's reference arm (CPU oracle port) — host-only checks of its two corpus modes."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(mode, extra_env=None):
    env = dict(os.environ, NM_BENCH_REF_MODE=mode, **(extra_env or {}))
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--rows", "60000",
                          "--dim", "96", "--steps", "3", "--warmup", "1"], env=env, check=True,
                         capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def test_reference_arm_scores_the_full_corpus_in_both_modes():
    res = _run("resident")
    chk = _run("chunked", {"NM_BENCH_REF_CHUNK": "25000"})
    for line in (res, chk):
        assert line["impl"] == "reference" and line["gpu_launches"] == 0
        assert line["config"]["rows"] == 60000 and line["steps"] == 3 and line["warmup"] == 1
        assert line["cpu_baseline"]["kind"] == "port" and "ALL 60,000 rows" in line["cpu_baseline"]["sample"]
        assert line["e2e"]["value"] == line["value"] > 0
        # the printed step time is a measured one: steps x ms_per_step == the timed region
        assert abs(line["ms_per_step"] * 3e-3 - line["timed_region_s"]) < 1e-6
    assert res["config"]["corpus"] == "resident" and chk["config"]["corpus"] == "chunked"
    # same query (step 3 of 4 -> query 3), same corpus: identical top-k either way
    assert res["last_result_rows"] == chk["last_result_rows"]


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """The driver launches the reference arm like the GPU arm — for N > 1 through torchrun.  Rank 0
    alone runs it and prints ONE JSON line (n_gpus as asked); the other ranks exit 0 without work."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "bench.py"),
           "--impl", "reference", "--gpus", "2", "--rows", "40000", "--dim", "64", "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, env=dict(os.environ, NM_BENCH_REF_MODE="resident"), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["gpu_launches"] == 0
    assert line["config"]["rows"] == 40000 and line["value"] > 0 and line["e2e"]["value"] == line["value"]


# ---- the parity machinery of the GPU arm, on CPU: chunked oracle per shard + merge over ranks ----
class _FakeShard:
    """Stands in for a DeviceIndex shard: get_rows() serves the synthetic corpus from the host."""

    def __init__(self, dim, seed, global_lo):
        self.dim, self.seed, self.lo = dim, seed, global_lo

    def get_rows(self, first, n, out=None):
        import numpy as np
        import oracle_ffi as o
        rows = o.fill_synthetic(n, self.dim, self.seed, row_offset=self.lo + first, threads=2)
        if out is None:
            return rows
        out.reshape(-1)[:n * self.dim] = rows.reshape(-1)
        return out.reshape(-1)[:n * self.dim].reshape(n, self.dim)


def _parity_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import bench
    import oracle_ffi as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bench.CHUNK_ROWS = 7_000                       # several chunks per shard
        total, dim, k, nq = 50_001, 48, 10, 5
        lo, hi = total * rank // world, total * (rank + 1) // world
        qs = o.fill_synthetic(nq, dim, bench.SEED_QUERY)
        full = o.fill_synthetic(total, dim, bench.SEED_ROWS, threads=4)
        good = [o.search(full, qs[i], k, "cosine", threads=4) for i in range(nq)]
        bad = [(r.copy(), s.copy()) for r, s in good]
        bad[3][0][[4, 5]] = bad[3][0][[5, 4]]          # two ranks swapped in one query
        shard = _FakeShard(dim, bench.SEED_ROWS, lo)
        out = []
        for res in (good, bad):
            # (oracle_shard_topk takes chunk_rows as a default argument: pass the small one)
            orig = bench.oracle_shard_topk
            bench.oracle_shard_topk = lambda *a, **kw: orig(*a, chunk_rows=7_000, **kw)
            try:
                p = bench.parity_check(shard, hi - lo, lo, total, qs, k, "cosine",
                                       {"nm_search": res, "nm_search_device": good}, bench.SEED_ROWS,
                                       world, rank, "cpu test")
            finally:
                bench.oracle_shard_topk = orig
            out.append(p)
        q.put((rank, out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_bench_parity_pass_detects_a_wrong_rank_across_two_gloo_ranks():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_parity_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert isinstance(res[0], list), res[0]
    assert res[1] == [None, None]                       # only rank 0 reports
    ok, wrong = res[0]
    assert ok["ok"] and ok["ids_equal"] and ok["score_bits_equal"] and ok["generator_matches_host_twin"]
    assert ok["oracle_rows"] == 50_001 and ok["shards"] == 2 and ok["queries"] == 5
    assert not wrong["ok"] and not wrong["ids_equal"]
    assert wrong["mismatches"][0]["query"] == 3 and wrong["mismatches"][0]["path"] == "nm_search"
