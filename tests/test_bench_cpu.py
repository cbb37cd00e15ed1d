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
