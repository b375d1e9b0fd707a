"""bench.py prints ONE JSON line with the keys the driver reads (the contract in the task statement)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the reference's own CPU code (oracle/_ref, else the port) on a bounded sample"""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--n-bases", "2000000")
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "Gbases/s" and d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
def test_own_arm_line():
    # 20 Mb: at k = 12 more than half of the 4^12 k-mers must occur, or the median frequency is 0 and the
    # log2 score table holds +Inf (rejected by design)
    d = run_bench("--steps", "3", "--warmup", "3", "--n-bases", "20000000", "--e2e-steps", "2", "--config3-scale", "0.02")
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["data"] == "synthetic"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 20000000 and e["d2h_bytes_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] > 0
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    # the timed call is parity-gated against the oracle inside the run
    pc = d["parity_check"]
    assert pc["ok"] and pc["counts_bit_exact"] and pc["score_table_bit_exact"] and pc["spans_identical"]
    assert pc["rank_mode"]["ok"] and pc["call"] == "ks_dev_pipeline"
    assert {"rank_mode", "config1", "config5_subset"} <= set(d["extra"])
    assert d["extra"]["rank_mode"]["spans"] > 0 and d["extra"]["config1"]["ms_per_step"] > 0
    c3 = d["config3"]
    assert c3["parity_check"]["ok"] and c3["spans"] > 0 and c3["ms_per_step"] > 0 and c3["n_gpus"] == 1
