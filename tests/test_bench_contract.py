"""bench.py's contract where it can be checked without a GPU: the reference arm (`--impl reference`: the compiled, unmodified
reference timed on the host cores, or the oracle port where the reference tree is absent) runs end to end and prints ONE JSON
line with the keys the driver reads; our arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "cpu_baseline"]


def run_bench(*flags, timeout=900):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3", "--no-continuity-reference", "--no-exact-reference")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and not [k for k in REQUIRED if k not in d]
    assert d["metric"] == "sdf_queries_per_sec_256cubed_grid" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 1e5 and abs(d["value"] - 16777216 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]   # the full 256^3 grid per step
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "C2" in d["config"]["workload"] and "enoki" in d["config"]
    assert d["build"]["octree_c2"]["seconds"] > 0


def test_our_arm_has_no_cpu_fallback():
    r = run_bench("--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-config4", timeout=300)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert "CUDA" in (r.stderr + r.stdout)
