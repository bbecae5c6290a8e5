"""bench.py contract checks that do not need a GPU: the reference arm (the reference's own CPU implementation of the path,
oracle/_ref) prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtds_ref.so not built (needs /root/reference)")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3   # both arms clamp the warm-up to >= 3 (timing rules)
    assert d["metric"].startswith("env-steps/sec") and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["host"]["usable_cores"] >= d["cpu_baseline"]["cores"]
    assert d["cpu_baseline_codegen"]["value"] > 0 and set(d["config"]["threads_sweep_templated"]) >= {"1"}


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtds_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("config", ["cartpole64", "pendulum5_fd", "sphere2_16384", "humanoid4096"])
def test_reference_arm_of_every_config(config):
    """bench.py --impl reference --config <c>: one JSON line per BASELINE configuration, bounded sample, same keys."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config, "--steps", "2",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["cores"] == 1 and d["e2e"]["value"] == d["value"]
    assert "BASELINE.json configs" in d["config"]["workload"]
