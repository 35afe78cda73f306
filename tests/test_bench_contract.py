"""bench.py contract on CPU: the reference arm (CPU oracle port; the one place outside tests/ and smoke() that may execute
oracle/) prints exactly ONE JSON line on stdout with the keys the driver reads, whatever the libraries write elsewhere."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--cpu-centers", "1",
           "--prec", "1e-4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "poisson_apply_output_nodes_per_s" and d["unit"] == "nodes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("poisson_apply_k7_prec")
    # the arm states the workload it REALLY times: the bounded CPU sample, and which full workload it is a sample of
    assert d["config"]["centers"] == 1 and d["config"]["workload"].endswith("_cpu_sample")
    assert d["config"]["sample_of"].endswith("_gauss1000")


def test_reference_arm_config_c4_helmholtz_one_orbital():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c4", "--steps", "1", "--warmup", "3",
           "--prec", "1e-3", "--order", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert d["metric"] == "helmholtz_apply_output_nodes_per_s" and d["config"]["config"] == "c4" and d["config"]["centers"] == 1
    assert d["config"]["operator"].startswith("HelmholtzOperator")
