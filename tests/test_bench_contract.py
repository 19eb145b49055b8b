"""bench.py's contract on the CPU: the reference arm (`--impl reference`) runs without a GPU, prints ONE JSON line with the keys
the driver reads, names the same config as the B200 arm, and says that each step is a bounded sample; the argument defaults
the round-end runs rely on (N = 1, K / W that finish in minutes, overlap decided by the world size)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    j = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--records", "4000")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["impl"] == "reference" and j["gpu_launches"] == 0 and j["vs_baseline"] is None
    assert j["metric"] == "canonicalize+uniq records/sec" and j["unit"] == "records/s" and j["higher_is_better"] is True
    assert j["dtype"] == "u8" and j["data"] == "synthetic" and j["n_gpus"] == 1
    assert "config 2" in j["config"]["workload"] and "REDUCED" in j["config"]["workload"]      # --records: not the named config
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["value"] > 0 and j["ms_per_step"] > 0


def test_argument_defaults():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'default="c2"' in src                                   # the default workload is the config BASELINE.json's metric is quoted on
    assert bench.WORKLOADS["c2"]["records"] == 10_000_000 and bench.WORKLOADS["c5"]["records"] == 12_500_000
    assert bench.WORKLOADS["c1"]["records"] == 1_000_000 and bench.WORKLOADS["c3"]["records"] == 5_000_000
    assert bench.WORKLOADS["c4"]["records"] == 200_000
