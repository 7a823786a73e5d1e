"""bench.py's reference arm (the only arm that runs without a GPU) prints ONE JSON line with the keys the driver reads:
checked for the headline workload and for the vanishing-point workload (whose CPU side runs the reference's own code in
worker processes where oracle/_ref exists, the oracle port otherwise)."""
import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_ref(*extra):
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *extra],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("extra", [(), ("--workload", "V1", "--unique", "8")])
def test_reference_arm_line(extra):
    d = run_ref(*extra)
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["unit"] == "frames/s" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not extra:  # the headline workload: the very `config` object the B200 arm prints (bench.lsd_config)
        import argparse
        import bench
        a = argparse.Namespace(batch=4736, unique=96, slots=2, max_lines=0, upload_ahead=True, e2e_together=True)
        assert d["config"] == bench.lsd_config(a, bench.WORKLOADS["C2"], 1)
        assert d["metric"] == bench.METRIC and d["sample_frames_per_step"] >= 2


def test_other_ranks_do_no_work():
    """Under torchrun only rank 0 runs and prints the reference arm; the others exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
