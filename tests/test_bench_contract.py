"""bench.py contract checks that run without a GPU (the reference arm is CPU-only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout  # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["batch_per_gpu"] == 1024 and d["dtype"] == "f64"


def test_flop_and_byte_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes_per_solve(17, 6, 20) == 8152
    f = bench.algorithmic_flops_per_solve(17, 6, 20, 10)
    assert 5.0e6 < f < 5.6e6                                           # "about 5.3 Mflop/solve" at 10 factorisations
