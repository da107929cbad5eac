"""bench.py's JSON contract on the CPU: the reference arm prints one line whose `config` object is the one the B200 arm
prints for the same workload (the driver compares them), with the keys the harness reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_shared_config():
    sys.path.insert(0, ROOT)
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "1.5",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "snapshots/s" and line["higher_is_better"] is True
    assert line["config"] == bench.workload_config(*bench.GRIDS["1.5"])           # identical objects in both arms
    assert line["config"] == json.loads(json.dumps(bench.workload_config(240, 121)))
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "cpu_baseline",
                "e2e"):
        assert key in line
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert bench.cpu_split(1440, 16) == 8 and bench.cpu_split(240, 16) == 8
