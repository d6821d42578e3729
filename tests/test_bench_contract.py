"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_json():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                                   "--size", "48", "40", "32"], cwd=ROOT, timeout=600).decode().strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["metric"] == "demons_voxel_iterations_per_s" and line["unit"] == "Mvoxel*it/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=env, capture_output=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == b""
