"""bench.py's reference arm (`--impl reference`) on the CPU box: one JSON line with the contract's keys, whether stdout
is a pipe or a FILE (the arm runs the reference's machine code in a child with RLIMIT_FSIZE = 0; the parent relays the
line, otherwise a redirected print would fail with EFBIG and the driver would record no reference number)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(line):
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "ROIs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("3D ROIAlign") and "cfg2" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "ROIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_json_line_to_a_file(tmp_path):
    out = tmp_path / "ref.json"
    with open(out, "w") as f:
        rc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                            stdout=f, stderr=subprocess.PIPE, text=True, timeout=600)
    assert rc.returncode == 0, rc.stderr[-1000:]
    lines = [ln for ln in out.read_text().splitlines() if ln.strip()]
    assert len(lines) == 1
    _check(lines[0])


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert rc.returncode == 0 and rc.stdout.strip() == ""
