"""bench.py's reference arm (`--impl reference`): the CPU oracle timed on the host cores, no GPU and no product library involved.
Runs here (CPU container) in a few seconds; the b200 arm needs a device and is exercised by the driver and `tools/r02_refresh.sh`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUNNER = r"""
import runpy, sys
sys.argv = ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"]
try:
    runpy.run_path("bench.py", run_name="__main__")
except SystemExit as e:
    assert not e.code, e.code
assert "vkhrt_b200" not in sys.modules and "vkhrt_b200.api" not in sys.modules, "the reference arm must not load the product"
maps = open("/proc/self/maps").read()
assert "libvkhrt_b200" not in maps, "libvkhrt_b200.so is mapped in the reference arm"
import os
if os.environ.get("RANK", "0") == "0":
    assert "oracle" in maps          # the checker did the work
"""


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, "-c", RUNNER], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line_and_never_loads_the_product():
    lines = run()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s primary-ray hair hits" and d["unit"] == "Mrays/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c2:")                       # the same config as the b200 arm's default line
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "pixels" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_env_only_rank0_works():
    """N > 1: rank 0 alone runs and prints the line, the other ranks exit 0 without work"""
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
    lines = run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference"
