"""The reference arm of bench.py runs on the CPU alone (it times the oracle, the C++ restatement of the reference), so its
JSON contract can be checked without a GPU.  The GPU arm's line carries the same keys plus roofline / e2e details; its
static shape is checked here from the source."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline")


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample", "1500"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in REQUIRED:
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "plonk_by_hand_prove_plus_verify_throughput" and j["unit"] == "proof+verify/s"
    assert j["steps"] == 2 and j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_line_has_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("    line = {\n        \"metric\": METRIC"):]
    for k in REQUIRED + ("clocks", "gpu_launches", "roofline"):
        assert re.search(r'"%s":' % k, body), k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert re.search(r'"%s":' % k, body[body.index('"roofline"'):]), k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert re.search(r'"%s":' % k, src[src.index("cpu = {"):]), k
