"""bench.py's JSON contract, checked on CPU through the reference arm (the oracle on the host cores) and statically for the
GPU arm's line (no GPU here)."""
import json
import os
import re
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "cpu_baseline"]


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample-mib", "2"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED + ["impl"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("single 10 GiB-per-GPU synthetic slot (163840 blocks")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_line_has_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("        line = {\n            \"metric\""):]
    for k in REQUIRED + ["gpu_launches", "clocks", "roofline", "perms_per_s"]:
        assert f'"{k}"' in body, k
    roof = src[src.index("    roofline = {"):src.index("    traffic_file")]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert f'"{k}"' in roof, k
    assert re.search(r'"sm_mhz".*"sm_max_mhz".*"reasons"', src, re.S)


def test_work_model_matches_the_survey_appendix():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.total_perms(64) == 71679                      # config 1, one slot
    assert bench.total_perms(16384) == 18350079                # 1 GiB
    assert bench.total_perms(163840) == 183500801              # 10 GiB (odd nodes at 5 -> 3 -> 2 -> 1)
    assert bench.total_perms(1638400) == 1835008002            # 100 GiB
    assert bench.total_perms(2097152) == 2348810239            # 128 GiB
