"""bench.py's reference arm runs without a GPU: one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import util


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--lanes", "2", "--frames", "2",
                          "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, check=True, timeout=300).stdout.decode()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["config"]["workload"] == "h264_baseline_cavlc_1080p_ip" and j["dtype"] == "u8" and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "data"):
        assert k in j


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], stdout=subprocess.PIPE, env=env, timeout=60)
    assert r.returncode == 0 and not r.stdout.strip()
