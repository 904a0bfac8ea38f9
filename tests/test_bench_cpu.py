"""bench.py on a box without a GPU: the reference arm (CPU restatement on a bounded sample) prints one JSON line with the
contract's keys for every configuration; the B200 arm fails loudly instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config,rows", [("c2", 4), ("c2w", 4), ("c4", 1), ("c5", 1)])
def test_reference_arm_prints_the_contract_line(config, rows):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config, "--cpu-rows", str(rows),
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Mpx/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "reference_sample" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--cpu-rows", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_source_hash_guards_the_stored_traffic(tmp_path):
    """roofline.traffic is only reported while profiles/traffic_<config>.json belongs to the present kernel sources"""
    sys.path.insert(0, ROOT)
    import bench
    h = bench.source_hash()
    assert len(h) == 16 and h == bench.source_hash()
    t, info = bench.stored_traffic("c2", 4096, 1)
    path = os.path.join(ROOT, "profiles", "traffic_c2.json")
    if os.path.exists(path):
        stored = json.load(open(path))
        assert (t is not None) == (stored.get("source_hash") == h and stored.get("rows") == 4096)
    assert bench.stored_traffic("c2", 123, 1) == (None, None)
    assert bench.stored_traffic("nosuchconfig", 4096, 1) == (None, None)
