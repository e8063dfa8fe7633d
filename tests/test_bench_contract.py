"""CPU tier: the bench contract's reference arm runs here (it is the CPU path) and prints one JSON line with the agreed keys;
the GPU arm is exercised with the same key check on the box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
             "cpu_baseline", "gpu_launches"}


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--terrain-n", "96", "--steps", "2", "--warmup", "1", "--ref-log2-rays", "14"],
                         capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["metric"] == "incoherent_mrays_per_s" and line["unit"] == "Mrays/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_contract_keys(gpu):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--terrain-n", "300", "--log2-rays", "20", "--steps", "3", "--warmup", "3", "--cpu-log2-rays", "16",
                          "--spp", "2", "--bounces", "3"], capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert BASE_KEYS | {"roofline", "clocks", "hit_id_mismatches", "spp_per_s"} <= set(line)
    assert line["hit_id_mismatches"] == 0 and line["hit_t_max_ulp"] == 0 and line["brute_force_mismatches"] == 0
    assert line["gpu_launches"] == 3 and line["value"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == (1 << 20) * 32
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["frac"] < 2 and line["cpu_baseline"]["value"] > 0
