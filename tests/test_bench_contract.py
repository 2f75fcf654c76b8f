"""bench.py's reference arm runs on the CPU box: check the JSON line carries the contract's keys."""
import json
import os
import subprocess
import sys

from tests.conftest import REPO

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _run(args, env=None):
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "msda_fwd_bwd_achieved_hbm_gbps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the unmodified reference function when it is staged (baseline/_ref) or present (/root/reference), else the port
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import stage_reference
    assert d["cpu_baseline"]["kind"] == ("reference" if stage_reference.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "N=20" in d["cpu_baseline"]["sample"]
    assert "N=20" in d["config"]["workload"]                       # the arm runs the configuration it states
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0


def test_reference_arm_non_zero_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env) == []


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
