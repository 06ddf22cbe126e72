"""bench.py contract pieces that need no GPU: the reference arm's JSON line and the helpers around it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU path on the host cores) on the small config: one JSON line with the
    keys the driver reads, e2e equal to the line's own value and zero copy bytes"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "fast"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
                "cpu_baseline", "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    # ms_per_step is the measured time of one sampled step (what the driver multiplies by steps), not the scaled one
    assert line["ms_per_step"] > 0 and "ms_per_step_is" in line


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n = mod.ncu_summaries()
    t = n["corr_dram_bytes"]
    assert t is not None and 1e8 < t < mod.BYTES_PER_EDGE_FP16 * 47712       # DRAM bytes stay below the algorithmic bytes
    assert n["corr_l2_bytes"] > t and n["corr_source"].startswith("profiles/")
    hbm, tf, src = mod.peaks()
    assert 1000 < hbm < 9000 and 500 < tf < 2500
