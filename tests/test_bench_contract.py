"""CPU-only: the parts of bench.py's contract that do not need a GPU -- the reference arm prints
exactly one JSON line with the agreed keys (rank 0) or nothing (other ranks), nothing else
reaches stdout, and the helper parsers behave."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ref(rank):
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LASER_B200_REF_BUDGET_S="1.5")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                           "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_rank0_prints_one_json_line():
    out = run_ref(0)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    out = run_ref(1)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_parsing_and_peaks():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = ["0, 1965, 1965, 120.5, 0x0, Not Active, Not Active, Not Active, Not Active",
               "0, 1400, 1965, 990.1, 0x4, Not Active, Not Active, Not Active, Active",
               "0, 1380, 1965, 1001.0, 0x4, Not Active, Not Active, Not Active, Active"]
    c = s.stop()
    assert c["sm_max_mhz"] == 1965 and c["reasons"] == ["sw_power_cap"] and 1380 <= c["sm_mhz"] <= 1400
    p = bench.load_peaks()
    assert p["bf16"] > 0 and p["source"] in ("measured", "fallback")
