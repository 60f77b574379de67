"""CPU-side checks of the measurement contract (bench.py): the reference arm prints one JSON line with the agreed keys, the
product arm refuses to run without a CUDA device (no CPU fallback), the clock sampler copes with a slow nvidia-smi, and the
profile summariser reads the committed launch list."""
import json
import os
import stat
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--m", "24", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ca_lanczos_s_step_blocks_per_sec" and d["unit"] == "blocks/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("laplace3d_24^3_7pt_s8_newton_cholqr2")


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--m", "24", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_product_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-clocks"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stdout + out.stderr)


def test_clock_sampler_waits_for_a_slow_nvidia_smi(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.5\nwhile true; do echo '1965, 1965, 400.0, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cs = bench.ClockSampler(0, True)
    cs.mark()
    assert cs.rows, "mark() must wait until the sampler delivers"
    cs.mark(0.0)
    time.sleep(0.08)
    c = cs.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] >= 1


def test_profile_summary_reads_the_committed_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_summary.py"),
                          os.path.join(ROOT, "profiles", "r1_final_launches_bench_steps3.csv")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "k_spmv_selld" in out.stdout and "k_tile" in out.stdout and "MPK share under ncu" in out.stdout
