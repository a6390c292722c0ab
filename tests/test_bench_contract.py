"""CPU suite, part 4: the bench.py contract.  The reference arm runs here (it is the CPU oracle); the GPU arm's
JSON line is checked on the committed record of the last round-1 run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def test_reference_arm_prints_one_contract_line(built_libs):
    env = dict(os.environ, PFLARE_BENCH_CACHE="/nonexistent")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "96", "--steps", "2",
                          "--warmup", "1", "--no-cache"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "DOF/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_recorded_gpu_line_has_roofline_and_clocks():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_4096_n1_final.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * d["config"]["rows"] == d["e2e"]["d2h_bytes_per_step"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] is not None
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
