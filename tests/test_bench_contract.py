"""CPU suite, part 4: the bench.py contract.  The reference arm runs here (it is the CPU oracle); the GPU arm's
JSON line is checked on the committed record of the last round-2 run."""
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
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_4096_n1_final.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * d["config"]["rows"] == d["e2e"]["d2h_bytes_per_step"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] is not None
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    assert d["parity"]["rel_l2"] <= d["parity"]["tol"] == 1e-12 and d["parity"]["n"] == d["config"]["rows"]
    assert r["frac"] >= 0.60


def test_ncu_traffic_record_is_consistent():
    """profiles/r02_ncu_traffic.json (ncu --set full, dram__bytes per launch) is what bench.py reports as roofline.traffic when its
    stamp matches the kernel source: measured DRAM bytes must be close to the algorithmic bytes of the same launches."""
    t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    assert len(t["launches"]) >= 10 and 0.8 <= t["dram_over_algorithmic"] <= 1.2
    assert abs(t["mean_dram_bytes_per_launch"] - sum(x["dram_bytes"] for x in t["launches"]) / len(t["launches"])) < 1.0
    assert all(x["kernel"].startswith(("spmv_wc_kernel", "spmv_sv_kernel")) for x in t["launches"])
