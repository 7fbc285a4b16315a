"""The bench.py contract (one JSON line per run, the keys the driver reads) checked without a GPU: the reference arm runs on
the CPU, and the committed final line of the CUDA arm (profiles/) is checked for the same keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline"]


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    line = json.loads(lines[-1])
    for k in COMMON:
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "images/s"
    assert line["config"]["workload"].startswith("S3FD 640x640") and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1


def test_committed_cuda_line_has_the_contract_keys():
    line = json.load(open(os.path.join(ROOT, "profiles", "r01_final_bench.json")))
    for k in COMMON + ["clocks", "gpu_launches", "roofline"]:
        assert k in line, k
    assert line["gpu_launches"] == 7 * line["steps"] and line["dtype"] == "f32" and line["data"] == "synthetic"
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert line["clocks"]["reasons"] == [] or "sw_power_cap" in "".join(line["clocks"]["reasons"])
