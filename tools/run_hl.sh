python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_hl.json 2> gpurun_out/r2_hl.err
tail -c 1500 gpurun_out/r2_hl.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_hl.json"))
print(round(d["value"]), d["serial"]["ms_per_step"], d["kernel_ms"])
for k in ("e2e", "e2e_copy_all", "e2e_full"):
    print(k, round(d[k]["value"]), d[k]["ms_per_step"], d[k]["h2d_bytes_per_step"])
PY
