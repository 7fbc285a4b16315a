python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_hl2.json 2> gpurun_out/r2_hl2.err
tail -c 800 gpurun_out/r2_hl2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_hl2.json"))
print(round(d["value"]), d["serial"]["ms_per_step"], d["kernel_ms"])
for k in ("e2e", "e2e_copy_all", "e2e_full"):
    print(k, round(d[k]["value"]), d[k]["ms_per_step"], d[k]["h2d_bytes_per_step"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_greedy|enc_pass1" -s 6 -c 2 -f -o gpurun_out/r2_win_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/r2_win_ncu.log 2>&1
tail -2 gpurun_out/r2_win_ncu.log
