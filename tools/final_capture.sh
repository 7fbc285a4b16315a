#!/bin/bash
# Round evidence run (one GPU): parity tests, smoke, bench lines of both arms, ncu launch list, one ncu --set full capture.
# Usage on the GPU box:  bash tools/final_capture.sh [tag]      (writes everything under gpurun_out/<tag>_*)
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log
tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("${TAG}_bench_reference_arm", "${TAG}_bench"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              (d.get("cpu_baseline") or {}).get("value"), (d.get("serial") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"enc_pass|nms_greedy" \
    -s 6 -c 5 -f -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_*
