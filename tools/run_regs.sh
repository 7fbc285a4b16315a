for v in base r48 r40 base r48; do
  if [ $v = base ]; then unset DAN_B200_LIB; else export DAN_B200_LIB=build/variants/libdan_$v.so; fi
  python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/regs_$v.json 2> gpurun_out/regs_$v.err || tail -3 gpurun_out/regs_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/regs_$v.json"))
print("$v", round(d["value"]), d["serial"]["ms_per_step"], d["kernel_ms"], d["run"]["native_so_loaded"])
PY
done
