mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 150 -x 2>&1 | tail -3
timeout 200 python tools/phase_timing.py 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(round(d['value']), d['ms_per_step'], {k:round(v*1e3,1) for k,v in d['kernel_ms'].items()}, d['roofline']['frac'])
"
