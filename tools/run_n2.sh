python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2b_n2.json 2> gpurun_out/r2b_n2.err
tail -c 400 gpurun_out/r2b_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_n2.json").read().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["gather"][:40], d["run"]["gather_fallback"], round(d["e2e"]["value"]), d["gpu_launches"])
PY
