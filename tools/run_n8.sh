python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 5 > gpurun_out/r2b_n8.json 2> gpurun_out/r2b_n8.err
tail -c 600 gpurun_out/r2b_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_n8.json").read().splitlines()[-1])      # (NCCL prints its version on stdout first)
print(round(d["value"]), d["ms_per_step"], d["serial"]["ms_per_step"], d["run"].get("host_cores_per_rank"), d["run"].get("numa_node_rank0"))
for k in ("e2e", "e2e_copy_all", "e2e_full"):
    print(k, round(d[k]["value"]), d[k]["ms_per_step"], d[k].get("h2d_gbs_per_gpu"))
print(d["run"]["gather_check"])
PY
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|model name" | head
