#!/bin/bash
# 8-GPU measurements: e2e scaling of config 1 (default / bf16 host boundary / summary read-back), PCIe ceilings at N = 8, config 3
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n_topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $RUN bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/r2n_n8_default.err | grep '^{' > gpurun_out/r2n_n8_default.json; echo "default rc=$?"
timeout 400 $RUN bench.py --gpus 8 --steps 20 --warmup 3 --host-dtype bf16 2> gpurun_out/r2n_n8_bf16.err | grep '^{' > gpurun_out/r2n_n8_bf16.json; echo "bf16 rc=$?"
timeout 400 $RUN bench.py --gpus 8 --steps 20 --warmup 3 --host-dtype bf16 --d2h summary 2> gpurun_out/r2n_n8_bf16_summary.err | grep '^{' > gpurun_out/r2n_n8_bf16_summary.json; echo "bf16+summary rc=$?"
timeout 400 $RUN bench.py --gpus 8 --config 3 --steps 5 --warmup 3 2> gpurun_out/r2n_n8_cfg3.err | grep '^{' > gpurun_out/r2n_n8_cfg3.json; echo "cfg3 rc=$?"
python - <<PY
import json
for n in ("n8_default","n8_bf16","n8_bf16_summary","n8_cfg3"):
    try:
        d=json.load(open(f"gpurun_out/r2n_{n}.json"))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("roofline"))
PY
