#!/bin/bash
# 2-GPU validation of the NCCL paths (config 1 weak scaling, config 3 forward_sharded)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_n2_cfg1.json 2> gpurun_out/r2m_n2_cfg1.err; echo "cfg1 rc=$?"; tail -2 gpurun_out/r2m_n2_cfg1.err
timeout 600 $RUN bench.py --gpus 2 --config 3 --steps 5 --warmup 3 > gpurun_out/r2m_n2_cfg3.json 2> gpurun_out/r2m_n2_cfg3.err; echo "cfg3 rc=$?"; tail -2 gpurun_out/r2m_n2_cfg3.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2m_n1_cfg3.json 2> gpurun_out/r2m_n1_cfg3.err; echo "cfg3 n1 rc=$?"
python - <<PY
import json
for n in ("n2_cfg1","n2_cfg3","n1_cfg3"):
    try:
        d=json.load(open(f"gpurun_out/r2m_{n}.json"))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("roofline"))
PY
