#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2n_n1_default.err | grep '^{' > gpurun_out/r2n_n1_default.json
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --host-dtype bf16 2> gpurun_out/r2n_n1_bf16.err | grep '^{' > gpurun_out/r2n_n1_bf16.json
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --host-dtype bf16 --d2h summary 2> gpurun_out/r2n_n1_bf16_summary.err | grep '^{' > gpurun_out/r2n_n1_bf16_summary.json
timeout 400 python bench.py --config 3 --steps 5 --warmup 3 2> gpurun_out/r2n_n1_cfg3.err | grep '^{' > gpurun_out/r2n_n1_cfg3.json
python - <<PY
import json
for n in ("n1_default","n1_bf16","n1_bf16_summary","n1_cfg3"):
    try:
        d=json.load(open(f"gpurun_out/r2n_{n}.json"))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("roofline"))
PY
tail -3 gpurun_out/r2n_n1_*.err
