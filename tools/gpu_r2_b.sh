#!/bin/bash
# round-2 call B: full GPU suite, then bench A/B of the two attention cores inside one box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.txt
tail -5 gpurun_out/r2b_tests.txt
AXVS_ATTN_CORE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_core0.json 2> gpurun_out/r2b_bench_core0.err
AXVS_ATTN_CORE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_core1.json 2> gpurun_out/r2b_bench_core1.err


python - <<PY
import json
for c in (0,1):
    d=json.load(open(f"gpurun_out/r2b_bench_core{c}.json"))
    print(c, d["value"], d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["roofline"]["kernels"].items()})
PY
