#!/bin/bash
# tc attention v3 (unit-major layout): parity subset, wait profile, bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention_core or direct_qkv or axial_layer or encoder or fusion or trajectory_attention or shared_pos or cross or maps" > gpurun_out/r2h_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.txt
tail -4 gpurun_out/r2h_tests.txt
AXVS_LIB=axial_vs_b200/libaxvs_prof.so timeout 300 python tools/wait_profile.py attn 42 41 > gpurun_out/r2h_wait_attn41.txt 2>&1
AXVS_LIB=axial_vs_b200/libaxvs_prof.so timeout 300 python tools/wait_profile.py attn 42 21 > gpurun_out/r2h_wait_attn21.txt 2>&1
cat gpurun_out/r2h_wait_attn41.txt gpurun_out/r2h_wait_attn21.txt
AXVS_ATTN_CORE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_core0.json 2> gpurun_out/r2h_bench_core0.err
AXVS_ATTN_CORE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_core1.json 2> gpurun_out/r2h_bench_core1.err
python - <<PY
import json
for c in (0,1):
    d=json.load(open(f"gpurun_out/r2h_bench_core{c}.json"))
    print(c, d["value"], d["ms_per_step"], d["e2e"]["value"], {k:v["ms_per_step"] for k,v in d["roofline"]["kernels"].items()})
PY
