#!/bin/bash
# compute-sanitizer on the kernels added at the end of round 2 (MSDeformAttn front / tail / sampler, strided projections, per-frame fp32 attention)
mkdir -p gpurun_out
K="msda_front or msda_layer_golden or projections_level_slices or split_precision or convnext"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|at 0x" | head -20 > gpurun_out/r02_sanitizer_memcheck_msda_kernels.txt
cat gpurun_out/r02_sanitizer_memcheck_msda_kernels.txt
K2="msda_front or projections_level_slices"
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K2" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|Race reported|hazard" | head -20 > gpurun_out/r02_sanitizer_racecheck_msda_kernels.txt
cat gpurun_out/r02_sanitizer_racecheck_msda_kernels.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_front|msda_tail|msda_sample_lp" -s 6 -c 3 -f -o gpurun_out/r02h_msda_layer python tools/debug/bench_msda.py 32 > gpurun_out/r02h_ncu_msda.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02h_msda_layer.ncu-rep
