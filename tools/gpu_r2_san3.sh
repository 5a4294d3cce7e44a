#!/bin/bash
# compute-sanitizer on the kernels changed at the end of round 2: qkv_pair_kernel (bulk shared->global stores), spatial_attn_tc_kernel and
# traj_pair_kernel (frame-major rows, split o_ready barriers)
mkdir -p gpurun_out
K="pair_kernels_bit_identical or test_trajectory_attention_oracle or axial_layer"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|at 0x" | head -20 > gpurun_out/r02_sanitizer_memcheck_end.txt
cat gpurun_out/r02_sanitizer_memcheck_end.txt
K2="pair_kernels_bit_identical"
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K2" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|Race reported|hazard" | head -20 > gpurun_out/r02_sanitizer_racecheck_end.txt
cat gpurun_out/r02_sanitizer_racecheck_end.txt
