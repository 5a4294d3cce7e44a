#!/bin/bash
# ncu --set full of the res3 input projection (generic GEMM with the NCHW fp32 producers) at 84 images
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16" -s 52 -c 1 -f -o gpurun_out/r02i_inproj python tools/debug/bench_proj.py > gpurun_out/r02i_ncu_inproj.log 2>&1
echo rc=$?; ls -la gpurun_out/r02i_inproj.ncu-rep; tail -5 gpurun_out/r02i_ncu_inproj.log
