#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16" -s 2 -c 1 -f -o gpurun_out/r02g_gemm_out python tools/debug/bench_msda.py 32 > gpurun_out/r02g_ncu_gemm.log 2>&1
echo rc=$?; ls -la gpurun_out/r02g_gemm_out.ncu-rep
