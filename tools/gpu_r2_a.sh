#!/bin/bash
# round-2 bring-up call A: SW64 / MN-major operand forms, then the tcgen05 attention core tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 ./tools/microbench/umma_sw64 > gpurun_out/r2a_sw64.txt 2>&1; echo "sw64 rc=$?" >> gpurun_out/r2a_sw64.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention_core_tcgen05" > gpurun_out/r2a_tc.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_tc.txt
tail -30 gpurun_out/r2a_sw64.txt; tail -30 gpurun_out/r2a_tc.txt
