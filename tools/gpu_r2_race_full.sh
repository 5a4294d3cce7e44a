#!/bin/bash
# compute-sanitizer racecheck over the hot-path tests (trajectory attention, layers, encoder, whole module, pair kernels)
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory or axial or encoder or pair or wc_ or level_plumbing" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|Race reported|hazard|axvs::" | head -40 > gpurun_out/r02_sanitizer_racecheck_hotpath.txt
cat gpurun_out/r02_sanitizer_racecheck_hotpath.txt
