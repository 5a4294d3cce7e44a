#!/bin/bash
# compute-sanitizer memcheck over the WHOLE GPU test suite (every kernel path the tests reach)
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|at 0x|by thread|Address|axvs::" | head -60 > gpurun_out/r02_sanitizer_memcheck_full.txt
cat gpurun_out/r02_sanitizer_memcheck_full.txt
