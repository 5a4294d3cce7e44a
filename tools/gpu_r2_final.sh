#!/bin/bash
# final round-2 state: GPU tests, smoke, default bench line, whole-module breakdown, per-kernel layer times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02g_tests.txt; cat gpurun_out/r02g_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -c 600 gpurun_out/r02g_bench.err; cut -c1-300 gpurun_out/r02g_bench.json
timeout 300 python tools/debug/bench_module.py 16 > gpurun_out/r02g_module16.txt 2>&1; cat gpurun_out/r02g_module16.txt
timeout 200 python tools/debug/level_times.py 42 > gpurun_out/r02g_levels.txt 2>&1; cat gpurun_out/r02g_levels.txt
