#!/bin/bash
# round-2 evidence: launch list of the bench step + one `--set full` capture of a res4 layer (7 kernels) of the final default path
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --level-streams 0 --graph 0 > gpurun_out/r02h_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02h_launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"qkv_pair|spatial_attn_tc|traj_pair|ffn_n256_pair" -s 77 -c 7 -f -o gpurun_out/r02h_layer_res4 python tools/debug/level_times.py 42 > gpurun_out/r02h_ncu_layer.log 2>&1
echo "set full rc=$?"; ls -la gpurun_out/r02h_layer_res4.ncu-rep
