#!/bin/bash
mkdir -p gpurun_out
for k in traj ffn qkvd; do AXVS_LIB=axial_vs_b200/libaxvs_prof.so timeout 300 python tools/wait_profile.py $k 42; done > gpurun_out/r2_wp_base.txt 2>&1
cat gpurun_out/r2_wp_base.txt
