#!/bin/bash
# A/B inside one gpurun call: per-kernel times of one axial layer (42 clips) for each library variant named in $LIBS (suffixes of
# axial_vs_b200/libaxvs_*.so; "new" = libaxvs.so), twice, then (unless $1 = notest) the trajectory parity tests on libaxvs.so
mkdir -p gpurun_out
KF=${KF:-traj_ts}
for i in 1 2; do
  for lib in ${LIBS:-base new}; do
    f=axial_vs_b200/libaxvs_$lib.so; [ $lib = new ] && f=axial_vs_b200/libaxvs.so
    echo "== $lib ($i)"; AXVS_LIB=$f timeout 200 python tools/debug/level_times.py 42 2>&1 | grep -E "level|$KF"
  done
done
if [ "$1" != "notest" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${1:-trajectory or axial or encoder or fusion}" 2>&1 | tail -5; fi
