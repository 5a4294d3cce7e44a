# A/B of library variants inside one gpurun call: LIBS="libaxvs libaxvs_x" KF="traj|qkv" bash tools/debug/ab_libs.sh
for i in 1 2; do
  for lib in ${LIBS:-libaxvs}; do
    echo "== $lib ($i)"; AXVS_LIB=axial_vs_b200/$lib.so timeout 200 python tools/debug/level_times.py 42 2>&1 | grep -E "level|${KF:-kernel}"
  done
done
for lib in ${LIBS:-libaxvs}; do
  echo "== bench $lib"; AXVS_LIB=axial_vs_b200/$lib.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
