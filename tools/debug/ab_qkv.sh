for i in 1 2; do
  for lib in libaxvs_new2 libaxvs_new3; do
    echo "== $lib ($i)"; AXVS_LIB=axial_vs_b200/$lib.so timeout 200 python tools/debug/level_times.py 42 2>&1 | grep -E "level|qkv|total"
  done
done
AXVS_LIB=axial_vs_b200/libaxvs_new3.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or axial or encoder or fusion or pair" 2>&1 | tail -5
AXVS_LIB=axial_vs_b200/libaxvs_prof.so python tools/wait_profile.py qkvd 42 41 --pair-ctas 2>&1 | tail -16
