timeout 600 python -m pytest tests -m gpu -x -q -k "msda or within_clip" 2>&1 | tail -3
for lib in libaxvs libaxvs_pb4 libaxvs_pb1; do echo "== $lib"; AXVS_LIB=axial_vs_b200/$lib.so timeout 200 python tools/debug/bench_msda.py 32 2>&1 | grep -E "msda layer|msda_sample"; done
