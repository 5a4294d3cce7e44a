# A/B of the frame-major temporal row order (pair-mode bit 16) inside one gpurun call: per-kernel times of one axial layer, then tests
L=${L:-axial_vs_b200/libaxvs_new5.so}
for i in 1 2; do
  for pm in 14 30; do
    echo "== AXVS_PAIR=$pm ($i)"; AXVS_PAIR=$pm AXVS_LIB=$L timeout 200 python tools/debug/level_times.py 42 2>&1 | grep -E "level|kernel|total"
  done
done
AXVS_LIB=$L timeout 800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or axial or encoder or fusion or pair or tube or wc_ or module" 2>&1 | tail -5
