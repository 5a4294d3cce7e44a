#!/bin/bash
# ncu launch list (names + durations, in order) of ONE WithinClipTrackingModule.forward_features at 16 clips
mkdir -p gpurun_out
cat > /tmp/one_forward.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
exec(open("tools/debug/bench_module.py").read().split("for _ in range(2)")[0])
with torch.no_grad():
    for _ in range(2): m.forward_features(feats)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m.forward_features(feats)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02g_module_launches.csv python /tmp/one_forward.py 16 > gpurun_out/r02g_module_ncu.log 2>&1
echo rc=$?; wc -l gpurun_out/r02g_module_launches.csv
