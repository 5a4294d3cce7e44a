for c in ${CLIPS:-42 45 63 84}; do
  timeout 300 python bench.py --clips $c --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('clips',$c,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['step_frac_of_burst_peak'])"
done
