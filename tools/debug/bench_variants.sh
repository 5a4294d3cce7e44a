for v in "" "--level-streams 0" "--clip-chunks 2" "--clip-chunks 3"; do
  timeout 300 python bench.py $v --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ks=d['roofline']['kernels']
print('$v', 'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel-sum ms',round(sum(k['ms_per_step'] for k in ks.values()),3), 'hbm_view', d['roofline'].get('hbm_view'))"
done
