#!/bin/bash
O=gpurun_out/c11; mkdir -p $O
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?"
tail -5 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e',d.get('e2e')); print('run',d.get('run')); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms','rays_per_launch','extend_share_of_step')})
print('cpu',d.get('cpu_baseline'))
for k,v in (d.get('workloads') or {}).items(): print(k, round(v['value'],1),'Mrays/s', round(v['ms'],1),'ms', v['iterations'],'it', v['kernel_ms'])
print('c5',d.get('c5'))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
