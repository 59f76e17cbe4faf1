#!/bin/bash
O=gpurun_out/final1; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final1/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('e2e',d['e2e']['value']); print('run',d['run'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','traffic','extend_share_of_step')}); print('cpu',d['cpu_baseline']['value'], d['cpu_baseline'].get('at_reference_thread_count'))
for k,v in d['workloads'].items(): print(k, round(v['value'],1), round(v['ms'],1))
print('c5', d['c5']['value'], d['c5']['ms'])
r=json.loads(open('gpurun_out/final1/bench_reference.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
KT_POOL=125000000 KT_SPP=192 timeout 300 python tools/kernel_times.py final_scene | tail -1
