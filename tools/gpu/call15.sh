#!/bin/bash
O=gpurun_out/c15; mkdir -p $O
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt; tail -4 $O/pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench exit $?"
tail -3 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c15/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',d.get('e2e')); print('c5',d.get('c5'))
PY
timeout 300 raytracer-2022_b200/rt_render --scene final_scene --width 800 --height 800 --spp 64 --devices 2 --out $O/final2.ppm | tail -8
