#!/bin/bash
# round-2 GPU job 19 (2 GPUs): final kernels through the C-ABI NCCL all-gather path of bench.py, plus the N=1 line on the same box
O=gpurun_out/r02s
mkdir -p $O
T="timeout -k 10"
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
$T 300 python bench.py --cpu-images 0 --no-secondary > $O/bench_n1.json 2> $O/bench_n1.err
NCCL_DEBUG=WARN $T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "n2 rc=$?"; tail -c 800 $O/bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02s/bench_n1.json','gpurun_out/r02s/bench_n2.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f.split('/')[-1], 'n', d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d.get('rank_ms_per_step'), d.get('gathered_ok'), d['config']['collective'][:60])
    except Exception as e: print(f, 'ERR', e)
PY
