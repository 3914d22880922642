#!/bin/bash
# round-2 GPU job 5: non-default head configs + arena reuse tests, and a SAME-BOX A/B: round-1 build (_ab_old worktree) vs
# current vs current with CTA pairs down to N = 128, alternating, two rounds.
O=gpurun_out/r02e
mkdir -p $O
T="timeout -k 10"
$T 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -3 $O/canary.log
if [ $rc -ne 0 ]; then echo "CANARY FAILED -- aborting"; exit 1; fi
$T 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest.log
for round in 1 2; do
  (cd _ab_old && $T 200 python bench.py --cpu-images 0 > ../$O/ab_old_$round.json 2> ../$O/ab_old_$round.err)
  $T 200 python bench.py --cpu-images 0 --no-secondary > $O/ab_new_$round.json 2> $O/ab_new_$round.err
  DD3D_CONV_CTA2_MINN=128 $T 200 python bench.py --cpu-images 0 --no-secondary > $O/ab_new128_$round.json 2> $O/ab_new128_$round.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e/ab_*.json')):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['kernels_ms_per_step'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
DD3D_CONV_CTA2_MINN=128 $T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99_minn128.txt 2>&1
$T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
$T 200 python bench.py --workload dla34 --sweep 8,16,32,64 --cpu-images 0 > $O/sweep_dla34_bf16.jsonl 2> $O/sweep_dla34_bf16.err
$T 200 python bench.py --workload dla34 --dtype fp16 --sweep 8,16,32,64 --cpu-images 0 > $O/sweep_dla34_fp16.jsonl 2> $O/sweep_dla34_fp16.err
python - <<'PY'
import json
for f in ('gpurun_out/r02e/sweep_dla34_bf16.jsonl','gpurun_out/r02e/sweep_dla34_fp16.jsonl'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['global_batch'], round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
PY
