#!/bin/bash
# round-2 GPU job 15: A/B of CTA pairs for the N = 64 layers (operand-fetch model predicts 33 % -> 40 % tensor pipe)
O=gpurun_out/r02o
mkdir -p $O
T="timeout -k 10"
for round in 1 2; do
  $T 300 python bench.py --cpu-images 0 > $O/ab_a_default_$round.json 2> $O/ab_a_default_$round.err
  DD3D_CONV_CTA2_MINN=64 $T 300 python bench.py --cpu-images 0 > $O/ab_b_minn64_$round.json 2> $O/ab_b_minn64_$round.err
done
DD3D_CONV_CTA2_MINN=64 $T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99_minn64.txt 2>&1
DD3D_CONV_CTA2_MINN=64 $T 300 python tools/opprof.py dla34 8 > $O/op_times_dla34_minn64.txt 2>&1
$T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
$T 300 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); s=d.get('secondary',{})
        print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['kernels_ms_per_step']['conv_igemm'], '| dla34', round(s.get('value',0),1), round(s.get('ms_per_step',0),3))
    except Exception as e: print(f, 'ERR', e)
PY
paste <(head -12 gpurun_out/r02o/op_times_v2_99.txt) <(head -12 gpurun_out/r02o/op_times_v2_99_minn64.txt)
paste <(head -12 gpurun_out/r02o/op_times_dla34.txt) <(head -12 gpurun_out/r02o/op_times_dla34_minn64.txt)
