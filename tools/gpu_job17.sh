#!/bin/bash
# round-2 GPU job 17: ReLU fused into the conversion, division-free patch loads (stem_mma, dla_front): tests + bench + ncu durations
O=gpurun_out/r02q
mkdir -p $O
T="timeout -k 10"
$T 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "stem or dla_front" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -5 $O/canary.log
$T 900 python -m pytest tests/test_e2e_gpu.py tests/test_determinism_gpu.py -x -q > $O/canary2.log 2>&1
echo "e2e+determinism rc=$?"; tail -5 $O/canary2.log
for round in 1 2; do
  $T 400 python bench.py --cpu-images 0 > $O/bench_default_$round.json 2> $O/bench_default_$round.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for r in (1,2):
    d=json.loads([l for l in open(f'gpurun_out/r02q/bench_default_{r}.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
    print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks'],round(d['roofline']['frac'],3),d['kernels_ms_per_step'])
    print('dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('roofline',{}).get('frac'),s.get('kernels_ms_per_step'))
PY
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem|dla_front" --csv --log-file $O/launches_stem.csv python tools/one_forward.py v2_99 32 2 > $O/ncu_a.log 2>&1
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem|dla_front" --csv --log-file $O/launches_front.csv python tools/one_forward.py dla34 8 2 > $O/ncu_b.log 2>&1
grep -h "gpu__time" $O/launches_stem.csv $O/launches_front.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
