#!/bin/bash
# round-2 GPU job 18: stem_mma 256-bit stores: full suite, smoke, default bench, stem duration
O=gpurun_out/r02r
mkdir -p $O
T="timeout -k 10"
$T 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "stem" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -5 $O/canary.log
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
$T 400 python bench.py --cpu-images 0 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02r/bench_default.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks'],round(d['roofline']['frac'],3),d['kernels_ms_per_step'])
print('dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('roofline',{}).get('frac'),s.get('kernels_ms_per_step'))
PY
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"stem" --csv --log-file $O/launches_stem.csv python tools/one_forward.py v2_99 32 2 > $O/ncu_a.log 2>&1
grep -h "stem" $O/launches_stem.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-140
