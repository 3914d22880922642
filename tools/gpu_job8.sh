#!/bin/bash
# round-2 GPU job 8 (re-entry): validate HEAD (class-parallel NMS commit): full GPU suite, smoke, default bench, per-op times
O=gpurun_out/r02h
mkdir -p $O
T="timeout -k 10"
$T 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv or nms or decode" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -3 $O/canary.log
if [ $rc -ne 0 ]; then echo "CANARY FAILED -- aborting"; exit 1; fi
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
$T 400 python bench.py --cpu-images 0 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
$T 200 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
$T 200 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h/bench_default.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks'],d['roofline']['frac'],d['kernels_ms_per_step'])
print('dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('roofline',{}).get('frac'),s.get('kernels_ms_per_step'))
PY
