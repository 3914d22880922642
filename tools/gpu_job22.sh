#!/bin/bash
# round-2 GPU job 22: PDL only behind our own kernels (PdlScope): determinism file twice in fresh processes, full suite, smoke, bench
O=gpurun_out/r02v
mkdir -p $O
T="timeout -k 10"
for i in 1 2; do
  $T 600 python -m pytest tests/test_determinism_gpu.py -x -q > $O/determinism_$i.log 2>&1
  echo "determinism run $i rc=$?"; tail -2 $O/determinism_$i.log
done
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
$T 300 python bench.py --cpu-images 0 > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02v/bench_default.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks']['sm_mhz'],round(d['roofline']['frac'],3),d['kernels_ms_per_step'])
print('   dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('kernels_ms_per_step'))
PY
