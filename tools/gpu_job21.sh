#!/bin/bash
# round-2 GPU job 21: programmatic dependent launch for the small kernels (pre / post-processing, pools, eSE): full suite, A/B
O=gpurun_out/r02u
mkdir -p $O
T="timeout -k 10"
$T 900 python -m pytest tests/test_determinism_gpu.py tests/test_kernels_gpu.py -x -q -k "pdl or nms or decode or ese or maxpool or poisoned" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -5 $O/canary.log
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
DD3D_NO_PDL=1 $T 300 python bench.py --cpu-images 0 > $O/ab_nopdl.json 2> $O/ab_nopdl.err
$T 300 python bench.py --cpu-images 0 > $O/ab_pdl.json 2> $O/ab_pdl.err
python - <<'PY'
import json
for f in ('ab_nopdl','ab_pdl'):
    d=json.loads([l for l in open(f'gpurun_out/r02u/{f}.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
    print(f, round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks']['sm_mhz'],round(d['roofline']['frac'],3),d['kernels_ms_per_step'])
    print('   dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('kernels_ms_per_step'))
PY
