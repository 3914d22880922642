#!/bin/bash
# round-2 GPU job 13: stem_mma double-buffer race fixed, NMS scan survivor-to-survivor, ese_pool default off: full suite,
# racecheck / memcheck of the round-2 kernels, default bench, config-5 sweep, nuScenes workload
O=gpurun_out/r02m
mkdir -p $O
T="timeout -k 10"
$T 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "nms or stem or ese" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -5 $O/canary.log
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
$T 400 python bench.py --cpu-images 0 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02m/bench_default.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks'],d['roofline']['frac'],d['kernels_ms_per_step'])
print('dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('roofline',{}).get('frac'),s.get('kernels_ms_per_step'))
PY
K='stem_s2_mma or dla_front or nms_one_dominant or ese_with_fused_pool or weight_stationary'
$T 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > $O/racecheck_kernels.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/racecheck_kernels.log | tail -5
$T 900 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > $O/memcheck_kernels.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_kernels.log | tail -4
$T 600 compute-sanitizer --tool racecheck python -m pytest tests/test_e2e_gpu.py -x -q -k "non_default_head_configs and PER_LEVEL" > $O/racecheck_sparse.log 2>&1
echo "racecheck sparse rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_sparse.log | tail -3
$T 400 python bench.py --workload dla34 --dtype fp16 --sweep 8,16,32,64 --cpu-images 0 > $O/sweep_dla34_fp16.jsonl 2> $O/sweep_fp16.err
$T 400 python bench.py --workload dla34 --dtype bf16 --sweep 8,16,32,64 --cpu-images 0 > $O/sweep_dla34_bf16.jsonl 2> $O/sweep_bf16.err
$T 300 python bench.py --workload nusc_v2_99 --cpu-images 0 > $O/bench_nusc.json 2> $O/bench_nusc.err
$T 300 python bench.py --dtype fp16 --cpu-images 0 --no-secondary > $O/bench_v2_99_fp16.json 2> $O/bench_fp16.err
python - <<'PY'
import json,glob
for f in ['sweep_dla34_fp16.jsonl','sweep_dla34_bf16.jsonl','bench_nusc.json','bench_v2_99_fp16.json']:
    try:
        for l in open('gpurun_out/r02m/'+f):
            if l.startswith('{'):
                d=json.loads(l); print(f, d['config'].get('global_batch'), round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d['dtype'], round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
