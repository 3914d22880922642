#!/bin/bash
# round-2 GPU job 16: stem_mma with a 4-deep input ring: stem tests, full suite, smoke, default bench, launch list of the small kernels
O=gpurun_out/r02p
mkdir -p $O
T="timeout -k 10"
$T 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "stem" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -5 $O/canary.log
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
for round in 1 2; do
  $T 400 python bench.py --cpu-images 0 > $O/bench_default_$round.json 2> $O/bench_default_$round.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for r in (1,2):
    d=json.loads([l for l in open(f'gpurun_out/r02p/bench_default_{r}.json') if l.startswith('{')][-1]); s=d.get('secondary',{})
    print(round(d['value'],1),'img/s',round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1),d['clocks'],round(d['roofline']['frac'],3),d['kernels_ms_per_step'])
    print('dla34',round(s.get('value',0),1),s.get('ms_per_step'),s.get('roofline',{}).get('frac'),s.get('kernels_ms_per_step'))
PY
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"stem|maxpool|ese|preprocess|nms|decode|dense|select|clear|b3d" --csv --log-file $O/launches_small_v2_99.csv python tools/one_forward.py v2_99 32 2 > $O/ncu_small.log 2>&1
$T 200 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -x -q -k "stem_s2_mma" > $O/racecheck_stem.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_stem.log | tail -3
