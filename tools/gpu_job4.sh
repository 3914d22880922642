#!/bin/bash
# round-2 GPU job 4: 8-warp epilogue (hang fixed) + taps-in-N predictors + small-kernel tuning.  Every step is bounded by
# `timeout -k`; a canary (kernel-level conv tests) runs first and aborts the job if it fails or hangs.
O=gpurun_out/r02d
mkdir -p $O
T="timeout -k 10"
$T 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -4 $O/canary.log
if [ $rc -ne 0 ]; then echo "CANARY FAILED -- aborting"; exit 1; fi
$T 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest.log
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
$T 400 python bench.py --cpu-images 0 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
DD3D_CONV_TAPS=0 $T 300 python bench.py --cpu-images 0 --no-secondary > $O/bench_v2_99_notaps.json 2> $O/bench_v2_99_notaps.err
$T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
$T 120 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d/bench_*.json')):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['clocks'], d['kernels_ms_per_step'])
        if 'secondary' in d:
            s=d['secondary']; print('  secondary', round(s['value'],1), 'img/s', round(s['ms_per_step'],3), 'ms frac', round(s['roofline']['frac'],3), s['kernels_ms_per_step'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
$T 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"conv_igemm|conv_taps" -s 244 -c 122 --csv --log-file $O/conv_launches_v2_99.csv \
  python bench.py --steps 1 --warmup 3 --cpu-images 0 --no-secondary > $O/ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -2 $O/conv_launches_v2_99.csv | cut -c1-300
