#!/bin/bash
# round-2 GPU job 3: 8-warp lean epilogue -- correctness (all GPU tests incl. full-shape parity), bench, per-op times, ncu
O=gpurun_out/r02c
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
timeout 600 python bench.py --cpu-images 0 > $O/bench_v2_99.json 2> $O/bench_v2_99.err
timeout 600 python bench.py --workload dla34 --cpu-images 0 > $O/bench_dla34.json 2> $O/bench_dla34.err
timeout 600 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
timeout 600 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c/bench_*.json')):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', 'frac', round(d['roofline']['frac'],3), d['clocks'], d['kernels_ms_per_step'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:conv_igemm -s 244 -c 122 --csv --log-file $O/conv_launches_v2_99.csv \
  python bench.py --steps 1 --warmup 3 --cpu-images 0 > $O/ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -2 $O/conv_launches_v2_99.csv | cut -c1-300
