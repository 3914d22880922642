#!/bin/bash
# round-2 GPU job 2: probe under ncu (the r01 13-vs-15 anomaly), measured parity report, GPU tests incl. fp16, fp16 bench
O=gpurun_out/r02b
mkdir -p $O
python tools/determinism_probe.py --case dla34 --label plain 2>&1 | grep "^PROBE" > $O/probe_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/ncu_probe.csv \
  python tools/determinism_probe.py --case dla34 --label ncu 2>&1 | grep "^PROBE" > $O/probe_ncu.json
python -c "
import json
a=json.loads(open('$O/probe_plain.json').read()[6:]); b=json.loads(open('$O/probe_ncu.json').read()[6:])
ra, rb = a['runs'][0], b['runs'][0]
print('plain counts', ra['counts'], 'ncu counts', rb['counts'])
for k in ra:
    if k=='ops':
        d=[(x,y) for x,y in zip(ra['ops'],rb['ops']) if x!=y]; print('ops differing', len(d), d[:3])
    elif ra[k]!=rb[k]: print('DIFF', k, ra[k], rb[k])
" > $O/ncu_vs_plain.txt 2>&1
cat $O/ncu_vs_plain.txt
# smoke plain vs smoke under ncu, as the driver runs them
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/ncu_smoke.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_ncu.log 2>&1
grep -h "smoke:" $O/smoke_plain.log $O/smoke_ncu.log
timeout 1500 python tools/parity_report.py --out $O/parity_r02.json > $O/parity.log 2>&1
echo "parity rc=$?"; tail -12 $O/parity.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_full_gpu.py > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest.log
for b in 8 16 32 64; do
  timeout 300 python bench.py --workload dla34 --dtype fp16 --batch $b --cpu-images 0 > $O/bench_dla34_fp16_b$b.json 2> $O/bench_dla34_fp16_b$b.err
  timeout 300 python bench.py --workload dla34 --dtype bf16 --batch $b --cpu-images 0 > $O/bench_dla34_bf16_b$b.json 2> $O/bench_dla34_bf16_b$b.err
done
timeout 300 python bench.py --dtype fp16 --cpu-images 0 > $O/bench_v2_99_fp16.json 2> $O/bench_v2_99_fp16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b/bench_*.json')):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', 'frac', round(d['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
PY
