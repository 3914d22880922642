#!/bin/bash
# round-2 GPU job 9: fused DLA-34 front end (dla_front.cu): parity, A/B against the layer-by-layer path, ncu launch list
O=gpurun_out/r02i
mkdir -p $O
T="timeout -k 10"
$T 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dla_front" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -15 $O/canary.log
$T 300 python -m pytest tests/test_e2e_gpu.py -x -q -k "dla" > $O/canary2.log 2>&1
rc2=$?; echo "canary2 rc=$rc2"; tail -15 $O/canary2.log
for round in 1 2; do
  DD3D_DLA_FRONT=0 $T 200 python bench.py --workload dla34 --cpu-images 0 --steps 50 > $O/ab_unfused_$round.json 2> $O/ab_unfused_$round.err
  $T 200 python bench.py --workload dla34 --cpu-images 0 --steps 50 > $O/ab_fused_$round.json 2> $O/ab_fused_$round.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['kernels_ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
$T 200 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
$T 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_tensor.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none --csv --log-file $O/launches_dla34.csv \
   python tools/one_forward.py dla34 8 2 > $O/ncu_dla34.log 2>&1
echo "ncu rc=$?"
$T 300 ncu --set full --import-source on --clock-control none -k regex:dla_front -c 1 -s 1 -o $O/prof_dla_front -f python tools/one_forward.py dla34 8 2 > $O/ncu_front.log 2>&1
echo "ncu full rc=$?"
if [ $rc -eq 0 ] && [ $rc2 -eq 0 ]; then
  ( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
  echo "pytest rc=$?"; tail -6 $O/pytest.log
fi
