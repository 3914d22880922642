#!/bin/bash
# round-2 GPU job 12: weight-stationary halo convs (64->64), eSE scale + max-pool fusion, NMS rank sort, stem_mma with permuted channels / direct stores, sparse-box3d auto
# policy, NMS scan / mask tweaks: canaries, A/B, conv launch table, full suite
O=gpurun_out/r02l
mkdir -p $O
T="timeout -k 10"
$T 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv or nms or decode or dla_front or stem or ese or maxpool" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -12 $O/canary.log
$T 600 python -m pytest tests/test_e2e_gpu.py -x -q > $O/canary2.log 2>&1
rc2=$?; echo "canary2 rc=$rc2"; tail -15 $O/canary2.log
if [ $rc -ne 0 ]; then echo "kernel canary failed"; fi
for round in 1 2; do
  DD3D_CONV_WSTAT=0 DD3D_ESE_POOL=0 $T 300 python bench.py --cpu-images 0 > $O/ab_a_old_$round.json 2> $O/ab_a_old_$round.err
  $T 300 python bench.py --cpu-images 0 > $O/ab_b_new_$round.json 2> $O/ab_b_new_$round.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); s=d.get('secondary',{})
        print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['kernels_ms_per_step'], round(d['roofline']['frac'],3))
        print('    dla34', round(s.get('value',0),1), round(s.get('ms_per_step',0),3), s.get('kernels_ms_per_step'), s.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
$T 200 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
$T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
$T 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"conv_igemm|conv_taps" -s 121 -c 121 --csv --log-file $O/conv_launches_v2_99.csv \
  python tools/one_forward.py v2_99 32 2 > $O/ncu_conv.log 2>&1
echo "ncu conv rc=$?"
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"stem|maxpool|ese|preprocess|nms|decode|dense|select|clear|b3d" --csv --log-file $O/launches_small_v2_99.csv python tools/one_forward.py v2_99 32 2 > $O/ncu_small.log 2>&1
$T 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_dla34.csv python tools/one_forward.py dla34 8 2 > $O/ncu_dla34.log 2>&1
if [ $rc -eq 0 ] && [ $rc2 -eq 0 ]; then
  ( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
  echo "pytest rc=$?"; tail -8 $O/pytest.log
  $T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
fi
