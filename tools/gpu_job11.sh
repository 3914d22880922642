#!/bin/bash
# round-2 GPU job 11: stem_mma (VoVNet stem_1 on register fragments), sparse box3d predictor, NMS mask/scan v2, decode_final spread, n_split fix: full suite + A/B + launch lists
O=gpurun_out/r02k
mkdir -p $O
T="timeout -k 10"
$T 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "nms or decode or dla_front or stem" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -12 $O/canary.log
$T 600 python -m pytest tests/test_e2e_gpu.py -x -q > $O/canary2.log 2>&1
rc2=$?; echo "canary2 rc=$rc2"; tail -15 $O/canary2.log
for round in 1 2; do
  DD3D_SPARSE_BOX3D=0 DD3D_STEM_MMA=0 $T 300 python bench.py --cpu-images 0 > $O/ab_a_dense_$round.json 2> $O/ab_a_dense_$round.err
  DD3D_STEM_MMA=0 $T 300 python bench.py --cpu-images 0 > $O/ab_b_sparse_$round.json 2> $O/ab_b_sparse_$round.err
  $T 300 python bench.py --cpu-images 0 > $O/ab_c_sparse_stem_$round.json 2> $O/ab_c_sparse_stem_$round.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); s=d.get('secondary',{})
        print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['kernels_ms_per_step'], round(d['roofline']['frac'],3))
        print('    dla34', round(s.get('value',0),1), round(s.get('ms_per_step',0),3), s.get('kernels_ms_per_step'), s.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
$T 200 python tools/opprof.py dla34 8 > $O/op_times_dla34.txt 2>&1
$T 300 python tools/opprof.py v2_99 32 > $O/op_times_v2_99.txt 2>&1
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms|decode|dense|select|clear|b3d" --csv --log-file $O/launches_post_v2_99.csv python tools/one_forward.py v2_99 32 2 > $O/ncu_v299.log 2>&1
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_dla34.csv python tools/one_forward.py dla34 8 2 > $O/ncu_dla34.log 2>&1
if [ $rc -eq 0 ] && [ $rc2 -eq 0 ]; then
  ( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
  echo "pytest rc=$?"; tail -8 $O/pytest.log
  $T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
fi
