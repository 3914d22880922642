#!/bin/bash
# round-2 GPU job 7: class-parallel NMS (tests + bench), ncu --set full of the three worst non-conv kernels
O=gpurun_out/r02g
mkdir -p $O
T="timeout -k 10"
$T 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv or nms or decode" > $O/canary.log 2>&1
rc=$?; echo "canary rc=$rc"; tail -3 $O/canary.log
if [ $rc -ne 0 ]; then echo "CANARY FAILED -- aborting"; exit 1; fi
$T 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest.log
for round in 1 2; do
  DD3D_NMS_CLASS_PARALLEL=0 $T 300 python bench.py --cpu-images 0 > $O/ab_nms1cta_$round.json 2> $O/ab_nms1cta_$round.err
  $T 300 python bench.py --cpu-images 0 > $O/ab_nmscls_$round.json 2> $O/ab_nmscls_$round.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g/ab_*.json')):
    try:
        d=json.loads(open(f).read()); s=d.get('secondary',{})
        print(f.split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], 'nms', d['kernels_ms_per_step']['nms'], '| dla34', round(s.get('value',0),1), round(s.get('ms_per_step',0),3), 'nms', s.get('kernels_ms_per_step',{}).get('nms'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
for k in stem_tc ese_scale maxpool; do
  $T 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o $O/prof_$k -f \
    python bench.py --steps 1 --warmup 1 --cpu-images 0 --no-secondary > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la $O/*.ncu-rep
