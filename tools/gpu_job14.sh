#!/bin/bash
# round-2 GPU job 14: per-image sparse-box3d rule: full suite + smoke; ncu --set full of the 64->64 halo conv (VoVNet stem_2)
O=gpurun_out/r02n
mkdir -p $O
T="timeout -k 10"
( time $T 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log
$T 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
$T 400 ncu --set full --import-source on --clock-control none -k regex:conv_igemm_kernel -s 119 -c 1 -o $O/prof_stem2 -f python tools/one_forward.py v2_99 32 2 > $O/ncu_stem2.log 2>&1
echo "ncu stem2 rc=$?"
$T 400 ncu --set full --import-source on --clock-control none -k regex:conv_igemm_kernel -s 58 -c 1 -o $O/prof_dla_l2 -f python tools/one_forward.py dla34 8 2 > $O/ncu_dla_l2.log 2>&1
echo "ncu dla rc=$?"
ls -la $O/*.ncu-rep
