#!/bin/bash
# round-2 GPU job 23: ncu --set full of the dominant kernel of the final build (FCOS tower conv, 5 levels per launch, B = 32)
O=gpurun_out/r02w
mkdir -p $O
timeout -k 10 300 ncu --set full --import-source on --clock-control none -k regex:conv_igemm_kernel -s 226 -c 1 -o $O/prof_tower -f python tools/one_forward.py v2_99 32 2 > $O/ncu_tower.log 2>&1
echo "ncu rc=$?"; ls -la $O/*.ncu-rep
