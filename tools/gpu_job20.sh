#!/bin/bash
# round-2 GPU job 20: measured end-to-end parity of the final kernels (tools/parity_report.py -> profiles/parity_r02.json)
O=gpurun_out/r02t
mkdir -p $O
timeout -k 10 900 python tools/parity_report.py --out $O/parity_r02.json > $O/parity.log 2>&1
echo "parity rc=$?"; tail -12 $O/parity.log
