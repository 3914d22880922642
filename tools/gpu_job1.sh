#!/bin/bash
# round-2 GPU job 1: determinism probes (fresh processes, PDL on/off, zeroed / poisoned / dirty workspace), sanitizers
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
run() { "$@" 2>&1 | grep -E "^PROBE|Error|error|Traceback" >> $O/probe.log; }
for i in 1 2 3; do run python tools/determinism_probe.py --case dla34 --label plain$i; done
run python tools/determinism_probe.py --case dla34 --fill 0 --label fill00
run python tools/determinism_probe.py --case dla34 --fill 255 --label fillff
run python tools/determinism_probe.py --case dla34 --dirty 8192 --label dirty
DD3D_NO_PDL=1 run python tools/determinism_probe.py --case dla34 --label nopdl
DD3D_NO_PDL=1 run python tools/determinism_probe.py --case dla34 --fill 255 --label nopdl_fillff
for c in v2_99 dla34_full v2_99_full; do
  run python tools/determinism_probe.py --case $c --fill 0 --label fill00
  run python tools/determinism_probe.py --case $c --fill 255 --label fillff
  DD3D_NO_PDL=1 run python tools/determinism_probe.py --case $c --fill 255 --label nopdl_fillff
done
timeout 900 compute-sanitizer --tool initcheck --kernel-regex kns=dd3d --print-limit 40 --log-file $O/initcheck_dla34.log \
  python tools/determinism_probe.py --case dla34 --label initcheck > $O/initcheck_dla34.out 2>&1
echo "initcheck rc=$?" >> $O/probe.log
timeout 900 compute-sanitizer --tool racecheck --kernel-regex kns=dd3d --print-limit 40 --log-file $O/racecheck_dla34.log \
  python tools/determinism_probe.py --case dla34 --label racecheck > $O/racecheck_dla34.out 2>&1
echo "racecheck rc=$?" >> $O/probe.log
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=dd3d --print-limit 40 --log-file $O/memcheck_v2_99.log \
  python tools/determinism_probe.py --case v2_99 --label memcheck > $O/memcheck_v2_99.out 2>&1
echo "memcheck rc=$?" >> $O/probe.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/probe.log
timeout 600 python bench.py > $O/bench_v2_99.json 2> $O/bench_v2_99.err
timeout 600 python bench.py --workload dla34 > $O/bench_dla34.json 2> $O/bench_dla34.err
tail -3 $O/pytest.log; cat $O/bench_v2_99.json | cut -c1-600
