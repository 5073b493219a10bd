#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01m}
KREGEX=${2:-bwd_input_tc}
timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json | cut -c1-3000; tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/${TAG}_ncu2.log | cut -c1-200
