#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01b}
timeout 600 python tools/tc_probe.py > gpurun_out/${TAG}_tcprobe.log 2>&1; echo "probe exit $?"; cat gpurun_out/${TAG}_tcprobe.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/${TAG}_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_kernel\|finalize -s 4 -c 4 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/${TAG}_ncu2.log
