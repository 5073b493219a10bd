#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01g}
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "pytest tc exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest_tc.log | head -20; grep -E "^E  " gpurun_out/${TAG}_pytest_tc.log | head -20
CMD="python bench.py --engine tc --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 2 -c 2 -o gpurun_out/${TAG}_prof_tc $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/${TAG}_ncu2.log | cut -c1-300
