#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01c}
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -m gpu > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "pytest tc exit $?"; tail -15 gpurun_out/${TAG}_pytest_tc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --engine tc --no-cpu > gpurun_out/${TAG}_bench_tc.json 2> gpurun_out/${TAG}_bench_tc.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench_tc.json; tail -3 gpurun_out/${TAG}_bench_tc.err
