#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01d}
timeout 600 python tools/tc_probe.py mnmajor_gradfilter mn_a_dense mn_b_dense mn_a_gtile mn_b_toeplitz mn_b_toeplitz_lbo16 > gpurun_out/${TAG}_tcprobe.log 2>&1; echo "probe exit $?"; cat gpurun_out/${TAG}_tcprobe.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
