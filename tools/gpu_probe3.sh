#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01e}
timeout 600 python tools/tc_probe.py map_a_128_256 map_a_256_128 map_a_128_128 map_b_128_256 map_b_128_16 map_b_16_128 map_a_128_416 > gpurun_out/${TAG}_tcprobe.log 2>&1; echo "probe exit $?"; cat gpurun_out/${TAG}_tcprobe.log
