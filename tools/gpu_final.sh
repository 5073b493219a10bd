#!/bin/bash
# Round-end evidence in one gpurun call: GPU tests, smoke, bench (ours + reference arm), context peaks, ncu launch
# list of one bench command and one ncu --set full capture of every hot kernel.  Outputs -> gpurun_out/<TAG>_*.
set -u
mkdir -p gpurun_out
TAG=${1:-r01final}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 150 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "backward and 1000" > gpurun_out/${TAG}_preflight.log 2>&1
rc=$?; echo "preflight exit $rc"; if [ $rc -ne 0 ]; then tail -30 gpurun_out/${TAG}_preflight.log; exit 1; fi
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
timeout 300 python tools/probe_peaks.py > gpurun_out/${TAG}_peaks.json 2>&1; cat gpurun_out/${TAG}_peaks.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc16_kernel\|finalize -s 4 -c 4 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/${TAG}_ncu2.log | cut -c1-200
