#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, context peaks, ncu launch list + full capture.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01a}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1; cat gpurun_out/${TAG}_bench_ref.json
timeout 300 python tools/probe_peaks.py > gpurun_out/${TAG}_peaks.json 2>&1; cat gpurun_out/${TAG}_peaks.json
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
  timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_kernel\|finalize\|tc_kernel -s 4 -c 4 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
  echo "ncu full exit $?"
fi
