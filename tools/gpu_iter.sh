#!/bin/bash
# One iteration on the GPU: all GPU tests, the bench line, then one ncu --set full capture of each tcgen05 kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r01n}
KREGEX=${2:-tc_kernel}
NCAP=${3:-3}
# pre-flight: one small backward on the tensor-core engine under a short timeout -- a protocol bug must not eat the budget
timeout 150 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "backward and 1000" > gpurun_out/${TAG}_preflight.log 2>&1
rc=$?; echo "preflight exit $rc"; if [ $rc -ne 0 ]; then tail -30 gpurun_out/${TAG}_preflight.log; exit 1; fi
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | head -20; grep -E "^E  " gpurun_out/${TAG}_pytest.log | head -20
timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    for k, v in d["kernels"].items(): print(" ", k, v)
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $NCAP -c $NCAP -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/${TAG}_ncu2.log | cut -c1-200
fi
