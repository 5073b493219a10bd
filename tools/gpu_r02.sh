#!/bin/bash
# one gpurun call of round 2: parity suite, bench, cfg-3 harness.  usage: tools/gpu_r02.sh <tag> [pytest args]
tag=${1:-r02x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 -rP "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench.err
python examples/cfg3_harness.py --out gpurun_out/${tag}_cfg3.json > gpurun_out/${tag}_cfg3.log 2>&1; echo "cfg3 rc=$?"
tail -c 800 gpurun_out/${tag}_cfg3.log
