#!/bin/bash
# final validation of a round on one B200: full GPU suite, default bench (both arms), cfg-3 harness, smoke.  usage: tools/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -rP > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "reference arm rc=$?"
timeout 300 python examples/cfg3_harness.py --out gpurun_out/${tag}_cfg3.json > gpurun_out/${tag}_cfg3.log 2>&1; echo "cfg3 rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["launch"], d.get("eager"), d["gpu_launches"])
print({k:(round(v["ms"],4), round(v["frac_of_pipe_peak"],3)) for k,v in d["kernels"].items()})
print(d["e2e"]["value"], d["e2e"]["launch"], d["e2e"]["ms_per_step_runs"]); print(d["sustained"]); print(d["gpu_reference"]); print(d["cpu_baseline"])
print(d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["traffic"]); print(d.get("cfg4")); print(d.get("cfg5"))
PY
