mkdir -p gpurun_out
timeout 240 python bench.py --no-cpu --no-sub --no-gpu-ref --sustained 200 > gpurun_out/r06_n1.json 2> gpurun_out/r06_n1.err; echo rc=$?
tail -c 400 gpurun_out/r06_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r06_n1.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["launch"], d.get("eager"), d["gpu_launches"])
print(d["e2e"]["value"], d["e2e"]["launch"], d["e2e"]["ms_per_step_runs"])
PY
