mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --sustained 200 > gpurun_out/r06_n2.json 2> gpurun_out/r06_n2.err; echo rc=$?
tail -c 300 gpurun_out/r06_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r06_n2.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d["launch"], d.get("eager"))
print(d["e2e"]["value"], d["e2e"]["launch"], d["e2e"]["ms_per_step_runs"])
print(d.get("cfg4")); print(d.get("cfg5"))
PY
