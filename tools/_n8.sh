mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r05_n8.json 2> gpurun_out/r05_n8.err; echo rc=$?
tail -c 300 gpurun_out/r05_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r05_n8.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d["e2e"]["value"], d["e2e"]["ms_per_step_runs"])
print(d.get("sustained")); print(d.get("cfg4")); print(d.get("cfg5"))
PY
