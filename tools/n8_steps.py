"""Per-step device times of the first steps at N ranks (torchrun): where does a short burst lose time?
   torchrun --nproc-per-node 8 tools/n8_steps.py [steps]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import sincformer_metacog_speech_enhancement_b200 as sb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
B, T, C, K = 32, 48000, 64, 251
x = torch.randn(B, 1, T, device=dev); g = torch.randn(B, C, T, device=dev)
warm = torch.zeros(2 * C, device=dev)
for _ in range(10): dist.all_reduce(warm)
torch.cuda.synchronize()


def run(sync, label):
    m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
    if sync: m.enable_grad_sync()
    def step():
        xr = x.detach().requires_grad_(True)
        m.low_hz_.grad = None; m.band_hz_.grad = None
        m(xr).backward(g)
    for _ in range(3): step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host = []
    evs[0].record()
    for i in range(steps):
        t0 = time.perf_counter()
        step()
        host.append((time.perf_counter() - t0) * 1e6)
        evs[i + 1].record()
    torch.cuda.synchronize()
    dt = torch.tensor([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(steps)], device=dev)
    mx = dt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    h = torch.tensor(host, device=dev); hm = h.clone(); dist.all_reduce(hm, op=dist.ReduceOp.MAX)
    if rank == 0:
        f = lambda t, a, b: round(float(t[a:b].mean()), 1)
        print(json.dumps({"case": label, "device_us_per_step_max_over_ranks": {"1-20": f(mx, 0, 20), "21-40": f(mx, 20, 40), "41-80": f(mx, 40, 80), f"81-{steps}": f(mx, 80, steps)},
                          "host_us_per_step_max_over_ranks": {"1-20": f(hm, 0, 20), "21-40": f(hm, 20, 40), "41-80": f(hm, 40, 80), f"81-{steps}": f(hm, 80, steps)},
                          "rank0_device_first_10": [round(float(v)) for v in dt[:10]]}), flush=True)


run(True, "all-reduce inside backward")
run(False, "no exchange")
run(True, "all-reduce inside backward (second time)")
dist.destroy_process_group()
