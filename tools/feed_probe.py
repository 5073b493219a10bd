"""Ceiling of the HBM -> shared-memory operand feed on one B200 (diagnostics library): the (32,64,48128) fp32 activation
streamed by 148 persistent CTAs with no consumer, for several access patterns and depths; torch's own streaming kernels
for reference."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
B, C, T = 32, 64, 48128          # 94 x 512: whole tiles
g = torch.randn(B, C, T, device=dev); sink = torch.zeros(4, device=dev)
nbytes = g.numel() * 4
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
P = lambda t: ctypes.c_void_p(t.data_ptr())
names = {0: "bulk 64 x 512 B", 1: "bulk 64 x 1 KB", 2: "bulk 64 x 2 KB", 3: "LDG.128, 8 in flight per thread"}
for mode, depths in ((0, (2, 4, 6)), (1, (1, 2, 3)), (2, (1,)), (3, (1,))):
    for depth in depths:
        def run():
            assert lib.sinc_feed_probe(P(g), B, C, T, mode, depth, P(sink), None) == 0
        ms = timeit(run)
        kb = 0 if mode == 3 else depth * 64 * (128, 256, 512)[mode] * 4 // 1024
        print(json.dumps({"pattern": names[mode], "stages_in_flight": depth, "KB_in_flight_per_SM": kb, "ms": round(ms, 4),
                          "TB_per_s": round(nbytes / ms / 1e9, 3)}), flush=True)
y = torch.empty_like(g)
for name, fn, tot in (("torch sum (read)", lambda: g.sum(), nbytes), ("torch copy (read + write)", lambda: y.copy_(g), 2 * nbytes),
                      ("torch fill (write)", lambda: y.fill_(1.0), nbytes)):
    ms = timeit(fn)
    print(json.dumps({"pattern": name, "ms": round(ms, 4), "TB_per_s": round(tot / ms / 1e9, 3)}), flush=True)
