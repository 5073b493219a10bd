"""Does traffic from the other warps of the CTA slow tcgen05.mma down?  N = 32 / 64 MMAs from a lean stream (pattern 1)
while warps 1-3 do nothing / tcgen05.st / LDS.128 / tcgen05.ld (sinc_tc_probe.cu, RateArgs::stress)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
names = {0: "idle", 1: "tcgen05.st", 2: "LDS.128", 3: "tcgen05.ld"}
for N in (32, 64, 112):
    for stress in (0, 1, 2, 3):
        cyc = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        count = 51 * 2000
        flags = 1 | 2 | (1 << 2) | (stress << 4) | (148 << 8)
        for rep in range(2):
            rc = lib.sinc_tc_mma_rate(N, 1, count, flags, ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
            torch.cuda.synchronize()
        print(json.dumps({"N": N, "other_warps": names[stress], "rc": rc, "status": int(st.item()),
                          "cycles_per_mma": round(cyc.item() / count, 2)}), flush=True)
