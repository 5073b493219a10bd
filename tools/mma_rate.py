"""Cycles per tcgen05.mma (kind::f16, M=128) on one SM for the MMA shapes the kernels use (run on the GPU box)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
out = []
for a_tmem in (1, 0):
    for N in (32, 64, 112, 128, 256):
        for nacc in (1, 2, 4):
            if nacc * N > 448:
                continue
            cyc = torch.zeros(1, dtype=torch.int64, device=dev)
            st = torch.zeros(1, dtype=torch.int32, device=dev)
            count = 4096
            for _ in range(2):
                rc = lib.sinc_tc_mma_rate(N, nacc, count, a_tmem, ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
                torch.cuda.synchronize()
            out.append({"a": "tmem" if a_tmem else "smem", "N": N, "nacc": nacc, "rc": rc, "status": int(st.item()),
                        "cycles_per_mma": round(cyc.item() / count, 2), "ideal": N / 2})
            print(json.dumps(out[-1]), flush=True)
