"""Does alternating between accumulators cost tensor-pipe time?  (sinc_tc_probe.cu pattern 1, nacc = 1 / 2 / 3)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
names = {1: "one accumulator", 2: "alternating every MMA", 3: "runs of three per accumulator"}
for N in (32, 64):
    for nacc in (1, 2, 3):
        cyc = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        count = 51 * 2000
        flags = 1 | 2 | (1 << 2) | (148 << 8)
        for rep in range(2):
            rc = lib.sinc_tc_mma_rate(N, nacc, count, flags, ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
            torch.cuda.synchronize()
        print(json.dumps({"N": N, "accumulators": names[nacc], "rc": rc, "status": int(st.item()),
                          "cycles_per_mma": round(cyc.item() / count, 2)}), flush=True)
