"""tcgen05.mma rate with the WHOLE chip busy: cycles per MMA (clock64 on SM 0) and nanoseconds per MMA (CUDA events)
for 1 vs 148 CTAs and zero vs non-zero operands.  cycles/ns is the SM clock the tensor pipe actually ran at."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
for N in (112, 64, 32, 256):
    for grid in (1, 148):
        for nonzero in (0, 1):
            cyc = torch.zeros(1, dtype=torch.int64, device=dev)
            st = torch.zeros(1, dtype=torch.int32, device=dev)
            count = 400000   # ~15 ms at N = 112: long enough for the power management to react
            flags = 1 | (nonzero << 1) | (grid << 8)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(2):
                e0.record()
                rc = lib.sinc_tc_mma_rate(N, 1, count, flags, ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            cpm = cyc.item() / count
            print(json.dumps({"N": N, "ctas": grid, "nonzero": nonzero, "rc": rc, "status": int(st.item()),
                              "cycles_per_mma": round(cpm, 2), "ns_per_mma": round(ms * 1e6 / count, 2),
                              "implied_sm_mhz": round(cyc.item() / (ms * 1e3), 0),
                              "chip_tflops": round(grid * count * 2 * 128 * N * 16 / (ms * 1e9), 1)}), flush=True)
