"""Cycles per wait on an already completed mbarrier phase (one thread, dependent chain)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
lib.sinc_mbar_wait_cost.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
names = {0: "try_wait + suspend hint", 1: "try_wait", 2: "test_wait", 3: "tc::mbar_wait"}
for mode in (0, 1, 2, 3):
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        rc = lib.sinc_mbar_wait_cost(mode, 10000, ctypes.c_void_p(cyc.data_ptr()), None)
        torch.cuda.synchronize()
    print(json.dumps({"wait": names[mode], "rc": rc, "cycles_per_wait": cyc[0].item() / 10000, "succeeded": int(cyc[1].item())}))
