"""What does a tcgen05.commit in the instruction stream cost the tensor pipe?  Pattern 3 of the rate probe: the forward
kernel's MMA sequence with a commit (to a barrier nobody waits on) every k K-steps = every 3k MMAs."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sincformer_metacog_speech_enhancement_b200 import _lib
lib = _lib.load_diag()
dev = torch.device("cuda:0")
for N in (32, 64, 112):
    for k in (0, 1, 2, 4):
        cyc = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        count = 51 * 2000
        pattern = 3 if k else 1
        flags = 1 | 2 | (pattern << 2) | (148 << 8)
        for rep in range(2):
            rc = lib.sinc_tc_mma_rate(N, max(k, 1), count, flags, ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(st.data_ptr()), None)
            torch.cuda.synchronize()
        print(json.dumps({"N": N, "commit_every_mmas": 3 * k if k else None, "rc": rc, "status": int(st.item()),
                          "cycles_per_mma": round(cyc.item() / count, 2)}), flush=True)
