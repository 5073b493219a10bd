"""Role ablation of the three 16-bit tcgen05 kernels (an -DSINC_ABLATION build of the library; results are garbage by
construction): which role paces each kernel?   SINC_EXTRA_NVCC_FLAGS=-DSINC_ABLATION python tools/ablate.py [fw|bi|gf ...]
Also prints, for reference, what a plain device fill / copy of the same 393 MB activation achieves."""
import ctypes, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PLANS = {
    "fw": ("SINC_ABLATE_FW", {0: "full", 1: "epilogue idle (stores of stale staging)", 2: "no stores to y", 3: "epilogue idle + no stores",
                              4: "no MMAs", 8: "producers idle", 12: "no MMAs, producers idle (stores only)",
                              15: "barriers only"}),
    "bi": ("SINC_ABLATE_BI", {0: "full", 32: "full, boxes from L2 (perfect feed)", 3: "epilogue idle", 12: "producers idle (TMA boxes still arrive)",
                              15: "MMA issue only", 47: "MMA issue only, boxes from L2", 16: "no MMAs", 31: "barriers + TMA boxes only",
                              63: "barriers + TMA boxes from L2 only"}),
    "gf": ("SINC_ABLATE_GF", {0: "full", 8: "no wait for the previous window's flush (results may be wrong)"}),
}

if len(sys.argv) > 2 and sys.argv[1] == "child":
    import torch
    import sincformer_metacog_speech_enhancement_b200 as sb
    from sincformer_metacog_speech_enhancement_b200 import _lib
    from sincformer_metacog_speech_enhancement_b200.sinc import _bank
    which = sys.argv[2]
    dev = torch.device("cuda:0"); lib = _lib.load()
    C, K, T, B = 64, 251, 48000, 32
    m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
    fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    x = torch.randn(B, 1, T, device=dev); g = torch.randn(B, C, T, device=dev); gx = torch.empty(B, 1, T, device=dev)
    y = torch.empty(B, C, T, device=dev); gp = torch.empty(2, C, device=dev)
    nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if which == "fw":
        run = lambda: lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), 0, B, T, P(ws), nbytes, st, None)
    elif which == "bi":
        run = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, P(gx), None, None, None, B, T, P(ws), nbytes, st, None)
    elif which == "gf":
        run = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, None, P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None)
    else:   # reference: device fill and copy of the activation
        def timeit(fn):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 20
        nb = y.numel() * 4
        tf = timeit(lambda: y.fill_(1.0)); tcp = timeit(lambda: y.copy_(g)); ts = timeit(lambda: g.sum())
        print(json.dumps({"fill_ms": tf, "fill_TBs": nb / tf / 1e9, "copy_ms": tcp, "copy_TBs_rw": 2 * nb / tcp / 1e9,
                          "sum_ms": ts, "read_TBs": nb / ts / 1e9}))
        sys.exit(0)
    for _ in range(5): assert run() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"ms": round(e0.elapsed_time(e1) / 20, 4)}))
else:
    todo = sys.argv[1:] or ["ref", "fw", "bi", "gf"]
    for which in todo:
        if which == "ref":
            p = subprocess.run([sys.executable, __file__, "child", "ref"], capture_output=True, text=True)
            print("reference", p.stdout.strip() or p.stderr[-400:], flush=True)
            continue
        var, plan = PLANS[which]
        for mask, what in plan.items():
            env = dict(os.environ); env[var] = str(mask)
            p = subprocess.run([sys.executable, __file__, "child", which], env=env, capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            print(f"{which} mask {mask:2d} {what:45s}", line[-1] if line else p.stderr[-500:], flush=True)
