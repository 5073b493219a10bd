"""Fused finalize (tail of the 16-bit grad-filter kernel) against the separate finalize_kernel, and where the tail's time
goes.  Needs the -DSINC_ABLATION build:  SINC_EXTRA_NVCC_FLAGS=-DSINC_ABLATION python tools/gf_tail.py"""
import ctypes, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import sincformer_metacog_speech_enhancement_b200 as sb
    from sincformer_metacog_speech_enhancement_b200 import _lib
    from sincformer_metacog_speech_enhancement_b200.sinc import _bank
    dev = torch.device("cuda:0"); lib = _lib.load()
    C, K, T, B = 64, 251, 48000, 32
    m = sb.SincConv1d(C, K, sample_rate=16000).to(dev)
    fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    x = torch.randn(B, 1, T, device=dev); g = torch.randn(B, C, T, device=dev); gp = torch.empty(2, C, device=dev)
    nbytes = lib.sinc_workspace_bytes(B, T, C, K); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    run = lambda: lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), 0, None, P(gp[0]), P(gp[1]), None, B, T, P(ws), nbytes, st, None)
    n = int(sys.argv[2])
    if n == 1:   # isolated launches: per-CTA timeline (SINC_ABLATE_GF=64)
        for _ in range(3):
            assert run() == 0
            torch.cuda.synchronize()
            print("----", flush=True)
        sys.exit(0)
    for _ in range(min(n, 5)): assert run() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"ms": round(e0.elapsed_time(e1) / n, 4), "sum": float(gp.double().abs().sum())}))
else:
    plans = (("timeline", {"SINC_ABLATE_GF": "96"}, 1),) if "--timeline" in sys.argv else None
    if "--tail" in sys.argv:
        plans = (("fused, tail timing", {"SINC_ABLATE_GF": "32"}, 3),)
    for label, env, n in plans or (("fused", {}, 20), ("separate finalize_kernel", {"SINC_SPLIT_FINALIZE": "1"}, 20),
                          ("fused, tail timing (one launch after warm-up)", {"SINC_ABLATE_GF": "32"}, 2),
                          ("fused, no flush guard", {"SINC_ABLATE_GF": "8"}, 20),
                          ("timeline", {"SINC_ABLATE_GF": "96"}, 1)):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child", str(n)], env=e, capture_output=True, text=True)
        if label == "timeline":
            runs = r.stdout.split("----")
            import re
            rows = [tuple(map(int, m)) for m in re.findall(r"cta (\d+) sm (\d+) start (\d+) mainend (\d+)", runs[-2])]
            t0 = min(x[2] for x in rows)
            rows.sort(key=lambda x: x[3])
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "gf_timeline.txt"), "w") as fh:
                for x in sorted(rows):
                    fh.write(f"{x[0]} {x[1]} {x[2] - t0} {x[3] - t0}\n")
            print("timeline of the last isolated launch (ns from the first CTA start): cta sm start mainend")
            for x in rows[:5] + rows[-25:]:
                print("  ", x[0], x[1], x[2] - t0, x[3] - t0)
            print("  main-end spread:", rows[0][3] - t0, "...", rows[-1][3] - t0, " starts:", min(x[2] for x in rows) - t0, max(x[2] for x in rows) - t0)
            print(" ".join(l for l in runs[-2].splitlines() if "tail" in l))
            continue
        print(f"{label}: {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)
