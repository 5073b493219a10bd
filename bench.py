#!/usr/bin/env python
"""bench.py -- SincConv fwd+bwd throughput in audio-seconds/second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine auto|fma|tc]

One "step" = one forward + backward (grad-input + parameter grads) of the SincConv hot path over one batch of
synthetic waveforms.  Workload per rank: BASELINE config[1] -- batch 32 x 3 s @ 16 kHz (48 000 samples), fp32,
C=64 filters, K=251 taps; with N ranks the global batch is 32*N (weak scaling) and the 2*C-float parameter
gradient is all-reduced over NCCL every step.  audio-seconds = B*T/16000.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline` (dominant kernel, measured live with
CUDA events around direct C-ABI launches), `kernels` (all three), `cpu_baseline` (the oracle port of the
reference's CPU torch path on this box's host cores), `e2e` (module API with pinned host input, H2D + D2H inside
the timed region), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR_AUDIO = 16000
C, K = 64, 251
FLOP_FWD = 2 * C * K            # per input sample
FLOP_PER_SAMPLE = 3 * FLOP_FWD  # fwd + grad-filter + grad-input = 96 384
BYTES_PER_SAMPLE_F32 = 524      # fwd 4+256, bwd 256+4+4

WORKLOADS = {
    # name: (B per rank, T, out dtype, sample_rate of the module, needs input grad)
    "cfg2_b32_t48000_fp32": (32, 48000, "fp32", 16000, True),
    "cfg1_b1_t16000_fwd": (1, 16000, "fp32", 8000, True),
    "cfg4_b128_t48000_fp32": (128, 48000, "fp32", 16000, True),
    "cfg5_b8_t960000_bf16": (8, 960000, "bf16", 16000, True),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pw.append(0.0)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
            busy_pw = [w for v, w in zip(sm, pw) if v >= 0.5 * max(sm) and w > 0]
            if busy_pw:
                out["power_w"] = max(busy_pw)
        out["reasons"] = sorted(reasons)
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU torch path
# ------------------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(B, T, sr):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sinc_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    m = O.OracleSincConv1d(C, K, sample_rate=sr)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 1, T, generator=gen)
    g = torch.randn(B, C, T, generator=gen)

    def step():
        xr = x.clone().requires_grad_(True)
        for p in m.parameters():
            p.grad = None
        y = m(xr)
        y.backward(g)
        return float(m.low_hz_.grad.sum())
    return step


def time_cpu(step, reps, warm=1):
    for _ in range(warm):
        step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    B, T, dt, sr, _ = WORKLOADS[args.workload]
    Bs = min(B, 32)
    step = cpu_oracle_step_fn(Bs, T, sr)
    ts = time_cpu(step, args.steps, warm=min(args.warmup, 2))
    total = sum(ts)
    audio = Bs * T / SR_AUDIO
    value = audio * len(ts) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "sincconv_fwd_bwd_audio_seconds_per_second", "value": value,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_step": Bs, "samples": T, "filters": C, "taps": K,
                   "note": "reference CPU torch path (oracle port: same torch ops in the same order), all host threads"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"B={Bs} x T={T} fwd+bwd (param + input grads), {len(ts)} steps"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import sincformer_metacog_speech_enhancement_b200 as sb
    from sincformer_metacog_speech_enhancement_b200 import _lib
    from sincformer_metacog_speech_enhancement_b200.sinc import _bank

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the SincConv path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    sb.set_engine(args.engine)
    peaks = load_peaks()

    B, T, dt, sr, need_x = WORKLOADS[args.workload]
    out_dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    m = sb.SincConv1d(C, K, sample_rate=sr, out_dtype=out_dtype).to(dev)
    if world > 1:
        m.enable_grad_sync()   # 2*C-float SUM all-reduce inside backward, overlapped with the grad-input kernel
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = torch.randn(B, 1, T, generator=gen).pin_memory()
    x = x_host.to(dev)
    g = torch.randn(B, C, T, device=dev, dtype=torch.float32).to(out_dtype)
    audio_per_step = B * T / SR_AUDIO

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        xr = x.detach().requires_grad_(need_x)
        m.low_hz_.grad = None
        m.band_hz_.grad = None
        y = m(xr)
        y.backward(g)
        return xr.grad

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # --- headline: whole step, inputs resident in HBM --------------------------------------------------
    with ClockSampler(local) as clk:
        n0 = sb.kernel_launch_count()
        ms_total = timed(step_resident, args.steps, args.warmup)
        launches = (sb.kernel_launch_count() - n0) * args.steps // (args.steps + args.warmup)
    clocks = clk.summary()
    # nvidia-smi reports the clock the driver ASKS for; under the combined tensor + HBM load of this path the SMs run
    # slower (power management).  The kernels stamp the SM cycle counter and the nanosecond timer at entry and exit
    # (CTA 0, last launch of the timed region): cycles / ns is the clock they really ran at.
    try:
        stamps = (ctypes.c_longlong * 20)()
        if lib.sinc_tc_phase_stamps(stamps) == 0:
            real = {}
            for k, name in enumerate(("fwd", "bwd_input", "bwd_filter")):
                c0, t0, c2, t2 = stamps[6 * k], stamps[6 * k + 1], stamps[6 * k + 4], stamps[6 * k + 5]
                if t2 > t0 and c2 > c0:
                    real[name] = round((c2 - c0) / ((t2 - t0) / 1e3), 1)
            if real:
                clocks["sm_mhz_in_kernel"] = real
    except Exception:
        pass
    ms_step = ms_total / args.steps
    value = world * audio_per_step / (ms_step * 1e-3)

    # --- e2e: module API, pinned host input -> H2D, D2H of grad-input and parameter grads ---------------
    # A training loop's input pipeline: the H2D copy of step k+1 and the D2H read-back of step k-1 run on a
    # copy stream while step k computes (two device buffers); every byte still moves inside the timed region.
    gx_host = torch.empty(B, 1, T).pin_memory()
    gp_host = torch.empty(2, C, 1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    xdev = [torch.empty(B, 1, T, device=dev) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    buf_free = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0, "primed": False}

    def issue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(buf_free[slot])
            xdev[slot].copy_(x_host, non_blocking=True)
            h2d_done[slot].record(copy_stream)

    def step_e2e():
        main = torch.cuda.current_stream(dev)
        k = state["k"]
        slot = k & 1
        if not state["primed"]:
            for sl in range(2):
                buf_free[sl].record(main)
            issue_h2d(slot)
            state["primed"] = True
        issue_h2d(slot ^ 1)                      # prefetch the next step's waveforms
        main.wait_event(h2d_done[slot])
        xd = xdev[slot].requires_grad_(need_x)
        m.low_hz_.grad = None
        m.band_hz_.grad = None
        y = m(xd)
        y.backward(g)
        done = torch.cuda.Event()
        done.record(main)
        gxd = xd.grad if need_x else None
        glow, gband = m.low_hz_.grad, m.band_hz_.grad
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            if need_x:
                gx_host.copy_(gxd, non_blocking=True)
                gxd.record_stream(copy_stream)
            gp_host[0].copy_(glow, non_blocking=True)
            gp_host[1].copy_(gband, non_blocking=True)
            glow.record_stream(copy_stream)
            gband.record_stream(copy_stream)
            buf_free[slot].record(copy_stream)
        xdev[slot].requires_grad_(False)
        xdev[slot].grad = None
        state["k"] = k + 1

    def timed_e2e(steps, warmup):
        for _ in range(warmup):
            step_e2e()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e()
        torch.cuda.current_stream(dev).wait_stream(copy_stream)   # the last read-back is part of the region
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    e2e_runs = [timed_e2e(args.steps, args.warmup) / args.steps]
    if e2e_runs[0] > 1.5 * ms_step:      # host-link hiccup (seen on some boxes: 2-5 ms for 12 MB): measure once more
        e2e_runs.append(timed_e2e(args.steps, args.warmup) / args.steps)
    ms_e2e = min(e2e_runs)
    e2e_value = world * audio_per_step / (ms_e2e * 1e-3)

    def copy_gbs(dst, src, reps=5):      # the host link on its own, for context (pinned memory, copy stream idle)
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None or t < best else best
        return src.numel() * src.element_size() / best / 1e6
    link = {"h2d_gbs": copy_gbs(xdev[0], x_host), "d2h_gbs": copy_gbs(gx_host, xdev[0])}
    h2d = x_host.numel() * 4
    d2h = (gx_host.numel() * 4 if need_x else 0) + gp_host.numel() * 4

    line = {
        "metric": "sincconv_fwd_bwd_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dt == "fp32" else "f32 math, bf16 activations", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": B * world, "samples": T,
                   "filters": C, "taps": K, "module_sample_rate": sr, "engine": args.engine,
                   "grads": "input + low_hz_/band_hz_", "exchange": "allreduce(2*C fp32) over NCCL inside backward, overlapped with grad-input" if world > 1 else "none",
                   "l2_policy": "activations (393 MB y, 393 MB grad_y per step) exceed the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": ms_e2e, "ms_per_step_runs": e2e_runs,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_link": link,
                "note": "nn.Module API; every step copies its waveforms from pinned host memory and reads grad-input + parameter "
                        "grads back to pinned host memory (copy stream, double-buffered: copies overlap the neighbouring "
                        "steps' compute); grad_y resident (produced on-device by the next layer)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    # --- per-kernel timings through the C ABI (rank 0, N == 1 only) -------------------------------------
    if rank == 0 and world == 1:
        fb = _bank(m.low_hz_.detach(), m.band_hz_.detach(), m.window, m.n_, m.sample_rate, m.min_low_hz, m.min_band_hz)
        nbytes = lib.sinc_workspace_bytes(B, T, C, K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        y = torch.empty(B, C, T, dtype=out_dtype, device=dev)
        gx = torch.empty(B, 1, T, device=dev)
        gp = torch.empty(2, C, device=dev)
        code = _lib.DTYPE_BF16 if dt == "bf16" else _lib.DTYPE_F32
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def k_fwd():
            _lib.check(lib.sinc_forward(ctypes.byref(fb), P(x), T, P(y), code, B, T, P(ws), nbytes, stream), "fwd")

        def k_bwd_input():
            _lib.check(lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), code, P(gx), None, None, None, B, T,
                                         P(ws), nbytes, stream), "bwd_input")

        def k_bwd_filter():
            _lib.check(lib.sinc_backward(ctypes.byref(fb), P(x), T, P(g), code, None, P(gp[0]), P(gp[1]), None, B, T,
                                         P(ws), nbytes, stream), "bwd_filter")

        esize = 2 if dt == "bf16" else 4
        n = B * T
        kern = {}
        for name, fn, flops, byts in (
                ("fwd", k_fwd, FLOP_FWD * n, n * (4 + C * esize)),
                ("bwd_input", k_bwd_input, FLOP_FWD * n, n * (C * esize + 4)),
                ("bwd_filter+finalize", k_bwd_filter, FLOP_FWD * n, n * (C * esize + 4))):
            ms = timed(fn, max(3, args.steps // 2), 3) / max(3, args.steps // 2)
            mask = lib.sinc_last_engine_mask()
            tc_bit, h16_bit = {"fwd": (1, 8), "bwd_input": (2, 16)}.get(name, (4, 32))
            eng = "fma" if not (mask & tc_bit) else ("tc16" if (mask & h16_bit) else "tc_tf32")
            kern[name] = {"ms": ms, "tflops": flops / ms / 1e9, "gbs": byts / ms / 1e6, "engine": eng}
        line["kernels"] = kern
        dom = max(kern, key=lambda k: kern[k]["ms"])
        # Roofline of the dominant kernel.  The contraction is compute-bound on every pipe (184 FLOP/B).
        # FMA engine: FP32 pipe, nominal 148 SM x 128 lanes x 2 x clocks.max.sm (not in MEASURED_PEAKS).
        # tcgen05 engines: every algorithmic product costs three MMAs (hi*hi + hi*lo + lo*hi), so the ceiling for
        # ALGORITHMIC flops is the measured dense bf16 rate / 3 (kind::f16 operands: fp16 or bf16 split) or
        # / 2 / 3 (kind::tf32 operands).
        sm_max = clocks.get("sm_max_mhz") or 1965.0

        def peak_of(eng):
            if eng == "tc16":
                return peaks["bf16_tflops"] / 3, "tensor", (f"MEASURED_PEAKS bf16 {peaks['bf16_tflops']} TF/s ({peaks['source']}) / 3 "
                                                           "(16-bit split operands: three kind::f16 MMAs per product)")
            if eng == "tc_tf32":
                return peaks["bf16_tflops"] / 2 / 3, "tensor", (f"MEASURED_PEAKS bf16 {peaks['bf16_tflops']} TF/s ({peaks['source']}) / 2 (tf32) "
                                                               "/ 3 (3xTF32 split)")
            return 148 * 128 * 2 * sm_max * 1e6 / 1e12, "fp32_fma", (f"nominal 148 SM x 128 FMA lanes x 2 x {sm_max:.0f} MHz "
                                                                     "(FP32 pipe is not in MEASURED_PEAKS)")
        for k in kern:
            kern[k]["frac_of_pipe_peak"] = kern[k]["tflops"] / peak_of(kern[k]["engine"])[0]
        peak, bound, peak_src = peak_of(kern[dom]["engine"])
        # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (tools/ncu_traffic.py)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
                tr = json.load(fh)
            kname = {"fwd": "fwd", "bwd_input": "bwd_input", "bwd_filter+finalize": "bwd_filter"}[dom] + \
                    {"tc16": "_tc16_kernel", "tc_tf32": "_tc_kernel", "fma": "_fma_kernel"}[kern[dom]["engine"]]
            if args.workload == "cfg2_b32_t48000_fp32" and kname in tr:
                traffic = tr[kname]["dram_bytes"]
        except Exception:
            traffic = None
        line["roofline"] = {"kernel": dom, "bound": bound, "achieved": kern[dom]["tflops"], "peak": peak,
                            "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / peak, "traffic": traffic,
                            "peak_source": peak_src,
                            "hbm": {"achieved_gbs": kern[dom]["gbs"], "peak_gbs": peaks["hbm_gbs"],
                                    "frac": kern[dom]["gbs"] / peaks["hbm_gbs"]},
                            "step": {"algorithmic_tflops": FLOP_PER_SAMPLE * n / ms_step / 1e9,
                                     "algorithmic_gbs": (BYTES_PER_SAMPLE_F32 if dt == "fp32" else 268) * n / ms_step / 1e6}}
        if not args.no_cpu:
            Bs = min(B, 32)
            step = cpu_oracle_step_fn(Bs, T, sr)
            ts = time_cpu(step, 3, warm=1)
            line["cpu_baseline"] = {"value": Bs * T / SR_AUDIO / min(ts), "unit": "audio-s/s",
                                    "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"oracle port of the reference CPU torch path, B={Bs} x T={T} fwd+bwd, best of 3"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "fma", "tc", "tc_tf32"])
    ap.add_argument("--workload", default="cfg2_b32_t48000_fp32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
