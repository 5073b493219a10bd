"""Host-side logic of the drop-in (no GPU): constructor parity with the reference (golden KATs), state_dict
layout, argument checking, the C ABI surface, and the world-size-2 gradient exchange over gloo."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import sincformer_metacog_speech_enhancement_b200 as sb
from sincformer_metacog_speech_enhancement_b200 import _lib, build
from conftest import ROOT


def _parse_key(key):
    c, k, sr = key.split("_")
    return int(c[1:]), int(k[1:]), (None if sr[2:] == "None" else int(sr[2:]))


def test_constructor_matches_reference_bit_for_bit(init_kat):
    for key in sorted({k.split("/")[0] for k in init_kat}):
        C, K, sr = _parse_key(key)
        m = sb.SincConv1d(C, K, sample_rate=sr)
        assert m.kernel_size == int(init_kat[f"{key}/kernel_size"])
        assert float(m.sample_rate) == float(init_kat[f"{key}/sample_rate"])
        sd = m.state_dict()
        assert list(sd.keys()) == ["low_hz_", "band_hz_", "window", "n_"]
        for n in sd:
            assert sd[n].dtype == torch.float32
            assert np.array_equal(sd[n].numpy(), init_kat[f"{key}/{n}"]), (key, n)
    assert [n for n, _ in sb.SincConv1d(4, 5).named_parameters()] == ["low_hz_", "band_hz_"]


def test_constructor_conventions():
    assert sb.SincConv1d(8, 250).kernel_size == 251          # even K is bumped (perception.py:44-45)
    assert sb.SincConv1d(8, 251).kernel_size == 251
    assert sb.SincConv1d(8, 11, sample_rate=0).sample_rate == 8000   # falsy -> config default (:41)
    m = sb.SincConv1d(64, 251, sample_rate=16000)
    assert m.low_hz_.shape == (64, 1) and m.band_hz_.shape == (64, 1)
    assert m.window.shape == (251,) and m.n_.shape == (1, 125)
    assert sb.SincConv1d(4, 1).n_.shape == (1, 0)
    with pytest.raises(ValueError):
        sb.SincConv1d(4, 5, out_dtype=torch.float16)


def test_state_dict_round_trip_with_reference_layout(init_kat):
    key = "C64_K251_sr16000"
    ref_sd = {n: torch.from_numpy(init_kat[f"{key}/{n}"]) for n in ("low_hz_", "band_hz_", "window", "n_")}
    m = sb.SincConv1d(64, 251, sample_rate=16000)
    with torch.no_grad():
        m.low_hz_.add_(1.0)
    m.load_state_dict(ref_sd, strict=True)
    assert torch.equal(m.low_hz_.detach(), ref_sd["low_hz_"])

    class Holder(torch.nn.Module):   # PerceptionAgent keeps it as `sinc_conv`
        def __init__(self):
            super().__init__()
            self.sinc_conv = sb.SincConv1d(64, 251, sample_rate=16000)
    assert sorted(Holder().state_dict()) == sorted(f"sinc_conv.{n}" for n in ref_sd)


def test_no_cpu_path_and_shape_rules():
    m = sb.SincConv1d(4, 9)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(2, 1, 50))
    with pytest.raises(RuntimeError):
        m(torch.randn(3, 50))           # 2-D with 3 "channels": the reference's conv1d raises too
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 2, 50))        # in_channels != 1


def test_install_into_reference_rebinds_the_class():
    class FakePerception:               # stands in for the reference's agents.perception module object
        SincConv1d = object
    prev = sb.install_into_reference(FakePerception)
    assert prev is object and FakePerception.SincConv1d is sb.SincConv1d
    sb.install_into_reference(FakePerception, out_dtype=torch.bfloat16)
    assert FakePerception.SincConv1d(4, 9).out_dtype == torch.bfloat16


def test_abi_library_exports_every_declared_symbol():
    build.build_library()
    header = open(os.path.join(ROOT, "include", "sinc_b200.h")).read()
    body = header.split('extern "C" {', 1)[1]
    declared = set(re.findall(r"\b(sinc_[a-z_0-9]+)\s*\(", body))
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    lib = ctypes.CDLL(build.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    # hardware probes / micro-benchmarks live in their own library; the product library must not export them
    build.build_diag_library()
    dheader = open(os.path.join(ROOT, "include", "sinc_b200_diag.h")).read()
    ddeclared = set(re.findall(r"\b(sinc_[a-z_0-9]+)\s*\(", dheader.split('extern "C" {', 1)[1]))
    assert ddeclared == set(_lib.DIAG_SYMBOLS), ddeclared ^ set(_lib.DIAG_SYMBOLS)
    dlib = ctypes.CDLL(build.DIAG_LIB_PATH)
    for sym in ddeclared:
        assert hasattr(dlib, sym), sym
        assert not hasattr(lib, sym), f"{sym} is a diagnostic: it must not be exported from libsinc_b200.so"
    # no process-wide setters in the ABI (per-call sinc_options instead)
    for gone in ("sinc_set_engine", "sinc_get_engine", "sinc_set_reserved_sms", "sinc_last_engine_mask", "sinc_tc_status",
                 "sinc_tc_phase_stamps"):
        assert not hasattr(lib, gone), gone


def test_abi_non_compute_calls():
    lib = _lib.load()
    assert lib.sinc_abi_version() == 2
    assert lib.sinc_strerror(0) == b"success"
    assert b"NULL" in lib.sinc_strerror(-1)
    assert b"odd K" in lib.sinc_strerror(-2)
    assert lib.sinc_workspace_bytes(32, 48000, 64, 251) >= 64 * 256 * 4
    # the workspace also holds the self-validating filter table (header + W[64][256]) behind the partial tiles
    assert lib.sinc_workspace_bytes(1, 1, 4, 3) >= 4096 + 64 * 256 * 4
    assert lib.sinc_workspace_bytes(1, 1, 70, 300) > lib.sinc_workspace_bytes(1, 1, 64, 251)
    for code, name in enumerate(("auto", "fma", "tc", "tc_tf32")):      # every engine name the package accepts
        sb.set_engine(name)
        assert sb.get_engine() == name and _lib.ENGINE_CODES[name] == code
    with pytest.raises(ValueError):
        sb.set_engine("cpu")
    sb.set_engine("auto")
    with pytest.raises(ValueError):
        sb.SincConv1d(4, 9, engine="cpu")
    assert sb.SincConv1d(4, 9, engine="fma").engine == "fma"
    # argument validation happens before any CUDA call
    fb = _lib.FilterBank()
    assert lib.sinc_forward(ctypes.byref(fb), None, 0, None, 0, 1, 1, None, 0, None, None) == -1
    fb.low_hz = fb.band_hz = fb.window = fb.n = 16
    fb.C, fb.K = 4, 4
    assert lib.sinc_forward(ctypes.byref(fb), None, 0, None, 0, 1, 1, None, 0, None, None) == -2   # even K


def test_shard_rows_partitions_the_batch():
    for n, w in [(1024, 8), (32, 3), (5, 8), (7, 2)]:
        spans = [sb.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def _allreduce_worker(rank, world, port, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = sb.SincConv1d(6, 9)
    m.low_hz_.grad = torch.full_like(m.low_hz_, float(rank + 1))
    if rank == 0:
        m.band_hz_.grad = torch.arange(6.0).view(6, 1)      # rank 1 has no band grad yet: treated as zero
    sb.allreduce_filter_grads(m)
    out[rank] = (m.low_hz_.grad.clone(), m.band_hz_.grad.clone())
    dist.destroy_process_group()


def test_allreduce_filter_grads_world_size_2_gloo():
    world, port = 2, 29000 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_allreduce_worker, args=(world, port, out), nprocs=world, join=True)
        for r in range(world):
            lo, bd = out[r]
            assert torch.equal(lo, torch.full((6, 1), 3.0))
            assert torch.equal(bd, torch.arange(6.0).view(6, 1))


def test_bench_reference_arm_runs_on_cpu_and_ours_refuses_without_a_gpu():
    """`bench.py --impl reference` (the staged reference module -- or the oracle port when baseline/_ref is absent --
    timed on the host cores) prints the contract's JSON line with the same `config` keys as our arm; our own arm has no
    CPU fallback and must say so instead of printing a number."""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    p = subprocess.run([sys.executable, bench, "--impl", "reference", "--workload", "cfg1_b1_t16000_fwd", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "audio-s/s"
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "reference", "agents", "perception.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench as bench_module
    assert line["config"] == bench_module.workload_config("cfg1_b1_t16000_fwd", 1)    # what our arm prints as `config`
    if not torch.cuda.is_available():
        q = subprocess.run([sys.executable, bench, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
        assert q.returncode != 0 and "no CPU fallback" in (q.stdout + q.stderr)
        assert not [l for l in q.stdout.splitlines() if l.startswith("{")]


def test_bench_reference_arm_picks_a_thread_count_that_fits_the_host():
    """The reference arm times the reference 'with all the host threads it can use' -- which is not one thread per visible
    CPU on every box (the 32-CPU multi-GPU leases of this pool ran it 10x slower that way): the count is calibrated."""
    import sys
    sys.path.insert(0, ROOT)
    import bench as bench_module
    n = bench_module.usable_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)
    cls, _ = bench_module.reference_module_class()
    before = torch.get_num_threads()
    try:
        best = bench_module.best_cpu_threads(cls, 16000)
        assert 1 <= best <= n
        assert line_cores_ok(bench_module, best)
    finally:
        torch.set_num_threads(before)


def line_cores_ok(bench_module, best):
    # the `cores` the JSON line reports is what torch really runs with after the calibration
    torch.set_num_threads(best)
    return torch.get_num_threads() == best
