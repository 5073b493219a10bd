"""the REAL reference PerceptionAgent with the B200 class swapped in at its only
construction site (agents/perception.py:152-156) and called through its own forward (:228-251).

The reference is imported unmodified -- from /root/reference in the build container, from the staged copy under
baseline/_ref/reference (oracle/stage_reference.py) on the GPU box.  Oracle for the comparison: the reference agent with
its own SincConv1d, evaluated on CPU in fp32 (outputs) and in fp64 (gradients; the reference's fp32 parameter gradients
of the layer are round-off noise).
"""
import copy
import os
import sys

import pytest
import torch

import sincformer_metacog_speech_enhancement_b200 as sb
from conftest import ROOT, max_rel

pytestmark = pytest.mark.gpu


def reference_root():
    for cand in (os.environ.get("SINC_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "reference")):
        if cand and os.path.exists(os.path.join(cand, "agents", "perception.py")):
            return cand
    return None


@pytest.fixture(scope="module")
def perception():
    root = reference_root()
    if root is None:
        pytest.skip("reference not available: run oracle/stage_reference.py in the build container")
    sys.path.insert(0, root)
    import agents.perception as perception_module
    yield perception_module
    sys.path.remove(root)


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_install_into_real_perception_agent(perception):
    prev_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False        # the agent's other layers run through cuDNN: keep them fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    reference_class = perception.SincConv1d
    try:
        torch.manual_seed(0)
        ref = perception.PerceptionAgent(sample_rate=16000)                    # the reference, untouched
        assert type(ref.sinc_conv) is reference_class
        replaced = sb.install_into_reference(perception)
        assert replaced is reference_class
        torch.manual_seed(0)
        pa = perception.PerceptionAgent(sample_rate=16000)                     # same source, new hot path
        assert type(pa.sinc_conv) is sb.SincConv1d
        assert set(pa.state_dict()) == set(ref.state_dict())                   # sinc_conv.low_hz_, .band_hz_, .window, .n_
        pa.load_state_dict(ref.state_dict())                                   # reference checkpoint loads as is
        ref.load_state_dict(pa.state_dict())                                   # ... and back
        pa = pa.to(dev())
        gen = torch.Generator().manual_seed(1)
        wav = 0.3 * torch.randn(3, 8000, generator=gen)                        # (batch, samples): forward unsqueezes (:228)
        gz = [torch.randn(3, 256, 500, generator=gen), torch.randn(3, 256, 500, generator=gen),
              torch.randn(3, 1, 500, generator=gen)]

        # reference on CPU: fp32 outputs, fp64 gradients
        out32 = ref(wav)
        ref64 = copy.deepcopy(ref).double()
        out64 = ref64(wav.double())
        sum((o * g.double()).sum() for o, g in zip(out64, gz)).backward()

        n0 = sb.kernel_launch_count()
        with pa.sinc_conv.capture() as rec:
            out = pa(wav.to(dev()))                                            # z_real, z_imag, sigma
            assert len(out) == 3
            sum((o * g.to(dev())).sum() for o, g in zip(out, gz)).backward()
            torch.cuda.synchronize()
        assert sb.kernel_launch_count() - n0 >= 2                              # forward, grad-filter (+ fused finalize): our kernels ran
        assert rec["fwd"] == "tc16" and rec["bwd_filter"] == "tc16" and "bwd_input" not in rec   # the waveform is data
        for name, o, o32, o64 in zip(("z_real", "z_imag", "sigma"), out, out32, out64):
            assert o.shape == o32.shape
            assert max_rel(o.detach().cpu(), o64.detach()) <= 2e-4, name      # five fp32 conv/GroupNorm stages downstream
            assert max_rel(o.detach().cpu(), o32.detach()) <= 4e-4, name
        # every parameter gradient of the agent against the fp64 reference
        got = dict(pa.named_parameters())
        worst = 0.0
        for name, p64 in ref64.named_parameters():
            assert got[name].grad is not None, name
            err = max_rel(got[name].grad.cpu(), p64.grad)
            if name.startswith("sinc_conv."):
                # grad_y reaches the layer through fp32 cuDNN layers (relative error ~1e-6) and the synthesis chain rule
                # amplifies a perturbation of the (C,K) grad-filter: gate 1e-4 + that conditioning floor
                assert err <= 2e-3, (name, err)
                print(f"{name}: grad vs fp64 reference agent {err:.2e}")
            else:
                worst = max(worst, err)
                assert err <= 2e-3, (name, err)
        print(f"worst non-sinc parameter-gradient error vs fp64 reference agent: {worst:.2e}")
    finally:
        perception.SincConv1d = reference_class
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev_tf32


def test_from_reference_on_an_existing_agent(perception):
    torch.manual_seed(3)
    pa = perception.PerceptionAgent()                                           # config.py defaults: 8 kHz
    with torch.no_grad():
        pa.sinc_conv.low_hz_.mul_(1.1)
    ref_layer = pa.sinc_conv
    pa.sinc_conv = sb.SincConv1d.from_reference(ref_layer)
    pa = pa.to(dev())
    x = torch.randn(2, 1, 4000, generator=torch.Generator().manual_seed(4))
    y_ref32 = ref_layer.cpu()(x).detach()
    y_ref64 = copy.deepcopy(ref_layer).double()(x.double()).detach()
    y = pa.sinc_conv(x.to(dev())).cpu()
    gap = max_rel(y_ref32, y_ref64)                  # the reference's own fp32 round-off (it synthesises in fp32)
    assert max_rel(y, y_ref64) <= 1e-5 and max_rel(y, y_ref32) <= 1e-5 + gap
    z_real, z_imag, sigma = pa(x.to(dev()))
    assert z_real.shape == (2, 256, 250) and sigma.shape == (2, 1, 250)


def test_cfg3_harness_small(perception):
    """examples/cfg3_harness.py end to end at a small size: 3a (the swap leaves the reference's training step
    bit-identical and launches none of our kernels) and 3b (PA -> pool -> CPEA -> MSA -> MSE step, loss equal to 1e-5
    between the reference layer and the B200 layer)."""
    import json
    import subprocess
    p = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "cfg3_harness.py"), "--batch", "4", "--seconds", "1",
                        "--steps", "2"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    rec = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["cfg3a"]["loss_identical"] and rec["cfg3a"]["sinc_kernel_launches_in_step"] == 0
    assert rec["cfg3b"]["loss_equal_to_1e-5"], rec["cfg3b"]
