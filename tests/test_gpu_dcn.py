"""-m gpu: the CUDA DCN boundary (through the C ABI: DCN shim -> rdfc_dcn_forward/backward) against the C oracle and
the golden vectors produced from the reference, plus the reference's own known-answer checks (deformconv/test.py)."""
import numpy as np
import pytest
import torch

from _synth import DCN_CASES, dcn_case_inputs

pytestmark = pytest.mark.gpu


def _geo(case):
    k, s, p, d, g, dg = (case[x] for x in ("k", "s", "p", "d", "g", "dg"))
    return (k, k, s, s, p, p, d, d, g, dg, 64)


@pytest.mark.parametrize("name", list(DCN_CASES))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forward_backward_vs_oracle_and_golden(name, dtype, golden_dir):
    from oracle import dcn as odcn
    from rdfc_gan_b200.dcn import DCN
    case = DCN_CASES[name]
    t = dcn_case_inputs(case, dtype=dtype)
    gold = np.load(f"{golden_dir}/dcn_{name}.npz")
    c = {k: torch.from_numpy(v).cuda() for k, v in t.items()}
    geo = _geo(case)
    if case["mask"]:
        out = DCN.modulated_deform_conv_forward(c["input"], c["weight"], c["bias"], c["offset"], c["mask"], *geo)
        grads = DCN.modulated_deform_conv_backward(c["input"], c["weight"], c["bias"], c["offset"], c["mask"],
                                                   c["grad_output"], *geo)
        names = ["grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"]
    else:
        out = DCN.deform_conv_forward(c["input"], c["weight"], c["bias"], c["offset"], *geo)
        grads = DCN.deform_conv_backward(c["input"], c["weight"], c["bias"], c["offset"], c["grad_output"], *geo)
        names = ["grad_input", "grad_offset", "grad_weight", "grad_bias"]
    assert out.is_contiguous() and out.dtype == c["input"].dtype
    ref_out = odcn.modulated_deform_conv_forward(t["input"], t["weight"], t["bias"], t["offset"], t.get("mask"), *geo)
    ref_g = odcn.modulated_deform_conv_backward(t["input"], t["weight"], t["bias"], t["offset"], t.get("mask"),
                                                t["grad_output"], *geo)
    ref_g = dict(zip(["grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"], ref_g))
    tol = 1e-10 if dtype == np.float64 else 2e-5      # fp32: accumulation-order noise of O(100)-term sums
    scale = lambda a: max(1.0, float(np.abs(a).max()))
    assert np.abs(out.cpu().numpy() - ref_out).max() <= tol * scale(ref_out)
    assert np.abs(out.cpu().numpy() - gold["output"]).max() <= tol * scale(ref_out)
    for n, gr in zip(names, grads):
        assert np.abs(gr.cpu().numpy() - ref_g[n]).max() <= tol * scale(ref_g[n]), n
        assert np.abs(gr.cpu().numpy() - gold[n]).max() <= tol * scale(ref_g[n]), n


def test_errors_like_reference():
    from rdfc_gan_b200.dcn import DCN
    x = torch.randn(2, 4, 6, 6, device="cuda")
    w = torch.randn(4, 4, 3, 3, device="cuda")
    b = torch.zeros(4, device="cuda")
    off = torch.zeros(2, 18, 6, 6, device="cuda")
    m = torch.ones(2, 9, 6, 6, device="cuda")
    with pytest.raises(RuntimeError, match="contiguous"):
        DCN.modulated_deform_conv_forward(x.transpose(2, 3), w, b, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        DCN.modulated_deform_conv_forward(x.cpu(), w.cpu(), b.cpu(), off.cpu(), m.cpu(), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    x3 = torch.randn(3, 4, 6, 6, device="cuda")
    with pytest.raises(RuntimeError, match="im2col_step"):     # batch 3 % min(3, 2) != 0
        DCN.modulated_deform_conv_forward(x3, w, b, torch.zeros(3, 18, 6, 6, device="cuda"),
                                          torch.ones(3, 9, 6, 6, device="cuda"), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2)
    with pytest.raises(RuntimeError, match="kernel"):
        DCN.modulated_deform_conv_forward(x, w, b, off, m, 5, 5, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    with pytest.raises(RuntimeError):
        DCN.deform_psroi_pooling_forward()


# ---- the reference's own checks (deformconv/test.py), same shapes and thresholds ------------------------------
N, inC, inH, inW, outC, kH, kW = 2, 4, 4, 4, 4, 3, 3


def _identity(weight, bias, groups):
    weight.data.zero_()
    bias.data.zero_()
    o, i, h, w = weight.shape
    oc = o // groups
    for p in range(i):
        for q in range(o):
            if p == q % oc:
                weight.data[q, p, h // 2, w // 2] = 1.0


def test_mdconv_and_dconv_zero_offset_equal_conv2d():      # test.py:36-110
    from rdfc_gan_b200.dcn import DeformConv, ModulatedDeformConv
    torch.manual_seed(3)
    x = torch.randn(N, inC, inH, inW).cuda()
    offset = torch.zeros(N, 2 * kH * kW, inH, inW).cuda()
    mask = 2 * torch.sigmoid(torch.zeros(N, kH * kW, inH, inW)).cuda()
    for cls in (ModulatedDeformConv, DeformConv):
        dcn = cls(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, groups=2, deformable_groups=1, im2col_step=1).cuda()
        pcn = torch.nn.Conv2d(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, groups=2).cuda()
        pcn.weight, pcn.bias = dcn.weight, dcn.bias
        out_d = dcn(x, offset, mask) if cls is ModulatedDeformConv else dcn(x, offset)
        assert (out_d - pcn(x)).abs().max() < 1e-5


def test_zero_offset_identity():                           # test.py:112-181
    from rdfc_gan_b200.dcn import DeformConv, ModulatedDeformConv
    x = torch.randn(N, inC, inH, inW).cuda()
    offset = torch.zeros(N, 2 * kH * kW, inH, inW).cuda()
    mask = torch.full((N, kH * kW, inH, inW), 0.5).cuda()
    d = ModulatedDeformConv(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, groups=2, deformable_groups=1, im2col_step=1).cuda()
    _identity(d.weight, d.bias, 2)
    assert (2 * d(x, offset, mask) - x).abs().max() < 1e-10
    d1 = DeformConv(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, groups=2, deformable_groups=1, im2col_step=1).cuda()
    _identity(d1.weight, d1.bias, 2)
    assert (d1(x, offset) - x).abs().max() < 1e-10


def test_im2col_step_invariance_forward_backward():        # test.py:183-349
    from rdfc_gan_b200.dcn import _DeformConv, _ModulatedDeformConv
    torch.manual_seed(0)
    x = torch.rand(N, inC, inH, inW).cuda() * 0.01
    off = (torch.randn(N, 2 * kH * kW, inH, inW).cuda() * 2)
    msk = torch.sigmoid(torch.randn(N, kH * kW, inH, inW).cuda())
    w = torch.randn(outC, inC // 2, kH, kW).cuda()
    b = torch.rand(outC).cuda()
    res = []
    for step in (1, 2):
        xs, os_, ms, ws, bs = (t.clone().requires_grad_(True) for t in (x, off, msk, w, b))
        out = _ModulatedDeformConv(xs, os_, ms, ws, bs, 1, 1, 1, 2, 1, step)
        out.sum().backward()
        res.append((out.detach(), xs.grad, os_.grad, ms.grad, ws.grad, bs.grad))
    assert (res[0][0] - res[1][0]).abs().max() < 1e-10
    # im2col_step is ignored here, so the two runs differ only by the fp32 atomicAdd order of the grad_input scatter
    # (the reference has the same atomics, cuh:249); offset / mask / weight / bias gradients are bit-identical.
    assert all(torch.equal(a, c) for a, c in zip(res[0][2:], res[1][2:]))
    assert (res[0][1] - res[1][1]).abs().max() < 1e-6
    res = []
    for step in (1, 2):
        xs, os_, ws, bs = (t.clone().requires_grad_(True) for t in (x, off, w, b))
        out = _DeformConv(xs, os_, ws, bs, 1, 1, 1, 2, 1, step)
        out.sum().backward()
        res.append((out.detach(), xs.grad, os_.grad, ws.grad, bs.grad))
    assert (res[0][0] - res[1][0]).abs().max() < 1e-10
    assert all(torch.equal(a, c) for a, c in zip(res[0][2:], res[1][2:]))
    assert (res[0][1] - res[1][1]).abs().max() < 1e-6


def test_gradcheck():                                      # test.py:375-434 (fp64 like check_gradient_dconv)
    from torch.autograd import gradcheck
    from rdfc_gan_b200.dcn import _DeformConv, _ModulatedDeformConv
    torch.manual_seed(1)
    x = (torch.rand(N, inC, inH, inW, dtype=torch.float64).cuda() * 0.01).requires_grad_(True)
    off = (torch.randn(N, 2 * kH * kW, inH, inW, dtype=torch.float64).cuda() * 2).requires_grad_(True)
    msk = torch.sigmoid(torch.rand(N, kH * kW, inH, inW, dtype=torch.float64).cuda()).requires_grad_(True)
    w = torch.randn(outC, inC // 2, kH, kW, dtype=torch.float64).cuda().requires_grad_(True)
    b = torch.rand(outC, dtype=torch.float64).cuda().requires_grad_(True)
    assert gradcheck(_ModulatedDeformConv, (x, off, msk, w, b, 1, 1, 1, 2, 1, 64), eps=1e-6, atol=1e-5, rtol=1e-4,
                     nondet_tol=1e-9)
    assert gradcheck(_DeformConv, (x, off, w, b, 1, 1, 1, 2, 1, 64), eps=1e-6, atol=1e-5, rtol=1e-4, nondet_tol=1e-9)


def test_example_pack_modules():                           # test.py:506-530
    from rdfc_gan_b200.dcn import DeformConvPack, ModulatedDeformConvPack
    x = torch.randn(2, 64, 32, 32).cuda()
    for cls in (ModulatedDeformConvPack, DeformConvPack):
        m = cls(64, 128, kernel_size=(3, 3), stride=1, padding=1, deformable_groups=2).cuda()
        out = m(x)
        assert out.shape == (2, 128, 32, 32)
        ref = torch.nn.functional.conv2d(x.double().cpu(), m.weight.double().cpu(), m.bias.double().cpu(), 1, 1)
        if cls is ModulatedDeformConvPack:
            ref = 0.5 * (ref - m.bias.double().cpu().view(1, -1, 1, 1)) + m.bias.double().cpu().view(1, -1, 1, 1)
        assert (out.double().cpu() - ref).abs().max() < 1e-3
        out.mean().backward()
        assert m.weight.grad is not None and torch.isfinite(m.weight.grad).all()


@pytest.mark.parametrize("case", [
    # B, Cin, Cout, H, W, k, s, p, d, groups, dg, mask      (first: the reference's timing shape, deformconv/test.py:519-530)
    (2, 64, 128, 128, 128, 3, 1, 1, 1, 1, 2, True), (2, 64, 64, 37, 45, 3, 2, 1, 1, 2, 4, True), (1, 32, 48, 20, 33, 3, 1, 2, 2, 1, 1, False),
    (3, 128, 256, 19, 26, 1, 1, 0, 1, 1, 2, True)])
def test_forward_on_tensor_cores_matches_strip_kernel(case):
    """GEMM-sized DCN layers ((Cin / groups) * kh * kw a multiple of 32, >= 16 outputs per group) take the tcgen05 path: sampled
    columns as split fp16 halves + one 1x1 implicit GEMM per group at fp32 fidelity.  It must agree with the CUDA-core strip
    kernel (RDFC_DCN_TC = 0, the path the golden / oracle tests above pin) to fp32 accumulation noise, and -- at the first,
    smaller-batch shape -- with the fp64 C oracle."""
    from oracle import dcn as odcn
    from rdfc_gan_b200 import _cabi as C
    from rdfc_gan_b200.dcn import DCN
    B, Cin, Cout, H, W, k, s, p, d, g, dg, with_mask = case
    gen = torch.Generator(device="cuda").manual_seed(sum(case[:11]))
    Ho, Wo = (H + 2 * p - (d * (k - 1) + 1)) // s + 1, (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = torch.randn(B, Cin, H, W, device="cuda", generator=gen)
    w = torch.randn(Cout, Cin // g, k, k, device="cuda", generator=gen) * 0.1
    b = torch.randn(Cout, device="cuda", generator=gen) * 0.1
    off = torch.randn(B, dg * 2 * k * k, Ho, Wo, device="cuda", generator=gen) * 2.0
    m = torch.rand(B, dg * k * k, Ho, Wo, device="cuda", generator=gen) if with_mask else None
    geo = (k, k, s, s, p, p, d, d, g, dg, 64)
    outs = []
    for tc in (1, 0):
        C.set_knob("RDFC_DCN_TC", tc)
        n0 = C.launch_count()
        outs.append(DCN.modulated_deform_conv_forward(x, w, b, off, m, *geo) if with_mask else DCN.deform_conv_forward(x, w, b, off, *geo))
        launches = C.launch_count() - n0
        assert (launches >= 3 + g) == bool(tc), (tc, launches)          # im2col + per-group (pack + GEMM [+ split]) + transpose vs one strip kernel
    C.set_knob("RDFC_DCN_TC", None)
    scale = float(outs[1].abs().max())
    assert outs[0].shape == (B, Cout, Ho, Wo) and outs[0].is_contiguous()
    assert float((outs[0] - outs[1]).abs().max()) <= 3e-5 * scale, float((outs[0] - outs[1]).abs().max()) / scale
    if H * W <= 2000:
        ref = odcn.modulated_deform_conv_forward(x.cpu().numpy().astype(np.float64), w.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64),
                                                 off.cpu().numpy().astype(np.float64), None if m is None else m.cpu().numpy().astype(np.float64), *geo)
        assert np.abs(outs[0].cpu().numpy() - ref).max() <= 3e-5 * scale


@pytest.mark.parametrize("case", [
    # B, Cin, Cout, H, W, k, s, p, d, groups, dg, mask, magnitude of the input / grad_output
    (2, 64, 128, 128, 128, 3, 1, 1, 1, 1, 2, True, 1.0), (2, 64, 64, 37, 45, 3, 2, 1, 1, 2, 4, True, 1.0), (1, 32, 64, 20, 33, 3, 1, 2, 2, 1, 1, False, 1.0),
    (3, 128, 256, 19, 26, 1, 1, 0, 1, 1, 2, True, 1.0), (2, 64, 64, 23, 31, 3, 1, 1, 1, 1, 2, True, 1e-4), (2, 64, 64, 23, 31, 3, 1, 1, 1, 1, 2, True, 3e3)])
def test_backward_on_tensor_cores_matches_cuda_core_kernels(case):
    """GEMM-sized DCN layers run both contractions of the backward on tcgen05 (columns' gradient = W^T . grad_output through conv_umma,
    grad_weight = grad_output . columns^T through the filter-gradient kernel, split fp16 operands scaled by a power of two of each tensor's
    maximum).  All five gradients must agree with the CUDA-core kernels (RDFC_DCN_TC = 0, the path the goldens / gradcheck pin) to fp32
    accumulation noise -- also for tensors far from unit magnitude (fp16's exponent range) -- and, at small sizes, with the fp64 C oracle."""
    from oracle import dcn as odcn
    from rdfc_gan_b200 import _cabi as C
    from rdfc_gan_b200.dcn import DCN
    B, Cin, Cout, H, W, k, s, p, d, g, dg, with_mask, mag = case
    gen = torch.Generator(device="cuda").manual_seed(sum(case[:11]) + 7)
    Ho, Wo = (H + 2 * p - (d * (k - 1) + 1)) // s + 1, (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = torch.randn(B, Cin, H, W, device="cuda", generator=gen) * mag
    w = torch.randn(Cout, Cin // g, k, k, device="cuda", generator=gen) * 0.1
    b = torch.randn(Cout, device="cuda", generator=gen) * 0.1
    off = torch.randn(B, dg * 2 * k * k, Ho, Wo, device="cuda", generator=gen) * 2.0
    m = torch.rand(B, dg * k * k, Ho, Wo, device="cuda", generator=gen) if with_mask else None
    go = torch.randn(B, Cout, Ho, Wo, device="cuda", generator=gen) * mag
    geo = (k, k, s, s, p, p, d, d, g, dg, 64)
    res = []
    for tc in (1, 0):
        C.set_knob("RDFC_DCN_TC", tc)
        n0 = C.launch_count()
        res.append(DCN.modulated_deform_conv_backward(x, w, b, off, m, go, *geo) if with_mask else DCN.deform_conv_backward(x, w, b, off, go, *geo))
        launches = C.launch_count() - n0
        assert (launches >= 8) == bool(tc), (tc, launches)
        if tc:                                   # forward at this magnitude too
            y_tc = DCN.modulated_deform_conv_forward(x, w, b, off, m, *geo) if with_mask else DCN.deform_conv_forward(x, w, b, off, *geo)
        else:
            y_cc = DCN.modulated_deform_conv_forward(x, w, b, off, m, *geo) if with_mask else DCN.deform_conv_forward(x, w, b, off, *geo)
    C.set_knob("RDFC_DCN_TC", None)
    names = ("input", "offset", "mask", "weight", "bias") if with_mask else ("input", "offset", "weight", "bias")
    assert float((y_tc - y_cc).abs().max()) <= 3e-5 * float(y_cc.abs().max())
    for name, a, r in zip(names, res[0], res[1]):
        assert a.shape == r.shape and a.is_contiguous()
        scale = float(r.abs().max())
        tol = 2e-4 if name == "input" else 5e-5              # grad_input sums atomics in a different order
        assert float((a - r).abs().max()) <= tol * scale, (name, float((a - r).abs().max()) / scale)
    if H * W <= 1000:
        f64 = lambda t: None if t is None else t.cpu().numpy().astype(np.float64)
        ref = odcn.modulated_deform_conv_backward(f64(x), f64(w), f64(b), f64(off), f64(m), f64(go), *geo)
        ref = [r for r in ref if r is not None]
        for name, a, r in zip(names, res[0], ref):
            scale = float(np.abs(r).max())
            assert np.abs(a.cpu().numpy() - r).max() <= 1e-4 * scale, (name, np.abs(a.cpu().numpy() - r).max() / scale)
