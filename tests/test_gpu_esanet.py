"""-m gpu: the ESANet guidance network (rdfc_gan_b200.esanet) against the outputs of the reference's ESANetOneModality on the same
synthetic weights and inputs (tests/golden/esanet_*.npz), and inside RDF-GAN's DCVGANGenerator as its global_guidance_module."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["r18_small", "r34_full"])
def test_esanet_against_reference(name, golden_dir):
    """bf16 tensor-core arithmetic against the fp32 reference: the stated bound is relative to the RMS of the reference's logits
    (RMSE <= 2 %, max-abs <= 25 %: twice the measured 0.8 % / 12 %)."""
    from make_esanet_golden import ESANET_CASES, esanet_input, esanet_weights
    from _synth import state_dict_digest
    from rdfc_gan_b200.esanet import ESANetOneModality
    c = ESANET_CASES[name]
    gold = np.load(f"{golden_dir}/esanet_{name}.npz")
    net = ESANetOneModality(**c["kw"]).eval()
    sd = esanet_weights(net, c["seed"])
    assert state_dict_digest(sd) == int(gold["digest"][0])
    net.load_state_dict(sd)
    net = net.cuda()
    x = esanet_input(c).cuda()
    y = net(x)
    assert tuple(y.shape) == (c["B"], 40, c["H"] // 4 * 4 if c["H"] % 4 == 0 else y.shape[2], y.shape[3])
    got = y.cpu().numpy()[:, :, ::c["stride"], ::c["stride"]]
    assert got.shape == gold["logits"].shape
    rms = float(gold["rms"])
    d = got - gold["logits"]
    rel_rmse, rel_max = float(np.sqrt(np.mean(d ** 2))) / rms, float(np.abs(d).max()) / rms
    import json, os
    if os.environ.get("RDFC_DUMP_PARITY"):
        with open(os.environ["RDFC_DUMP_PARITY"], "a") as f:
            f.write(json.dumps({"case": f"esanet:{name}", "errs": {"rel_rmse": rel_rmse, "rel_max": rel_max, "rms": rms}}) + "\n")
    assert rel_rmse <= 2e-2 and rel_max <= 2.5e-1, (rel_rmse, rel_max)
    y2 = net(x)                                    # graph replay
    assert torch.equal(y, y2)
    with torch.no_grad():                          # an in-place weight change is picked up
        net.decoder.conv_out.bias.add_(1.0)
    y3 = net(x)
    assert float((y3 - y).abs().max()) > 1e-2              # (the two learned up-samplings after conv_out carry synthetic filters too)


def test_esanet_inside_dcvgan_generator():
    """config 2 end to end: DCVGANGenerator(global_guidance_module=ESANet) -- the guidance logits feed the 40-channel stems."""
    from make_esanet_golden import ESANET_CASES, esanet_weights
    from _synth import synth_inputs, synth_state_dict
    from rdfc_gan_b200.esanet import ESANetOneModality
    from rdfc_gan_b200.generator import DCVGANGenerator
    nl = dict(prop_kernel=3, prop_time=6, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
    esa = ESANetOneModality(**dict(ESANET_CASES["r18_small"]["kw"], height=64, width=96)).eval()
    esa.load_state_dict(esanet_weights(esa, 3))
    G = DCVGANGenerator(esa, pretrained_on_imagenet=False, semantic_channels_in=40, use_nlpsn_refine=True, nlspn_configs=nl).eval()
    sd = synth_state_dict(G, seed=4, recipe="scaled", nlspn_stress=True)
    for k in list(sd):                             # the guidance net keeps the weights loaded above
        if k.startswith("global_guidance_module."):
            sd[k] = G.state_dict()[k]
    G.load_state_dict(sd)
    G = G.cuda().set_precision("bf16")
    rgb, _, depth = synth_inputs(2, 64, 96, seed=4)
    with torch.no_grad():
        outs = G(rgb.cuda(), depth.cuda())
        guidance = esa(rgb.cuda())
        G2 = DCVGANGenerator(torch.nn.Identity(), pretrained_on_imagenet=False, semantic_channels_in=40, use_nlpsn_refine=True, nlspn_configs=nl).eval()
        G2.load_state_dict({k: v for k, v in G.state_dict().items() if not k.startswith("global_guidance_module.")})
        ref = G2.cuda().set_precision("bf16")(guidance, depth.cuda())
    assert len(outs) == 5 and all(torch.equal(a, b) for a, b in zip(outs, ref))
    # the pipelined host API runs the guidance network inside the stream: three host batches, same numbers as the module call
    keys = ("depth_map_1", "confidence_map_1", "depth_map_2", "confidence_map_2", "pred_depth")
    rgb2, _, depth2 = synth_inputs(2, 64, 96, seed=5)
    batches = [(rgb.pin_memory(), depth.pin_memory()), (rgb2.pin_memory(), depth2.pin_memory()), (rgb.pin_memory(), depth.pin_memory())]
    got = [{k: o[k].clone() for k in keys} for o in G.stream(iter(batches), outputs=keys)]
    assert len(got) == 3
    with torch.no_grad():
        outs2 = G(rgb2.cuda(), depth2.cuda())
    for i, want in enumerate((outs, outs2, outs)):
        for k, w in zip(keys, want):
            assert torch.equal(got[i][k], w.cpu()), (i, k)


@pytest.mark.parametrize("case", [(2, 3, 228, 304), (1, 3, 37, 50), (3, 1, 64, 33), (2, 4, 17, 129)])
def test_first_conv_kernels(case):
    """encoder.conv1 + bn1 + ReLU (7x7, stride 2, pad 3, fp32 NCHW -> bf16 NHWC): the register-tiled kernel and the plain one
    (RDFC_FIRSTCONV_FAST = 0) against F.conv2d in fp64, odd sizes and every input-channel count."""
    import ctypes
    import torch.nn.functional as F
    from rdfc_gan_b200 import _cabi as C
    B, Cin, H, W = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(64, Cin, 7, 7, device="cuda", generator=g) / (7 * Cin ** 0.5)
    sc, sh = torch.rand(64, device="cuda", generator=g) + 0.5, torch.randn(64, device="cuda", generator=g) * 0.1
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    want = torch.relu(F.conv2d(x.double(), w.double(), None, 2, 3) * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    outs = []
    try:
        for fast in (1, 0):
            if not fast and Cin == 4:            # the plain kernel keeps only the filter bank in shared memory and stops at 48 KB (Cin <= 3)
                continue
            C.set_knob("RDFC_FIRSTCONV_FAST", fast)
            out = torch.full((B, Ho, Wo, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
            v = C.view(out)
            C.check(C.lib.rdfc_first_conv_forward(C.ptr(x), B, Cin, H, W, C.ptr(w), 7, 2, 3, C.ptr(sc), C.ptr(sh), 1, ctypes.byref(v), C.stream_ptr()))
            torch.cuda.synchronize()
            outs.append(out)
            err = (out.double().permute(0, 3, 1, 2) - want).abs().max()
            assert float(err) <= 8e-3 * max(1.0, float(want.abs().max())), (fast, float(err))          # one bf16 rounding of the output
    finally:
        C.set_knob("RDFC_FIRSTCONV_FAST", None)
    if len(outs) == 2:
        assert float((outs[0].float() - outs[1].float()).abs().max()) <= 8e-3 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("case", [(2, 40, 114, 152, 228, 304, False, True), (2, 40, 29, 38, 57, 76, False, True), (3, 128, 29, 38, 57, 76, True, False),
                                  (1, 64, 15, 19, 29, 38, True, False), (2, 40, 20, 31, 64, 65, False, False)])
def test_learned_upsampling_kernels(case):
    """The decoder's learned up-sampling (nearest resize + depth-wise 3x3 + bias + skip): the restructured kernels (shared-memory staged
    NCHW variant, 8-channel NHWC variant) give the same numbers as the plain kernel (RDFC_UPSAMPLE_FAST = 0), and both match PyTorch."""
    import ctypes
    import torch.nn.functional as F
    from rdfc_gan_b200 import _cabi as C
    B, Cc, Hi, Wi, Ho, Wo, with_skip, nchw = case
    g = torch.Generator(device="cuda").manual_seed(sum(case[:6]))
    x = torch.randn(B, Hi, Wi, Cc, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cc, 1, 3, 3, device="cuda", generator=g) * 0.3
    bias = torch.randn(Cc, device="cuda", generator=g) * 0.1
    skip = torch.randn(B, Ho, Wo, Cc, device="cuda", generator=g).to(torch.bfloat16) if with_skip else None
    up = F.interpolate(x.float().permute(0, 3, 1, 2), size=(Ho, Wo), mode="nearest")
    want = F.conv2d(up, w, bias, padding=1, groups=Cc)
    if with_skip:
        want = want + skip.float().permute(0, 3, 1, 2)
    outs = []
    try:
        for fast in (1, 0):
            C.set_knob("RDFC_UPSAMPLE_FAST", fast)
            vx, vs = C.view(x), C.view(skip)
            if nchw:
                out = torch.full((B, Cc, Ho, Wo), float("nan"), device="cuda")
                C.check(C.lib.rdfc_upsample_dw_forward(ctypes.byref(vx), C.ptr(w), C.ptr(bias), ctypes.byref(vs) if with_skip else None, None, C.ptr(out),
                                                       B, Hi, Wi, Ho, Wo, C.stream_ptr()))
                got = out
            else:
                out = torch.full((B, Ho, Wo, Cc), float("nan"), device="cuda", dtype=torch.bfloat16)
                vo = C.view(out)
                C.check(C.lib.rdfc_upsample_dw_forward(ctypes.byref(vx), C.ptr(w), C.ptr(bias), ctypes.byref(vs) if with_skip else None, ctypes.byref(vo), None,
                                                       B, Hi, Wi, Ho, Wo, C.stream_ptr()))
                got = out.float().permute(0, 3, 1, 2)
            torch.cuda.synchronize()
            outs.append(got.clone())
            tol = 1e-4 if nchw else 2e-2
            assert float((got - want).abs().max()) <= tol * max(1.0, float(want.abs().max())), (fast, float((got - want).abs().max()))
    finally:
        C.set_knob("RDFC_UPSAMPLE_FAST", None)
    assert torch.equal(outs[0], outs[1]), "same arithmetic order: the restructured kernels must reproduce the plain one bit for bit"
