"""-m gpu: fused NLSPN kernels (rdfc_nlspn_affinity_forward / rdfc_nlspn_propagate_forward through the NLPSN /
NLSPNRefineModule drop-ins) against the golden vectors of the reference's nlspn_model.py and the numpy/C oracle."""
import numpy as np
import pytest
import torch

from _synth import nlspn_stress_inputs
from make_golden import NLSPN_CASES, NLSPN_SHAPE

pytestmark = pytest.mark.gpu
TOL = 1e-4          # BASELINE.json: fp32 max-abs <= 1e-4 (measured ~1e-6)


def _module(cfg, x):
    from rdfc_gan_b200.nlspn import NLSPNRefineModule
    mod = NLSPNRefineModule(prop_kernel=3, prop_time=cfg["prop_time"], affinity=cfg["affinity"], affinity_gamma=0.5,
                            conf_prop=cfg["conf_prop"], preserve_input=cfg["preserve_input"]).cuda().eval()
    mod.prop_layer.conv_offset_aff.weight.data.copy_(torch.from_numpy(x["conv_w"]))
    mod.prop_layer.conv_offset_aff.bias.data.copy_(torch.from_numpy(x["conv_b"]))
    return mod


@pytest.mark.parametrize("name", list(NLSPN_CASES))
def test_fused_vs_golden(name, golden_dir):
    cfg = NLSPN_CASES[name]
    B, H, W = NLSPN_SHAPE
    x = nlspn_stress_inputs(B, H, W, cfg["seed"])
    gold = np.load(f"{golden_dir}/nlspn_{name}.npz")
    mod = _module(cfg, x)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    mod.prop_layer.return_intermediates = True
    with torch.no_grad():
        y, inter, offset, aff, scale = mod.prop_layer(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
        y2, conf = mod(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
    assert len(inter) == cfg["prop_time"] and conf is t["confidence"]
    assert torch.equal(y, y2)
    assert np.abs(offset.cpu().numpy() - gold["offset"]).max() <= 1e-5
    assert np.abs(aff.cpu().numpy() - gold["aff"]).max() <= 1e-5
    assert np.abs(inter[0].cpu().numpy() - gold["first"]).max() <= TOL
    assert np.abs(y.cpu().numpy() - gold["y"]).max() <= TOL
    assert float(scale) == float(gold["aff_scale"][0])


@pytest.mark.parametrize("name", ["tgass18", "as12_preserve"])
def test_grad_mode_forward_matches_inference(name):
    """With gradients enabled the module runs the two differentiable fused ops; same numbers as the inference path, and
    list_feat (return_intermediates) carries gradients too."""
    cfg = NLSPN_CASES[name]
    B, H, W = NLSPN_SHAPE
    x = nlspn_stress_inputs(B, H, W, cfg["seed"])
    mod = _module(cfg, x)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    with torch.no_grad():
        y_fused, _ = mod(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
    g = t["guidance"].clone().requires_grad_(True)
    y_grad, _ = mod(t["pred_init"], g, t["confidence"], t["feat_fix"])
    assert (y_fused - y_grad).abs().max() <= 1e-6
    y_grad.sum().backward()
    assert torch.isfinite(g.grad).all() and g.grad.abs().sum() > 0
    assert mod.prop_layer.conv_offset_aff.weight.grad is not None
    # list_feat with gradients: a loss on an intermediate iteration against the composition on the general DCN kernels
    import _composition as comp
    mod.prop_layer.return_intermediates = True
    grads = []
    for fused in (True, False):
        p = t["pred_init"].clone().requires_grad_(True)
        g = t["guidance"].clone().requires_grad_(True)
        if fused:
            y, steps, _, _, _ = mod.prop_layer(p, g, t["confidence"], t["feat_fix"])
        else:
            y, steps = comp.refine(mod, p, g, t["confidence"], t["feat_fix"])
        assert len(steps) == cfg["prop_time"]
        (y.sum() + 0.5 * steps[2].square().sum() - steps[-1].sum()).backward()
        grads.append((p.grad.clone(), g.grad.clone()))
    for a, b in zip(*grads):
        assert (a - b).abs().max().item() <= 2e-4 * max(1.0, b.abs().max().item())


def test_unsupported_configurations_raise():
    """The fused kernels hard-code 8 guidance channels and a 3x3 propagation kernel: anything else must raise, not misread."""
    from rdfc_gan_b200.nlspn import NLPSN
    pl = NLPSN(channels_g=24, channels_f=1, k_g=3, k_f=5, prop_time=2, affinity="TGASS").cuda()
    with pytest.raises(NotImplementedError):
        pl(torch.zeros(1, 1, 16, 16, device="cuda"), torch.zeros(1, 24, 16, 16, device="cuda"), torch.ones(1, 1, 16, 16, device="cuda"))
    pl = NLPSN(channels_g=8, channels_f=1, k_g=3, k_f=3, prop_time=2, affinity="TGASS").cuda()
    with pytest.raises(RuntimeError):
        pl(torch.zeros(1, 1, 16, 16, device="cuda").double(), torch.zeros(1, 8, 16, 16, device="cuda").double(),
           torch.ones(1, 1, 16, 16, device="cuda").double())
    with pytest.raises(RuntimeError):
        pl(torch.zeros(1, 1, 16, 16), torch.zeros(1, 8, 16, 16), torch.ones(1, 1, 16, 16))


def test_oracle_at_odd_sizes():
    from oracle import nlspn as onl
    from rdfc_gan_b200.nlspn import NLSPNRefineModule
    for (B, H, W, seed) in ((1, 17, 33, 5), (3, 45, 31, 6)):
        x = nlspn_stress_inputs(B, H, W, seed)
        cfg = dict(prop_time=7, affinity="TGASS", conf_prop=True, preserve_input=False)
        mod = _module(cfg, x)
        t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
        with torch.no_grad():
            y, _ = mod(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
        ref, _, _ = onl.nlspn_forward(x["pred_init"], x["guidance"], x["confidence"], x["feat_fix"], x["conv_w"], x["conv_b"],
                                      np.array([4.0], np.float32), prop_time=7)
        assert np.abs(y.cpu().numpy() - ref).max() <= TOL


def test_full_size_properties():
    """Size-independent properties at the benchmark shape (B=8 here, 228x304, 18 iterations):
    (1) a constant map is a fixed point when the affinities sum to one and no tap leaves the image,
    (2) the propagation is linear in the feature map."""
    from rdfc_gan_b200 import _cabi as C
    B, H, W, T = 8, 228, 304, 18
    g = torch.Generator(device="cuda").manual_seed(0)
    off = torch.zeros(B, 18, H, W, device="cuda")
    off[:, :, 8:-8, 8:-8] = 3 * torch.randn(B, 18, H - 16, W - 16, device="cuda", generator=g).clamp(-2, 2)
    off[:, 8:10] = 0
    aff = torch.rand(B, 9, H, W, device="cuda", generator=g)
    aff = aff / aff.sum(1, keepdim=True)

    def prop(f):
        out, scratch = torch.empty_like(f), torch.empty_like(f)
        C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(f), C.ptr(off), C.ptr(aff), None, 0, C.ptr(out), C.ptr(scratch),
                                                   None, B, H, W, T, 0, None, C.stream_ptr()))
        return out
    const = torch.full((B, 1, H, W), 0.37, device="cuda")
    # taps that leave the image contribute zero, so only pixels farther than 18 x (1 + 2) px from the border qualify
    assert (prop(const)[:, :, 60:-60, 60:-60] - 0.37).abs().max() <= 2e-5
    a, b = torch.randn(B, 1, H, W, device="cuda", generator=g), torch.randn(B, 1, H, W, device="cuda", generator=g)
    assert (prop(2 * a - 3 * b) - (2 * prop(a) - 3 * prop(b))).abs().max() <= 1e-4


def _prop_inputs(B, H, W, seed, preserve):
    """offsets / affinities of the stress generator's own affinity stage + a random upstream gradient"""
    from oracle import nlspn as onl
    x = nlspn_stress_inputs(B, H, W, seed)
    off, aff = onl.get_offset_affinity(x["guidance"], x["confidence"], x["conv_w"], x["conv_b"], np.array([4.0], np.float32))[:2]
    rng = np.random.default_rng(seed + 100)
    gout = rng.standard_normal((B, 1, H, W)).astype(np.float32)
    return x, np.ascontiguousarray(off, np.float32), np.ascontiguousarray(aff, np.float32), gout


@pytest.mark.parametrize("preserve", [False, True])
def test_fused_backward_vs_oracle(preserve):
    """rdfc_nlspn_propagate_backward (through the autograd Function) against the CPU restatement of the reference's
    reverse loop (oracle.nlspn.nlspn_propagate_backward: one DCN backward per iteration)."""
    from oracle import nlspn as onl
    from rdfc_gan_b200.nlspn import _PropagateFused
    B, H, W, T = 2, 19, 27, 5
    x, off, aff, gout = _prop_inputs(B, H, W, 11, preserve)
    ref = onl.nlspn_propagate_backward(gout, x["pred_init"], off, aff, x["feat_fix"], preserve, T)
    f = torch.from_numpy(x["pred_init"]).cuda().requires_grad_(True)
    o = torch.from_numpy(off).cuda().requires_grad_(True)
    a = torch.from_numpy(aff).cuda().requires_grad_(True)
    fix = torch.from_numpy(x["feat_fix"]).cuda()
    y, _ = _PropagateFused.apply(f, o, a, fix if preserve else None, T, preserve)
    y.backward(torch.from_numpy(gout).cuda())
    for got, want, name in ((f.grad, ref[0], "feat_init"), (o.grad, ref[1], "offset"), (a.grad, ref[2], "aff")):
        want = torch.from_numpy(want)
        err = (got.cpu() - want).abs().max().item()
        assert err <= 1e-4 * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())


@pytest.mark.parametrize("name", ["tgass18", "as12_preserve", "tc12_noconf", "ass18"])
def test_fused_backward_matches_composition(name):
    """Training through NLSPNRefineModule: the fused forward/backward pair gives the gradients of the reference's composition
    (prop_time DCN Function calls on the general DCN kernels, tests/_composition.py) for the guidance, the initial depth and conv_offset_aff."""
    cfg = NLSPN_CASES[name]
    B, H, W = NLSPN_SHAPE
    x = nlspn_stress_inputs(B, H, W, cfg["seed"])
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    gout = torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    grads = {}
    import _composition as comp
    for fused in (True, False):
        mod = _module(cfg, x).train()
        g = t["guidance"].clone().requires_grad_(True)
        p = t["pred_init"].clone().requires_grad_(True)
        cf = t["confidence"].clone().requires_grad_(True)
        y, _ = mod(p, g, cf, t["feat_fix"]) if fused else comp.refine(mod, p, g, cf, t["feat_fix"])
        y.backward(gout)
        grads[fused] = (y.detach(), g.grad, p.grad, mod.prop_layer.conv_offset_aff.weight.grad, mod.prop_layer.conv_offset_aff.bias.grad,
                        mod.prop_layer.aff_scale_const.grad, cf.grad if cfg["conf_prop"] else None)
    assert grads[True][5] is not None or cfg["affinity"] != "TGASS"
    for a, b, nm in zip(grads[True], grads[False], ("y", "d guidance", "d pred_init", "d conv w", "d conv b", "d aff_scale", "d confidence")):
        if a is None and b is None:
            continue
        scale = max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= 2e-4 * scale, (nm, (a - b).abs().max().item(), scale)


def test_fused_backward_full_size_linearity():
    """At the bench size the oracle is too slow: the backward is linear in grad_out, and <A x, g> = <x, A^T g>."""
    from rdfc_gan_b200.nlspn import _PropagateFused
    B, H, W, T = 4, 228, 304, 18
    x, off, aff, _ = _prop_inputs(B, 32, 40, 21, False)
    rep = lambda v: torch.from_numpy(v).cuda().repeat(2, 1, 8, 8)[:B, :, :H, :W].contiguous()
    o, a = rep(off), rep(aff)
    gen = torch.Generator(device="cuda").manual_seed(5)
    f = torch.randn(B, 1, H, W, device="cuda", generator=gen, dtype=torch.float64).float().requires_grad_(True)
    g1 = torch.randn(B, 1, H, W, device="cuda", generator=gen)
    y, _ = _PropagateFused.apply(f, o, a, None, T, False)
    (gf,) = torch.autograd.grad(y, f, g1)
    lhs = (y.double() * g1.double()).sum().item()          # <A f, g>
    rhs = (f.detach().double() * gf.double()).sum().item()  # <f, A^T g>
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs)), (lhs, rhs)


@pytest.mark.parametrize("halo", [None, 0, 2, 12])
def test_large_offsets_leave_the_staged_band(halo):
    """Offsets far larger than the band kernel's halo (sigma = 7 px, some beyond the image): taps that leave the staged band
    take the global-memory path; every halo setting (RDFC_NLSPN_HALO) must give the oracle's numbers."""
    from oracle import dcn as odcn
    from rdfc_gan_b200 import _cabi as C
    C.set_knob("RDFC_NLSPN_HALO", halo)
    B, H, W, T = 3, 70, 92, 4
    rng = np.random.default_rng(17)
    off = (7.0 * rng.standard_normal((B, 18, H, W))).astype(np.float32)
    off[:, 8:10] = 0
    aff = rng.uniform(-1, 1, (B, 9, H, W)).astype(np.float32)
    aff /= np.abs(aff).sum(1, keepdims=True)
    f = rng.standard_normal((B, 1, H, W)).astype(np.float32)
    ref = odcn.nlspn_propagate(f, off, aff, None, False, 3, T)
    tf, to, ta = (torch.from_numpy(v).cuda() for v in (f, off, aff))
    out, scratch = torch.empty_like(tf), torch.empty_like(tf)
    C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(tf), C.ptr(to), C.ptr(ta), None, 0, C.ptr(out), C.ptr(scratch), None,
                                               B, H, W, T, 0, None, C.stream_ptr()))
    C.set_knob("RDFC_NLSPN_HALO", None)
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL


def test_full_size_against_oracle():
    """One 228x304 image, 18 iterations, stress offsets (sigma ~ 2.2 px): the CUDA path against the numpy / C oracle at the
    benchmark size (the oracle's OpenMP loop takes seconds for one image), fp32 planes, through the module."""
    from oracle import nlspn as onl
    B, H, W = 1, 228, 304
    x = nlspn_stress_inputs(B, H, W, 31)
    cfg = dict(prop_time=18, affinity="TGASS", conf_prop=True, preserve_input=False)
    mod = _module(cfg, x)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    with torch.no_grad():
        y, _ = mod(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
    ref, _, _ = onl.nlspn_forward(x["pred_init"], x["guidance"], x["confidence"], x["feat_fix"], x["conv_w"], x["conv_b"],
                                  np.array([4.0], np.float32), prop_time=18)
    assert np.abs(y.cpu().numpy() - ref).max() <= TOL


def _packed_to_planes(buf, B, H, W):
    """Decode the fp16 stream of rdfc_nlspn_affinity_forward_packed into the reference's fp32 offset / aff planes."""
    n = B * H * W
    nb = (n + 31) // 32
    h = buf.view(torch.float16).reshape(nb, 3, 32, 8).permute(0, 2, 1, 3).reshape(nb * 32, 24)[:n].float()     # [pixel][24]
    h = h.reshape(B, H, W, 24).permute(0, 3, 1, 2)
    off8, a8 = h[:, :16], h[:, 16:]
    zero = torch.zeros_like(off8[:, :2])
    offset = torch.cat([off8[:, :8], zero, off8[:, 8:]], 1)
    centre = 1.0 - (((a8[:, 0] + a8[:, 1]) + (a8[:, 2] + a8[:, 3])) + ((a8[:, 4] + a8[:, 5]) + (a8[:, 6] + a8[:, 7])))
    aff = torch.cat([a8[:, :4], centre[:, None], a8[:, 4:]], 1)
    return offset.contiguous(), aff.contiguous()


@pytest.mark.parametrize("shape,preserve,sub", [((2, 24, 32), False, None), ((3, 45, 31), True, None), ((2, 70, 92), False, 5),
                                                ((1, 228, 304), False, None)])
def test_packed_stream_path(shape, preserve, sub):
    """bf16-mode NLSPN: the affinity stage writes the packed fp16 stream, the packed propagation kernel consumes it and
    applies the output fusion in its last iteration.  (1) the stream decodes to the fp16 rounding of the fp32 planes;
    (2) the propagation is the oracle's arithmetic on the decoded (rounded) planes to fp32 tolerance; (3) the fused epilogue
    equals clamp + rdfc_fuse_depth_forward; (4) sub-band walking (RDFC_NLSPN_SUB) changes nothing."""
    from oracle import dcn as odcn
    from rdfc_gan_b200 import _cabi as C
    B, H, W = shape
    T = 6
    x = nlspn_stress_inputs(B, H, W, 41)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    scale = torch.tensor([4.0], device="cuda")
    off32 = torch.empty(B, 18, H, W, device="cuda")
    aff32 = torch.empty(B, 9, H, W, device="cuda")
    C.check(C.lib.rdfc_nlspn_affinity_forward(C.ptr(t["guidance"]), C.ptr(t["confidence"]), C.ptr(t["conv_w"]), C.ptr(t["conv_b"]),
                                              C.ptr(scale), C.AFFINITY["TGASS"], 1, C.ptr(off32), C.ptr(aff32), B, H, W, C.stream_ptr()))
    packed = torch.zeros(C.lib.rdfc_nlspn_packed_bytes(B, H, W), dtype=torch.uint8, device="cuda")
    C.check(C.lib.rdfc_nlspn_affinity_forward_packed(C.ptr(t["guidance"]), C.ptr(t["confidence"]), C.ptr(t["conv_w"]), C.ptr(t["conv_b"]),
                                                     C.ptr(scale), C.AFFINITY["TGASS"], 1, C.ptr(packed), B, H, W, C.stream_ptr()))
    off16, aff16 = _packed_to_planes(packed, B, H, W)
    assert torch.equal(off16, off32.half().float())
    nc = [0, 1, 2, 3, 5, 6, 7, 8]
    assert torch.equal(aff16[:, nc], aff32[:, nc].half().float())
    fix = t["feat_fix"] if preserve else None
    ref = odcn.nlspn_propagate(x["pred_init"], off16.cpu().numpy(), aff16.cpu().numpy(), x["feat_fix"] if preserve else None,
                               preserve, 3, T)
    d1 = torch.rand(B, 1, H, W, device="cuda") * 2 - 1
    c1, c2 = torch.rand(B, 1, H, W, device="cuda"), torch.rand(B, 1, H, W, device="cuda")
    outs = []
    for s in ([None, sub] if sub else [None]):
        C.set_knob("RDFC_NLSPN_SUB", s)
        out, scratch, pred = torch.empty_like(d1), torch.empty_like(d1), torch.empty_like(d1)
        fz = C.FuseOut(d1.data_ptr(), c1.data_ptr(), c2.data_ptr(), pred.data_ptr())
        import ctypes
        C.check(C.lib.rdfc_nlspn_propagate_forward_packed(C.ptr(t["pred_init"]), C.ptr(packed), C.ptr(fix), int(preserve), C.ptr(out),
                                                          C.ptr(scratch), B, H, W, T, 0, ctypes.byref(fz), C.stream_ptr()))
        outs.append((out, pred))
    C.set_knob("RDFC_NLSPN_SUB", None)
    out, pred = outs[0]
    assert np.abs(out.cpu().numpy() - np.clip(ref, -1, 1)).max() <= TOL
    d2c, pred2 = torch.empty_like(d1), torch.empty_like(d1)
    C.check(C.lib.rdfc_fuse_depth_forward(C.ptr(d1), C.ptr(c1), C.ptr(torch.from_numpy(ref).cuda()), C.ptr(c2), C.ptr(d2c), C.ptr(pred2),
                                          B * H * W, C.stream_ptr()))
    assert (pred - pred2).abs().max().item() <= TOL
    for o2, p2 in outs[1:]:
        assert torch.equal(o2, out) and torch.equal(p2, pred)
    # the fp32-plane kernel with the same fused epilogue and the same (decoded) planes gives the same numbers
    out3, pred3 = torch.empty_like(d1), torch.empty_like(d1)
    fz3 = C.FuseOut(d1.data_ptr(), c1.data_ptr(), c2.data_ptr(), pred3.data_ptr())
    C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(t["pred_init"]), C.ptr(off16), C.ptr(aff16), C.ptr(fix), int(preserve), C.ptr(out3),
                                               C.ptr(scratch), None, B, H, W, T, 0, ctypes.byref(fz3), C.stream_ptr()))
    assert (out3 - out).abs().max().item() <= 1e-5 and (pred3 - pred).abs().max().item() <= 1e-5
