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
def test_autograd_composition_matches_fused(name):
    """With gradients enabled the module runs the reference's composition on the general DCN kernels; same numbers."""
    cfg = NLSPN_CASES[name]
    B, H, W = NLSPN_SHAPE
    x = nlspn_stress_inputs(B, H, W, cfg["seed"])
    mod = _module(cfg, x)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    with torch.no_grad():
        y_fused, _ = mod(t["pred_init"], t["guidance"], t["confidence"], t["feat_fix"])
    g = t["guidance"].clone().requires_grad_(True)
    y_comp, _ = mod(t["pred_init"], g, t["confidence"], t["feat_fix"])
    assert (y_fused - y_comp).abs().max() <= 1e-5
    y_comp.sum().backward()
    assert torch.isfinite(g.grad).all() and g.grad.abs().sum() > 0
    assert mod.prop_layer.conv_offset_aff.weight.grad is not None


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
                                                   None, B, H, W, T, 0, C.stream_ptr()))
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
    y = _PropagateFused.apply(f, o, a, fix if preserve else None, T, preserve)
    y.backward(torch.from_numpy(gout).cuda())
    for got, want, name in ((f.grad, ref[0], "feat_init"), (o.grad, ref[1], "offset"), (a.grad, ref[2], "aff")):
        want = torch.from_numpy(want)
        err = (got.cpu() - want).abs().max().item()
        assert err <= 1e-4 * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())


@pytest.mark.parametrize("name", ["tgass18", "as12_preserve", "tc12_noconf", "ass18"])
def test_fused_backward_matches_composition(name):
    """Training through NLSPNRefineModule: the fused forward/backward pair gives the gradients of the reference's composition
    (prop_time DCN Function calls) for the guidance, the initial depth and conv_offset_aff."""
    cfg = NLSPN_CASES[name]
    B, H, W = NLSPN_SHAPE
    x = nlspn_stress_inputs(B, H, W, cfg["seed"])
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    gout = torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    grads = {}
    for fused in (True, False):
        mod = _module(cfg, x).train()
        mod.prop_layer.fused_backward = fused
        g = t["guidance"].clone().requires_grad_(True)
        p = t["pred_init"].clone().requires_grad_(True)
        cf = t["confidence"].clone().requires_grad_(True)
        y, _ = mod(p, g, cf, t["feat_fix"])
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
    y = _PropagateFused.apply(f, o, a, None, T, False)
    (gf,) = torch.autograd.grad(y, f, g1)
    lhs = (y.double() * g1.double()).sum().item()          # <A f, g>
    rhs = (f.detach().double() * gf.double()).sum().item()  # <f, A^T g>
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs)), (lhs, rhs)


@pytest.mark.parametrize("halo", [None, "0", "2", "12"])
def test_large_offsets_leave_the_staged_band(halo, monkeypatch):
    """Offsets far larger than the band kernel's halo (sigma = 7 px, some beyond the image): taps that leave the staged band
    take the global-memory path; every halo setting (RDFC_NLSPN_HALO) must give the oracle's numbers."""
    from oracle import dcn as odcn
    from rdfc_gan_b200 import _cabi as C
    if halo is not None:
        monkeypatch.setenv("RDFC_NLSPN_HALO", halo)
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
                                               B, H, W, T, 0, C.stream_ptr()))
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL
