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
