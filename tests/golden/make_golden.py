#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python (needs /root/reference; build container only).

    python tests/golden/make_golden.py [dcn] [nlspn] [generator] [generator_v2]

* dcn_<case>.npz        reference ModulatedDeformConvFunction / DeformConvFunction forward (the Function's call into
                        ``DCN`` is served by torchvision.ops.deform_conv2d, see _ref_import.py) and the five gradients
                        (torchvision autograd == the reference's col2im arithmetic).
* nlspn_<case>.npz      reference NLSPNRefineModule / NLPSN outputs (offset, aff, result) on the stress fixture.
* generator_<case>.npz  reference RDFGenerator outputs on synthetic weights (tests/_synth.py) and inputs.
Inputs and weights are NOT stored: tests regenerate them from the same seeds (a CRC of the weights is stored).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _ref_import import import_rdfc  # noqa: E402
from _synth import (DCN_CASES, dcn_case_inputs, nlspn_stress_inputs, state_dict_digest, synth_inputs,  # noqa: E402
                    synth_state_dict)

NLSPN_CASES = {
    "tgass18": dict(affinity="TGASS", prop_time=18, preserve_input=False, conf_prop=True, seed=0),
    "tgass1": dict(affinity="TGASS", prop_time=1, preserve_input=False, conf_prop=True, seed=1),
    "as12_preserve": dict(affinity="AS", prop_time=12, preserve_input=True, conf_prop=True, seed=2),
    "ass18": dict(affinity="ASS", prop_time=18, preserve_input=False, conf_prop=True, seed=3),
    "tc12_noconf": dict(affinity="TC", prop_time=12, preserve_input=False, conf_prop=False, seed=0),
    "tgass18_preserve": dict(affinity="TGASS", prop_time=18, preserve_input=True, conf_prop=True, seed=1),
}
NLSPN_SHAPE = (2, 24, 32)

_NL = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
GEN_CASES = {
    # name: (ctor kwargs, B, H, W, Cs, recipe, nlspn_stress, seed)
    "rdfc_small": (dict(use_nlspn_refine=True, nlspn_configs=_NL), 2, 36, 52, 3, "scaled", True, 1),
    "rdfc_small_init": (dict(use_nlspn_refine=True, nlspn_configs=_NL), 1, 36, 52, 3, "init", False, 2),
    "rdfc_full_init": (dict(use_nlspn_refine=True, nlspn_configs=_NL), 1, 228, 304, 3, "init", False, 0),
    "rdfc_full_scaled": (dict(use_nlspn_refine=True, nlspn_configs=_NL), 1, 228, 304, 3, "scaled", True, 3),
    "rdf_r34_weighting": (dict(encoder_rgb="resnet34", encoder_depth="resnet34", semantic_channels_in=40,
                               adain_weighting=True, use_nlspn_refine=True, nlspn_configs=_NL), 1, 40, 56, 40, "scaled", True, 4),
    "fuse_adain": (dict(fuse_depth_in_rgb_decoder="AdaIN", use_nlspn_refine=True, nlspn_configs=_NL), 1, 36, 52, 3, "scaled", True, 5),
    "fuse_in": (dict(fuse_depth_in_rgb_decoder="IN", use_nlspn_refine=True, nlspn_configs=_NL), 1, 36, 52, 3, "scaled", True, 6),
    "no_nlspn": (dict(use_nlspn_refine=False), 1, 36, 52, 3, "scaled", False, 7),
    "as12_preserve": (dict(use_nlspn_refine=True, nlspn_configs=dict(_NL, affinity="AS", prop_time=12, preserve_input=True)),
                      1, 36, 52, 3, "scaled", True, 8),
}


# Batched / full-size / RDF-GAN cases (round 2).  `store`: per output map (stride, image indices or None = all); maps that
# are not listed are not stored.  cls "rdfc" = RDFC-GAN's RDFGenerator, "rdf" = RDF-GAN's DCVGANGenerator (the F/ class itself,
# imported through baseline/ref_loader.py with the nlspn alias of SURVEY 8c; global_guidance_module = nn.Identity()).
_ALL5 = ("depth_map_1", "confidence_map_1", "depth_map_2", "confidence_map_2", "pred_depth")
GEN_CASES_V2 = {
    # the bench's own weights and inputs (init recipe + NLSPN stress, seed 0) at B = 4 and B = 32, 228x304
    "rdfc_full_b4": dict(cls="rdfc", kw=dict(use_nlspn_refine=True, nlspn_configs=_NL), B=4, H=228, W=304, Cs=3, recipe="init",
                         stress=True, seed=0, store={k: (4, None) for k in _ALL5}, full=("pred_depth", 3)),
    "rdfc_full_b32": dict(cls="rdfc", kw=dict(use_nlspn_refine=True, nlspn_configs=_NL), B=32, H=228, W=304, Cs=3, recipe="init",
                          stress=True, seed=0, store={k: (8, None) for k in ("depth_map_1", "depth_map_2", "pred_depth")},
                          full=("pred_depth", 31)),
    "dcvgan_r34": dict(cls="rdf", kw=dict(encoder_rgb="resnet34", encoder_depth="resnet34", semantic_channels_in=40, adain_weighting=True,
                                          use_nlpsn_refine=True, nlspn_configs=_NL), B=1, H=40, W=56, Cs=40, recipe="scaled",
                       stress=True, seed=9, store={k: (1, None) for k in _ALL5}),
    "dcvgan_r18_b2": dict(cls="rdf", kw=dict(semantic_channels_in=40, use_nlpsn_refine=True, nlspn_configs=_NL), B=2, H=36, W=52, Cs=40,
                          recipe="scaled", stress=True, seed=10, store={k: (1, None) for k in _ALL5}),
}


def make_generator_v2():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "baseline"))
    import ref_loader
    assert ref_loader.install(), "reference checkout not present"
    for name, c in GEN_CASES_V2.items():
        torch.manual_seed(0)
        if c["cls"] == "rdfc":
            G = ref_loader.load_rdfc()(pretrained_on_imagenet=False, **c["kw"]).eval()
        else:
            G = ref_loader.load_rdf_gan()[0](torch.nn.Identity(), pretrained_on_imagenet=False, **c["kw"]).eval()
        sd = synth_state_dict(G, seed=c["seed"], recipe=c["recipe"], nlspn_stress=c["stress"])
        G.load_state_dict(sd, strict=True)
        rgb, stem, depth = synth_inputs(c["B"], c["H"], c["W"], seed=c["seed"], Cs=c["Cs"])
        with torch.no_grad():
            out = G(rgb, depth, stem) if c["cls"] == "rdfc" else dict(zip(_ALL5, G(stem, depth)))
        res = {"digest": np.array([state_dict_digest(sd)], np.int64)}
        for k, (stride, imgs) in c["store"].items():
            v = out[k].numpy()
            res[k] = (v if imgs is None else v[list(imgs)])[:, :, ::stride, ::stride]
        if "full" in c:
            k, i = c["full"]
            res[f"full_{k}_{i}"] = out[k][i].numpy()
        np.savez_compressed(os.path.join(HERE, f"generator_{name}.npz"), **res)
        print("generator", name, {k: (float(v.min()), float(v.max())) for k, v in out.items()})


def make_dcn():
    from torchvision.ops import deform_conv2d
    import_rdfc()
    from lib.models.generator.rdf_generator.nlspn.modulated_deform_conv_func import ModulatedDeformConvFunction
    for name, case in DCN_CASES.items():
        t = {k: torch.from_numpy(v).double() for k, v in dcn_case_inputs(case, dtype=np.float64).items()}
        k, s, p, d, g, dg = (case[x] for x in ("k", "s", "p", "d", "g", "dg"))
        inp, w, b, off = (t[n].clone().requires_grad_(True) for n in ("input", "weight", "bias", "offset"))
        msk = t["mask"].clone().requires_grad_(True) if case["mask"] else None
        if g == 1 and case["mask"]:
            # through the reference's own Function (forward only: its backward needs the CUDA extension)
            with torch.no_grad():
                out_fn = ModulatedDeformConvFunction.apply(t["input"], t["offset"], t["mask"], t["weight"], t["bias"], s, p,
                                                           d, g, dg, 64)
        else:
            out_fn = None      # torchvision infers groups from the weight shape; same arithmetic
        out = deform_conv2d(inp, off, w, b, stride=s, padding=p, dilation=d, mask=msk)
        if out_fn is not None:
            assert torch.equal(out_fn, out.detach())
        out.backward(t["grad_output"])
        res = dict(output=out.detach().numpy(), grad_input=inp.grad.numpy(), grad_offset=off.grad.numpy(),
                   grad_weight=w.grad.numpy(), grad_bias=b.grad.numpy())
        if msk is not None:
            res["grad_mask"] = msk.grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"dcn_{name}.npz"), **res)
        print("dcn", name, res["output"].shape)


def make_nlspn():
    _, NLSPNRefineModule, _ = import_rdfc()
    B, H, W = NLSPN_SHAPE
    for name, cfg in NLSPN_CASES.items():
        x = nlspn_stress_inputs(B, H, W, cfg["seed"])
        mod = NLSPNRefineModule(prop_kernel=3, prop_time=cfg["prop_time"], affinity=cfg["affinity"], affinity_gamma=0.5,
                                conf_prop=cfg["conf_prop"], preserve_input=cfg["preserve_input"]).eval()
        pl = mod.prop_layer
        pl.conv_offset_aff.weight.data.copy_(torch.from_numpy(x["conv_w"]))
        pl.conv_offset_aff.bias.data.copy_(torch.from_numpy(x["conv_b"]))
        tt = {k: torch.from_numpy(v) for k, v in x.items()}
        with torch.no_grad():
            y, inter, offset, aff, _ = pl(tt["pred_init"], tt["guidance"], tt["confidence"], tt["feat_fix"])
            y2, conf = mod(tt["pred_init"], tt["guidance"], tt["confidence"], tt["feat_fix"])
        assert torch.equal(y, y2)
        np.savez_compressed(os.path.join(HERE, f"nlspn_{name}.npz"), y=y.numpy(), offset=offset.numpy(), aff=aff.numpy(),
                            first=inter[0].numpy(), aff_scale=pl.aff_scale_const.data.numpy())
        print("nlspn", name, float(offset.std()), float(aff[:, [0, 1, 2, 3, 5, 6, 7, 8]].sum(1).mean()), float(y.abs().max()))


def make_generator():
    RDFGenerator, _, _ = import_rdfc()
    for name, (kw, B, H, W, Cs, recipe, stress, seed) in GEN_CASES.items():
        torch.manual_seed(0)
        G = RDFGenerator(pretrained_on_imagenet=False, **kw).eval()
        sd = synth_state_dict(G, seed=seed, recipe=recipe, nlspn_stress=stress)
        G.load_state_dict(sd, strict=True)
        rgb, stem, depth = synth_inputs(B, H, W, seed=seed, Cs=Cs)
        with torch.no_grad():
            out = G(rgb, depth, stem)
        full = H > 100
        res = {"digest": np.array([state_dict_digest(sd)], np.int64)}
        for k, v in out.items():
            v = v.numpy()
            res[k] = v[:, :, ::4, ::4] if (full and k in ("depth_map_1", "confidence_map_1", "confidence_map_2")) else v
        np.savez_compressed(os.path.join(HERE, f"generator_{name}.npz"), **res)
        print("generator", name, {k: (float(v.min()), float(v.max())) for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1:] or ["dcn", "nlspn", "generator"]
    torch.set_num_threads(8)
    if "dcn" in what:
        make_dcn()
    if "nlspn" in what:
        make_nlspn()
    if "generator" in what:
        make_generator()
    if "generator_v2" in what:
        make_generator_v2()
