"""CPU: pins the oracle (oracle/) to the golden vectors that tests/golden/make_golden.py produced from the
reference's own Python -- every oracle function the GPU parity tests rely on is checked here first."""
import numpy as np
import pytest
import torch

from _synth import (DCN_CASES, dcn_case_inputs, nlspn_stress_inputs, state_dict_digest, synth_inputs,
                    synth_state_dict)
from make_golden import GEN_CASES, NLSPN_CASES, NLSPN_SHAPE


@pytest.mark.parametrize("name", list(DCN_CASES))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_dcn_oracle_matches_reference(name, dtype, tol, golden_dir):
    from oracle import dcn as odcn
    case = DCN_CASES[name]
    t = dcn_case_inputs(case, dtype=dtype)
    gold = np.load(f"{golden_dir}/dcn_{name}.npz")
    k, s, p, d, g, dg = (case[x] for x in ("k", "s", "p", "d", "g", "dg"))
    geo = (k, k, s, s, p, p, d, d, g, dg)
    out = odcn.modulated_deform_conv_forward(t["input"], t["weight"], t["bias"], t["offset"], t.get("mask"), *geo)
    grads = odcn.modulated_deform_conv_backward(t["input"], t["weight"], t["bias"], t["offset"], t.get("mask"),
                                                t["grad_output"], *geo)
    assert out.dtype == dtype
    assert np.abs(out - gold["output"]).max() <= tol * max(1, np.abs(gold["output"]).max())
    for n, gr in zip(("grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"), grads):
        if gr is None:
            assert n == "grad_mask" and not case["mask"]
            continue
        assert np.abs(gr - gold[n]).max() <= tol * max(1, np.abs(gold[n]).max()), n


def test_dcn_oracle_known_answers():
    """deformconv/test.py:69-110,142-181: zero offsets (+ mask 1) == Conv2d; identity filters with mask 0.5 halve."""
    from oracle import dcn as odcn
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 4, 4))
    w = rng.standard_normal((4, 2, 3, 3))
    b = rng.standard_normal(4)
    off = np.zeros((2, 18, 4, 4))
    out = odcn.modulated_deform_conv_forward(x, w, b, off, np.ones((2, 9, 4, 4)), 3, 3, 1, 1, 1, 1, 1, 1, 2, 1)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), 1, 1, 1, 2).numpy()
    assert np.abs(out - ref).max() < 1e-12
    wi = np.zeros((4, 2, 3, 3))
    for q in range(4):
        wi[q, q % 2, 1, 1] = 1.0
    out = odcn.modulated_deform_conv_forward(x, wi, np.zeros(4), off, np.full((2, 9, 4, 4), 0.5), 3, 3, 1, 1, 1, 1, 1, 1, 2, 1)
    assert np.abs(2 * out - x).max() < 1e-12
    assert np.abs(odcn.deform_conv_forward(x, wi, np.zeros(4), off, 3, 3, 1, 1, 1, 1, 1, 1, 2, 1) - x).max() < 1e-12


@pytest.mark.parametrize("name", list(NLSPN_CASES))
def test_nlspn_oracle_matches_reference(name, golden_dir):
    from oracle import nlspn as onl
    cfg = NLSPN_CASES[name]
    x = nlspn_stress_inputs(*NLSPN_SHAPE, cfg["seed"])
    gold = np.load(f"{golden_dir}/nlspn_{name}.npz")
    y, off, aff, inter = onl.nlspn_forward(x["pred_init"], x["guidance"], x["confidence"], x["feat_fix"], x["conv_w"],
                                           x["conv_b"], gold["aff_scale"], prop_time=cfg["prop_time"],
                                           affinity=cfg["affinity"], conf_prop=cfg["conf_prop"],
                                           preserve_input=cfg["preserve_input"], return_inter=True)
    y_fast, _, _ = onl.nlspn_forward(x["pred_init"], x["guidance"], x["confidence"], x["feat_fix"], x["conv_w"],
                                     x["conv_b"], gold["aff_scale"], prop_time=cfg["prop_time"], affinity=cfg["affinity"],
                                     conf_prop=cfg["conf_prop"], preserve_input=cfg["preserve_input"])
    assert np.abs(off - gold["offset"]).max() <= 1e-6
    assert np.abs(aff - gold["aff"]).max() <= 5e-6
    assert np.abs(inter[0] - gold["first"]).max() <= 5e-6
    assert np.abs(y - gold["y"]).max() <= 5e-6
    assert np.abs(y_fast - y).max() <= 1e-6          # the fused C loop == the call-per-iteration restatement
    # structure the reference guarantees (nlspn_model.py:77-80,131-136)
    assert np.all(off[:, 8:10] == 0) and np.abs(aff.sum(1) - 1).max() < 1e-5


@pytest.mark.parametrize("name", [n for n in GEN_CASES if "full" not in n] + ["rdfc_full_init"])
def test_generator_oracle_matches_reference(name, golden_dir):
    from oracle import generator as ogen
    from rdfc_gan_b200.generator import RDFGenerator
    kw, B, H, W, Cs, recipe, stress, seed = GEN_CASES[name]
    G = RDFGenerator(pretrained_on_imagenet=False, **kw).eval()
    sd = synth_state_dict(G, seed=seed, recipe=recipe, nlspn_stress=stress)
    gold = np.load(f"{golden_dir}/generator_{name}.npz")
    assert state_dict_digest(sd) == int(gold["digest"][0])
    rgb, stem, depth = synth_inputs(B, H, W, seed=seed, Cs=Cs)
    out = ogen.generator_forward(sd, stem, depth, fuse=kw.get("fuse_depth_in_rgb_decoder", "WAdaIN"),
                                 adain_weighting=kw.get("adain_weighting", False), use_nlspn_refine=kw["use_nlspn_refine"],
                                 nlspn_configs=kw.get("nlspn_configs"))
    for k in ("depth_map_1", "confidence_map_1", "depth_map_2", "confidence_map_2", "pred_depth"):
        v, g = out[k].numpy(), gold[k]
        if v.shape != g.shape:
            v = v[:, :, ::4, ::4]
        assert np.abs(v - g).max() <= 2e-5, k


@pytest.mark.parametrize("name", ["dcvgan_r34", "dcvgan_r18_b2", "rdfc_full_b4"])
def test_generator_oracle_matches_round2_goldens(name, golden_dir):
    """The oracle restatement against the round-2 goldens: RDF-GAN's DCVGANGenerator class itself (same body over a 40-channel
    stem input) and the bench recipe at B = 4, 228x304."""
    from make_golden import GEN_CASES_V2
    from oracle import generator as ogen
    from _synth import synth_inputs, synth_state_dict
    c = GEN_CASES_V2[name]
    gold = np.load(f"{golden_dir}/generator_{name}.npz")
    # state-dict keys and shapes of both generators are those of the product's parameter containers (pinned against the
    # reference's key list in test_cabi_and_layout.py); the digest proves the weights are the golden's
    from rdfc_gan_b200.generator import DCVGANGenerator, RDFGenerator
    import torch
    G = (RDFGenerator(pretrained_on_imagenet=False, **c["kw"]) if c["cls"] == "rdfc" else
         DCVGANGenerator(torch.nn.Identity(), pretrained_on_imagenet=False, **c["kw"]))
    sd = synth_state_dict(G, seed=c["seed"], recipe=c["recipe"], nlspn_stress=c["stress"])
    from _synth import state_dict_digest
    assert state_dict_digest(sd) == int(gold["digest"][0])
    rgb, stem, depth = synth_inputs(c["B"], c["H"], c["W"], seed=c["seed"], Cs=c["Cs"])
    torch.set_num_threads(8)
    out = ogen.generator_forward(sd, stem, depth, adain_weighting=c["kw"].get("adain_weighting", False), use_nlspn_refine=True,
                                 nlspn_configs=c["kw"]["nlspn_configs"])
    for k, (stride, imgs) in c["store"].items():
        v = out[k].numpy()
        v = (v if imgs is None else v[list(imgs)])[:, :, ::stride, ::stride]
        assert np.abs(v - gold[k]).max() <= 2e-5, (k, np.abs(v - gold[k]).max())


def test_oracle_nlspn_backward_matches_finite_differences():
    """oracle.nlspn.nlspn_propagate_backward (the reference's reverse loop on the C DCN oracle) against central finite
    differences of the oracle forward: pins the checker the GPU backward is compared with."""
    from oracle import dcn as odcn
    from oracle import nlspn as onl
    from _synth import nlspn_stress_inputs
    B, H, W, T = 1, 9, 11, 3
    x = nlspn_stress_inputs(B, H, W, 3)
    off, aff = onl.get_offset_affinity(x["guidance"], x["confidence"], x["conv_w"], x["conv_b"], np.array([4.0], np.float32))[:2]
    off, aff = np.ascontiguousarray(off, np.float32), np.ascontiguousarray(aff, np.float32)
    gout = np.random.default_rng(0).standard_normal((B, 1, H, W)).astype(np.float32)
    for preserve in (False, True):
        gf, go, ga = onl.nlspn_propagate_backward(gout, x["pred_init"], off, aff, x["feat_fix"], preserve, T)

        def loss(f, o, a):
            return float((odcn.nlspn_propagate(f, o, a, x["feat_fix"], preserve, 3, T).astype(np.float64) * gout).sum())
        for name, arr, grad, idxs in (("feat", x["pred_init"], gf, [(0, 0, 4, 5), (0, 0, 2, 9)]),
                                      ("aff", aff, ga, [(0, 3, 4, 5), (0, 7, 2, 9)])):
            for idx in idxs:
                e = 1e-2
                hi, lo = arr.copy(), arr.copy()
                hi[idx] += e
                lo[idx] -= e
                args = {"feat": lambda v: (v, off, aff), "aff": lambda v: (x["pred_init"], off, v)}[name]
                fd = (loss(*args(hi)) - loss(*args(lo))) / (2 * e)        # the op is linear in feat and in aff
                assert abs(fd - float(grad[idx])) <= 2e-3 * max(1.0, abs(fd)), (preserve, name, idx, fd, float(grad[idx]))


def test_oracle_metrics_vs_reference_golden(golden_dir):
    """oracle.metrics against the outputs of the reference's own RDFGANMetric (tests/golden/make_metric_golden.py)."""
    import json
    from oracle import metrics as om
    from _synth import metric_inputs
    gold = json.load(open(f"{golden_dir}/metric_golden.json"))
    for seed, g in gold.items():
        n_img, H, W, with_mask = g["cfg"]
        res = metric_inputs(int(seed), n_img, H, W, with_mask)
        got = om.evaluate_all(res)
        for k, v in g["evaluate_all"].items():
            assert abs(float(got[k]) - v) <= 2e-6 * max(1.0, abs(v)), (seed, k, float(got[k]), v)
        b = om.evaluate_batch(np.stack([r["gt"] for r in res]), np.stack([r["pd"] for r in res]))
        assert np.allclose(b[0], np.array(g["evaluate_batch"], np.float32), rtol=2e-5, atol=1e-6), (seed, b, g["evaluate_batch"])
