"""CPU: the C-ABI library loads and exports everything include/rdfc_b200.h declares; host-side argument validation;
state_dict compatibility with the reference; the product path never touches oracle/ and has no CPU fallback."""
import ctypes
import json
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rdfc_gan_b200 import _cabi as C
    hdr = open(os.path.join(ROOT, "include", "rdfc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rdfc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    nm = subprocess.run(["nm", "-D", "--defined-only", C.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rdfc_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    assert set(C.EXPORTS) <= exported
    assert C.lib.rdfc_abi_version() == 2


def test_host_side_validation_without_gpu():
    from rdfc_gan_b200 import _cabi as C
    s = C.DcnShape(2, 4, 6, 7, 8, 3, 3, 2, 2, 1, 1, 1, 1, 2, 1, 64)
    ho, wo = C.c_int(), C.c_int()
    assert C.lib.rdfc_dcn_out_size(s, ho, wo) == 0 and (ho.value, wo.value) == (3, 4)
    bad = C.DcnShape(3, 4, 6, 7, 8, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2)       # 3 % min(3,2) != 0
    assert C.lib.rdfc_dcn_out_size(bad, ho, wo) == -1
    assert b"im2col_step" in C.lib.rdfc_last_error()
    bad = C.DcnShape(2, 5, 6, 7, 8, 3, 3, 1, 1, 1, 1, 1, 1, 2, 1, 64)       # 5 channels, 2 groups
    assert C.lib.rdfc_dcn_out_size(bad, ho, wo) == -1 and b"group" in C.lib.rdfc_last_error()
    with pytest.raises(RuntimeError):
        C.check(C.lib.rdfc_conv_forward(None, None))
    assert C.lib.rdfc_instnorm_nchunk(228 * 304) == (228 * 304 + 255) // 256


def test_state_dict_matches_reference_layout(golden_dir):
    """tests/golden/state_dict_keys.json was dumped from the reference's RDFGenerator (make_golden.py)."""
    from rdfc_gan_b200.generator import DCVGANGenerator, RDFGenerator
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    nl = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
    for name, kw in (("rdfc_r18", dict(use_nlspn_refine=True, nlspn_configs=nl)),
                     ("rdfc_r34_weighting", dict(encoder_rgb="resnet34", encoder_depth="resnet34", semantic_channels_in=40,
                                                 adain_weighting=True, use_nlspn_refine=True, nlspn_configs=nl)),
                     ("rdfc_in_no_nlspn", dict(fuse_depth_in_rgb_decoder="IN", use_nlspn_refine=False))):
        sd = RDFGenerator(pretrained_on_imagenet=False, **kw).state_dict()
        assert [[k, list(v.shape)] for k, v in sd.items()] == ref[name], name
    G = DCVGANGenerator(None, pretrained_on_imagenet=False, use_nlpsn_refine=True, nlspn_configs=nl)
    assert "gd_dec0.0.weight" in G.state_dict() and "fuse_layer5.style.linear.weight_orig" in G.state_dict()
    # init details the reference fixes (nlspn_model.py:37-44, model_utils.py:60-61)
    pl = G.nlspn_refine_module.prop_layer
    assert float(pl.aff_scale_const) == 4.0 and not pl.w.requires_grad and float(pl.conv_offset_aff.weight.abs().sum()) == 0
    b = G.fuse_layer1.style.linear.bias
    assert torch.equal(b[:512], torch.ones(512)) and torch.equal(b[512:], torch.zeros(512))


def test_no_cpu_fallback_and_no_oracle_in_product():
    from rdfc_gan_b200.dcn import DCN, ModulatedDeformConv
    from rdfc_gan_b200.generator import RDFGenerator
    from rdfc_gan_b200.nlspn import NLSPNRefineModule
    x = torch.randn(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):       # modulated_deform_conv.h:43
        DCN.modulated_deform_conv_forward(x, torch.ones(1, 1, 3, 3), torch.zeros(1), torch.zeros(1, 18, 8, 8),
                                          torch.ones(1, 9, 8, 8), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    with pytest.raises(RuntimeError):
        ModulatedDeformConv(1, 1, 3, 1, 1)(x, torch.zeros(1, 18, 8, 8), torch.ones(1, 9, 8, 8))
    with pytest.raises(RuntimeError):
        NLSPNRefineModule()(x, torch.zeros(1, 8, 8, 8), torch.ones(1, 1, 8, 8), x)
    with pytest.raises(RuntimeError, match="CUDA"):
        RDFGenerator(pretrained_on_imagenet=False).eval()(torch.zeros(1, 3, 32, 32), torch.zeros(1, 1, 32, 32),
                                                           torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="parameter container"):
        RDFGenerator(pretrained_on_imagenet=False).rgb_branch_encoder_decoder(x)
    pkg = os.path.join(ROOT, "rdfc_gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libdcn_oracle" not in src and "/root/reference" not in src.replace("/root/reference/RDF", "REF"), f


def test_dcn_shim_installs_as_module():
    import sys
    from rdfc_gan_b200.dcn import DCN
    saved = sys.modules.pop("DCN", None)
    try:
        DCN.install()
        import DCN as imported
        for n in ("deform_conv_forward", "deform_conv_backward", "modulated_deform_conv_forward",
                  "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward"):
            assert callable(getattr(imported, n))          # deformconv/src/vision.cpp:7-12
    finally:
        sys.modules.pop("DCN", None)
        if saved is not None:
            sys.modules["DCN"] = saved


def test_patchgan_state_dict_matches_reference(golden_dir):
    """tests/golden/patchgan_state_dict_keys.json was dumped from the reference's PatchGANDiscriminator(in_channels=1)."""
    from rdfc_gan_b200.discriminator import PatchGANDiscriminator
    want = json.load(open(os.path.join(golden_dir, "patchgan_state_dict_keys.json")))
    got = {k: list(v.shape) for k, v in PatchGANDiscriminator(in_channels=1).state_dict().items()}
    assert got == want


def test_esanet_state_dict_matches_reference(golden_dir):
    """tests/golden/esanet_r34_state_dict_keys.json was dumped from the reference's ESANetOneModality (F/bash/test_nyuv2_Ts2T.sh flags)."""
    from make_esanet_golden import ESANET_CASES
    from rdfc_gan_b200.esanet import ESANetOneModality
    want = json.load(open(os.path.join(golden_dir, "esanet_r34_state_dict_keys.json")))
    got = {k: list(v.shape) for k, v in ESANetOneModality(**ESANET_CASES["r34_full"]["kw"]).state_dict().items()}
    assert got == want
    with pytest.raises(NotImplementedError):
        ESANetOneModality(pretrained_on_imagenet=False, encoder='resnet50')


def test_nhwc_view_descriptors_without_gpu():
    """_cabi.view / nhwc_viewable: dense NHWC tensors and CHANNEL slices of them become (pointer, channels, pixel stride) views --
    what the training ops hand to the kernels instead of .contiguous() copies; anything else is refused (host logic only, CPU tensors)."""
    import torch
    from rdfc_gan_b200 import _cabi as C
    big = torch.zeros(2, 5, 7, 192, dtype=torch.bfloat16)
    v = C.view(big)
    assert (v.C, v.pix_stride) == (192, 192)
    sl = big[..., 64:128]
    assert not sl.is_contiguous() and C.nhwc_viewable(sl)
    vs = C.view(sl)
    assert (vs.C, vs.pix_stride) == (64, 192) and vs.ptr == big.data_ptr() + 64 * 2
    v2 = C.view(big, 32, 16)                                   # explicit channel window of a dense tensor
    assert (v2.C, v2.pix_stride) == (32, 192) and v2.ptr == big.data_ptr() + 16 * 2
    assert not C.nhwc_viewable(big[:, :, ::2])                 # spatial slices, permutations and misaligned slices are not views
    assert not C.nhwc_viewable(big.permute(0, 3, 1, 2))
    assert not C.nhwc_viewable(big[..., 4:68])                 # 8-byte aligned start: the kernels need 16
    with pytest.raises(AssertionError):
        C.view(big[:, :, ::2])
