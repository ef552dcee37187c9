"""init_weights (C/lib/models/init_weights.py:5-33) on the drop-in generator reproduces the reference's tensors bit for
bit: same module traversal order, same draws.  Golden checksum: tests/golden/make_init_golden.py."""
import json
import os

import pytest
import torch

from golden.make_init_golden import NL, digest


@pytest.mark.parametrize("init_type", ["normal", "kaiming"])
def test_init_weights_matches_reference_checksum(init_type, golden_dir):
    from rdfc_gan_b200.generator import RDFGenerator
    from rdfc_gan_b200.init_weights import init_weights
    gold = json.load(open(os.path.join(golden_dir, "init_weights.json")))
    if gold["torch"].split("+")[0] != torch.__version__.split("+")[0]:
        pytest.skip("RNG streams are only comparable within one torch version")
    G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=NL)
    torch.manual_seed(0)
    init_weights(G, init_type=init_type)
    assert digest(G.state_dict()) == gold["crc32"][init_type]


def test_init_weights_semantics():
    from rdfc_gan_b200.generator import RDFGenerator
    from rdfc_gan_b200.init_weights import init_weights
    G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=NL)
    lin = G.fuse_layer1.style.linear
    w_orig, b_lin = lin.weight_orig.clone(), lin.bias.clone()
    init_weights(G)
    pl = G.nlspn_refine_module.prop_layer
    assert float(pl.conv_offset_aff.weight.abs().sum()) > 0 and float(pl.conv_offset_aff.bias.abs().sum()) == 0   # zero init overwritten
    assert torch.equal(lin.weight_orig, w_orig) and torch.equal(lin.bias, b_lin)                                  # EqualLinear skipped
    assert abs(float(G.rgb_branch_encoder_decoder.en2[0].bn1.weight.mean()) - 1.0) < 0.02
    assert abs(float(G.rgb_pred_dec1[0].weight.std()) - 0.02) < 0.002
    with pytest.raises(NotImplementedError):
        init_weights(G, init_type="bogus")
