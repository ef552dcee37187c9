#!/usr/bin/env python
"""tests/golden/esanet_*.npz: outputs of the REFERENCE's ESANetOneModality (F/lib/models/segmentator/esa_net, imported through
baseline/ref_loader.py; build container only) on synthetic weights (tests/_synth.py 'scaled' recipe) and inputs.

    python tests/golden/make_esanet_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "baseline"))

_COMMON = dict(num_classes=40, pretrained_on_imagenet=False, encoder_block='BasicBlock', encoder_decoder_fusion='add', context_module='ppm',
               weighting_in_encoder='SE-add', upsampling='learned-3x3-zeropad', pyramid_supervision=False)
ESANET_CASES = {
    # F/bash/test_nyuv2_Ts2T.sh:7-16 at the NYUv2 size
    "r34_full": dict(kw=dict(_COMMON, height=228, width=304, encoder='resnet34', channels_decoder=[512, 256, 128], nr_decoder_blocks=[3, 3, 3]),
                     B=1, H=228, W=304, seed=41, stride=4),
    "r18_small": dict(kw=dict(_COMMON, height=72, width=104, encoder='resnet18', channels_decoder=[128, 128, 128], nr_decoder_blocks=[1, 1, 1]),
                      B=2, H=72, W=104, seed=42, stride=2),
}


def esanet_weights(net, seed):
    """tests/_synth.py 'scaled' recipe, with the last BatchNorm of every residual branch damped: ESANet-34 stacks 16 + 9 residual
    blocks in front of eval-mode BatchNorms whose (synthetic) running statistics do not track the activations, so undamped branches
    let the signal grow ~2x per block (logits of 1e7)."""
    from _synth import synth_state_dict
    sd = synth_state_dict(net, seed=seed, recipe="scaled")
    for k in sd:
        if k.endswith("bn2.weight"):
            sd[k] = sd[k] * 0.25
    return sd


def esanet_input(c):
    from _synth import _rng
    r = _rng(c["seed"], "esanet_rgb")
    return torch.from_numpy(r.uniform(-1, 1, (c["B"], 3, c["H"], c["W"])).astype(np.float32))


def main():
    import ref_loader
    from _synth import state_dict_digest
    assert ref_loader.install()
    _, ESANet = ref_loader.load_rdf_gan()
    torch.set_num_threads(8)
    for name, c in ESANET_CASES.items():
        net = ESANet(**c["kw"]).eval()
        sd = esanet_weights(net, c["seed"])
        net.load_state_dict(sd)
        with torch.no_grad():
            y = net.forward_net(esanet_input(c))
        np.savez_compressed(os.path.join(HERE, f"esanet_{name}.npz"), logits=y.numpy()[:, :, ::c["stride"], ::c["stride"]],
                            digest=np.array([state_dict_digest(sd)], np.int64), rms=np.array(float(y.square().mean().sqrt())))
        print(name, tuple(y.shape), "rms", float(y.square().mean().sqrt()), "max", float(y.abs().max()))


if __name__ == "__main__":
    main()
