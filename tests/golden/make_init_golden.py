"""Generates tests/golden/init_weights.json in the BUILD container (needs /root/reference): the reference's own
RDFGenerator initialised by the reference's own init_weights (C/lib/models/init_weights.py:5-33), reseeded right before
the call, checksummed over the sorted state dict without the constructor-drawn ``*.weight_orig`` tensors."""
import json
import sys
import types
import zlib

import numpy as np
import torch

NL = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)


def digest(sd):
    h = 0
    for k in sorted(sd):
        if 'weight_orig' in k:
            continue
        h = zlib.crc32(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes(), h)
        h = zlib.crc32(k.encode(), h)
    return h


if __name__ == "__main__":
    sys.modules["DCN"] = types.ModuleType("DCN")
    REF = "/root/reference/RDFC-GAN"
    for name, path in [("lib", "/lib"), ("lib.models", "/lib/models"), ("lib.models.generator", "/lib/models/generator")]:
        pkg = types.ModuleType(name)
        pkg.__path__ = [REF + path]
        sys.modules[name] = pkg
    sys.path.insert(0, REF)
    from lib.models.generator.rdf_generator.rdf_generator import RDFGenerator
    from lib.models.init_weights import init_weights
    out = {}
    for it in ("normal", "kaiming"):
        G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=NL)
        torch.manual_seed(0)
        init_weights(G, init_type=it)
        out[it] = digest(G.state_dict())
    json.dump({"recipe": "G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=TGASS/18); "
                         "torch.manual_seed(0); init_weights(G, init_type)", "crc32": out, "torch": torch.__version__},
              open(__file__.replace("make_init_golden.py", "init_weights.json"), "w"), indent=1)
    print(out)
