"""Times the reference's OWN NLSPN (unmodified nlspn_model.py driving the reference's own CUDA DCN extension, built for
sm_100a by baseline/build_ref_gpu.py into the git-ignored baseline/_ref/) next to this repo's fused kernels, on the same
B200, same weights and inputs (SURVEY 8d "also time").  Development / reporting aid: not part of tests, smoke or bench."""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFD = os.path.join(HERE, "_ref")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_reference():
    spec = importlib.util.spec_from_file_location("DCN", os.path.join(REFD, "build", "DCN.so"))
    dcn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dcn)
    sys.modules["DCN"] = dcn
    pkg = types.ModuleType("refnlspn")
    pkg.__path__ = [REFD]
    sys.modules["refnlspn"] = pkg
    import refnlspn.nlspn_model as m
    return m


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]


def main():
    from rdfc_gan_b200.nlspn import NLSPNRefineModule
    ref = load_reference()
    H, W = 228, 304
    cfg = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
    for B in (1, 32):
        ours = NLSPNRefineModule(**cfg).eval()
        theirs = ref.NLSPNRefineModule(**cfg).eval()
        g = torch.Generator().manual_seed(0)
        pl = ours.prop_layer
        with torch.no_grad():
            pl.conv_offset_aff.weight[:16].normal_(0, 0.25, generator=g)
            pl.conv_offset_aff.bias[:16].uniform_(-1.5, 1.5, generator=g)
            pl.conv_offset_aff.weight[16:].normal_(0, 0.02, generator=g)
            pl.conv_offset_aff.bias[16:].uniform_(0.3, 2.0, generator=g)
        theirs.load_state_dict(ours.state_dict())
        ours, theirs = ours.cuda(), theirs.cuda()
        guide = torch.randn(B, 8, H, W, generator=g).cuda()
        conf = torch.rand(B, 1, H, W, generator=g).cuda()
        init = (2 * torch.rand(B, 1, H, W, generator=g) - 1).cuda()
        fix = torch.zeros(B, 1, H, W).cuda()
        with torch.no_grad():
            yo, _ = ours(init, guide, conf, fix)
            yt, _ = theirs(init, guide, conf, fix)
            torch.cuda.synchronize()
            diff = float((yo - yt).abs().max())
            t_ours = timeit(lambda: ours(init, guide, conf, fix))
            t_ref = timeit(lambda: theirs(init, guide, conf, fix))
        note = "" if B == 1 else "  (B > 1: the reference kernel reads non-contiguous offset views as contiguous, SURVEY 8a quirk 5)"
        print(f"NLSPN refine B={B} {H}x{W} 18 it.: reference CUDA extension {t_ref:8.2f} ms | this repo {t_ours:7.3f} ms | "
              f"speed-up {t_ref / t_ours:6.1f}x | max-abs diff {diff:.2e}{note}")


if __name__ == "__main__":
    main()
