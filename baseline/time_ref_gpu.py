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


def full_generator():
    """The reference's unmodified RDFGenerator (PyTorch / cuDNN convs + its own DCN extension) against this repo's, same
    synthetic state dict and inputs: parity at B = 1 (fp32 mode here), throughput at B = 32."""
    import bench
    from _synth import synth_inputs, synth_state_dict
    from rdfc_gan_b200.generator import RDFGenerator
    sys.path.insert(0, REFD)
    # rdf_generator.py imports ESANet (never used by RDFGenerator.forward); its own imports need the rest of the repo, so a
    # stub module stands in for it -- the generator files themselves stay unmodified
    for name in ("rdf_generator.segmentator", "rdf_generator.segmentator.esa_net"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    stub = types.ModuleType("rdf_generator.segmentator.esa_net.esa_net_one_modality")
    stub.ESANetOneModality = type("ESANetOneModality", (), {})
    sys.modules[stub.__name__] = stub
    from rdf_generator.rdf_generator import RDFGenerator as RefG
    nl = bench.NLSPN_CFG
    ours = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=nl).eval()
    sd = synth_state_dict(ours, seed=0, recipe="init", nlspn_stress=True)
    ours.load_state_dict(sd)
    theirs = RefG(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=nl).eval()
    theirs.load_state_dict(sd)
    ours, theirs = ours.cuda(), theirs.cuda()
    for B in (1, 32):
        rgb, normal, depth = (t.cuda() for t in synth_inputs(B, 228, 304, seed=0))
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = False          # parity check against true-fp32 cuDNN convs ...
            ot = theirs(rgb, depth, normal)
            o32 = ours.set_precision("fp32")(rgb, depth, normal)
            diff32 = max(float((o32[k] - ot[k]).abs().max()) for k in ot)
            torch.backends.cudnn.allow_tf32 = True           # ... timing with torch's default (TF32 convs allowed)
            ot = theirs(rgb, depth, normal)
            diff_tf32 = max(float((o32[k] - ot[k]).abs().max()) for k in ot)
            t_ref = timeit(lambda: theirs(rgb, depth, normal), reps=5)
            o16 = ours.set_precision("bf16")(rgb, depth, normal)
            diff16 = max(float((o16[k] - ot[k]).abs().max()) for k in ot)
            t_ours = timeit(lambda: ours(rgb, depth, normal), reps=5)
        print(f"RDFGenerator forward B={B} 228x304: reference (PyTorch/cuDNN fp32 + its DCN extension) {t_ref:8.2f} ms = {B / t_ref * 1e3:7.1f} maps/s | "
              f"this repo bf16 {t_ours:7.2f} ms = {B / t_ours * 1e3:7.1f} maps/s | speed-up {t_ref / t_ours:5.1f}x | "
              f"max-abs diff vs reference: fp32 mode {diff32:.2e} (reference with TF32 convs, torch's default: {diff_tf32:.2e}), bf16 mode {diff16:.2e}" + ("" if B == 1 else "  (B > 1: reference quirk 5)"))


def json_mode(gen, B, H, W):
    """bench.py's `ref_gpu`: the unmodified reference generator on this GPU (its own DCN extension, PyTorch/cuDNN convs with
    torch's default TF32 setting), same synthetic weights and inputs as the bench.  Imports nothing of the product package."""
    import json
    import bench
    import ref_loader
    from _synth import synth_inputs
    cfg = dict(bench.CONFIGS["c3" if gen == "rdfc" else "c2"], H=H, W=W)
    G = bench.build_reference(cfg, gpu=True).cuda()
    rgb, stem, depth = (t.cuda() for t in synth_inputs(B, H, W, seed=0, Cs=cfg["cs"]))
    with torch.no_grad():
        t = timeit(lambda: bench.call_generator(G, cfg, rgb, stem, depth), reps=5)
    print(json.dumps({"value": B / t * 1e3, "unit": "maps/s", "ms_per_step": t, "batch": B,
                      "what": "unmodified reference generator (PyTorch/cuDNN fp32, TF32 convs allowed = torch default) + its own DCN "
                              "CUDA extension compiled for sm_100a, same GPU, same weights / inputs; median of 5"}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--json":
        sys.path.insert(0, HERE)
        json_mode(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    else:
        main()
        full_generator()
