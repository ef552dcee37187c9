"""Development aid: BASELINE config 2 (RDF-GAN body: ResNet-34 encoders, 40-channel guidance stem, W-AdaIN with
weighting, NLSPN 18 it.) at B=32, 228x304, bf16: per-step plan timing."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _synth import synth_inputs, synth_state_dict
from rdfc_gan_b200.generator import RDFGenerator
from rdfc_gan_b200 import _cabi as C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nl = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False)
G = RDFGenerator(encoder_rgb="resnet34", encoder_depth="resnet34", semantic_channels_in=40, adain_weighting=True,
                 pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=nl).eval()
G.load_state_dict(synth_state_dict(G, seed=0, recipe="init", nlspn_stress=True))
G = G.cuda().set_precision("bf16")
rgb, stem, depth = synth_inputs(B, 228, 304, seed=0, Cs=40)
with torch.no_grad():
    G(rgb.cuda(), depth.cuda(), stem.cuda())
plan = next(iter(G.engine()._plans.values()))
s = C.stream_ptr()
acc = [[] for _ in plan.steps]
for rep in range(5):
    evs = []
    for f in plan.steps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(s); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    if rep >= 2:
        for i, (a, b) in enumerate(evs): acc[i].append(a.elapsed_time(b) * 1e3)
med = [sorted(v)[len(v) // 2] for v in acc]
tot = sum(med)
print(f"config 2 plan B={B}: {len(med)} steps, {tot/1e3:.2f} ms -> {B/tot*1e6:.0f} maps/s")
for t, n in sorted(zip(med, plan.names), reverse=True)[:14]:
    print(f"  {t:9.1f} us  {100*t/tot:5.1f}%  {n}")
