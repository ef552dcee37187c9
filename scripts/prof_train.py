"""Development aid: where the training step's GPU time goes (torch.profiler, kernels grouped by name).
    python scripts/prof_train.py [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench_train as bt
from _synth import synth_state_dict
from rdfc_gan_b200.discriminator import PatchGANDiscriminator
from rdfc_gan_b200.generator import RDFGenerator
from rdfc_gan_b200.rdf_gan import RDFGAN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = RDFGenerator(pretrained_on_imagenet=False, use_nlspn_refine=True, nlspn_configs=bt.NLSPN_CFG)
G.load_state_dict(synth_state_dict(G, seed=0, recipe="init", nlspn_stress=True))
D = PatchGANDiscriminator(in_channels=1); D.load_state_dict(synth_state_dict(D, seed=1, recipe="init"))
m = RDFGAN(G, D, device="cuda", args=bt.ARGS); m.train()
data = {k: v.cuda() for k, v in bt.synth_batch(B, 0).items()}
for _ in range(3):
    m.set_input(data); m.optimize_parameters()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.set_input(data); m.optimize_parameters(); torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("<")[0].split("(")[0][:70]
        r = rows.setdefault(n, [0.0, 0]); r[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time; r[1] += 1
tot = sum(r[0] for r in rows.values())
print(f"B={B}: {tot/1e3:.1f} ms of GPU kernel time in one step")
for n, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
if os.environ.get("OPS", "1") == "1":
    # the PyTorch ops behind the at::native kernels: self device time by (op, input shapes)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
        m.set_input(data); m.optimize_parameters(); torch.cuda.synchronize()
    rows2 = [(e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:110]) for e in prof2.key_averages(group_by_input_shape=True)
             if e.key.startswith("aten::") and e.self_device_time_total > 0]
    print("PyTorch ops by self device time (op, input shapes):")
    for t, c, k, shp in sorted(rows2, reverse=True)[:60]:
        print(f"{t/1e3:8.2f} ms  x{c:4d}  {k:34s} {shp}")
