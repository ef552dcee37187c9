"""Development aid: BASELINE config 2 end to end (ESANet guidance network + RDF-GAN generator, B = 32): GPU time by kernel (torch.profiler;
the CUDA graphs are replayed, so kernels appear under their own names)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench
from _synth import synth_inputs
cfg = bench.CONFIGS[os.environ.get("CONFIG", "c2")]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
G = bench.build_product(cfg).to(dev).set_precision(os.environ.get("PRECISION", "bf16"))
rgb, normal, depth = synth_inputs(B, cfg["H"], cfg["W"], seed=0, Cs=cfg["cs"])
rgb, normal, depth = rgb.to(dev), normal.to(dev), depth.to(dev)
with torch.no_grad():
    for _ in range(3):
        bench.call_generator(G, cfg, rgb, normal, depth)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
    bench.call_generator(G, cfg, rgb, normal, depth)
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:90]
        r = rows.setdefault(n, [0.0, 0]); r[0] += e.device_time; r[1] += 1
tot = sum(r[0] for r in rows.values())
print(f"{os.environ.get('CONFIG', 'c2')} B={B}: {tot/1e3:.2f} ms of GPU kernel time per forward (lanes overlap: wall time is shorter)")
for n, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
