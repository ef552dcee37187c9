"""Development aid: ESANet-34 alone (config c2's guidance network, B = 32, 228x304): wall time per forward and GPU time by kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import bench
cfg = bench.CONFIGS["c2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
G = bench.build_product(cfg).cuda().set_precision("bf16")
esa = G.global_guidance_module
x = torch.randn(B, 3, cfg["H"], cfg["W"], device="cuda")
with torch.no_grad():
    for _ in range(3): esa(x)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
with torch.no_grad():
    for a, b in ev:
        a.record(); esa(x); b.record()
torch.cuda.synchronize()
print(f"ESANet forward B={B}: {sorted(a.elapsed_time(b) for a, b in ev)[5]:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
    esa(x); torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:80]
        r = rows.setdefault(n, [0.0, 0]); r[0] += e.device_time; r[1] += 1
tot = sum(r[0] for r in rows.values())
print(f"{tot/1e3:.2f} ms of GPU kernel time")
for n, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
