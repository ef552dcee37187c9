"""Development aid: run the plan of a B-image forward step by step with a device sync after each C-ABI call."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_generator  # noqa: E402
from rdfc_gan_b200 import _cabi as C  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
G = build_generator().cuda().set_precision("bf16")
eng = G.engine()
plan = eng._build_plan(B, 228, 304, 3, torch.device("cuda", 0), "bf16")
plan.stem_in.normal_()
plan.depth.zero_()
s = C.stream_ptr()
names = plan.names + ["?"] * (len(plan.steps) - len(plan.names))
ci = 0
for i, f in enumerate(plan.steps):
    t0 = time.perf_counter()
    try:
        f(s)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at step", i, e)
        break
    print(f"step {i:3d} ok {1e3 * (time.perf_counter() - t0):8.3f} ms", flush=True)
print("names:")
for n in plan.names:
    print("  ", n)
