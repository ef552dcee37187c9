"""Development aid: time of the 18-iteration NLSPN propagation alone (CUDA graph of the launches, L2 flushed between reps).
    python scripts/prof_nlspn.py [B]          (RDFC_NLSPN_* knobs apply)
"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from rdfc_gan_b200 import _cabi as C  # noqa: E402
from _synth import synth_inputs  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    G = bench.build_generator().to(dev).set_precision("bf16")
    rgb, normal, depth = synth_inputs(B, bench.H, bench.W, seed=0)
    with torch.no_grad():
        G(rgb.to(dev), depth.to(dev), normal.to(dev))
    plan = next(iter(G.engine()._plans.values()))
    T, H, W = 18, bench.H, bench.W

    def prop():
        C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(plan.pred_init), C.ptr(plan.offset), C.ptr(plan.aff), None, 0,
                                                   C.ptr(plan.d2raw), C.ptr(plan.scratch), None, B, H, W, T, 0, C.stream_ptr()))
    for _ in range(3):
        prop()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        prop()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    print(f"NLSPN x{T} B={B}: {ms*1e3:.1f} us total, {ms*1e3/T:.2f} us / launch, "
          f"{116.0*B*H*W*T/ms/1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
