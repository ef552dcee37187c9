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
    if os.environ.get("OFFSCALE"):          # how does the band halo behave with larger offsets?
        plan.offset.mul_(float(os.environ["OFFSCALE"]))
        print("offset std now", float(plan.offset.std()))

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




def train_timing(B=8):
    """fused forward + backward of the propagation against the reference's composition (prop_time DCN Function calls):
        python scripts/prof_nlspn.py --train [B]"""
    from rdfc_gan_b200.nlspn import NLSPNRefineModule
    from _synth import nlspn_stress_inputs
    H, W = bench.H, bench.W
    x = nlspn_stress_inputs(B, H, W, 1)
    t = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    gout = torch.randn(B, 1, H, W, device="cuda")
    for fused in (True, False):
        mod = NLSPNRefineModule(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True).cuda().train()
        mod.prop_layer.conv_offset_aff.weight.data.copy_(t["conv_w"])
        mod.prop_layer.conv_offset_aff.bias.data.copy_(t["conv_b"])
        mod.prop_layer.fused_backward = fused
        ts = []
        for rep in range(5):
            g = t["guidance"].clone().requires_grad_(True)
            p = t["pred_init"].clone().requires_grad_(True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y, _ = mod(p, g, t["confidence"], t["feat_fix"])
            y.backward(gout)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"NLSPN forward+backward B={B} 228x304 x18, {'fused propagate fwd/bwd' if fused else 'composition (18 DCN Function calls)'}: "
              f"{statistics.median(ts[1:]):.2f} ms")


if __name__ == "__main__" and "--train" in sys.argv:
    sys.argv.remove("--train")
    train_timing(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
    sys.exit(0)

if __name__ == "__main__":
    main()
