"""Development aid: run one hot kernel in isolation (for CUDA-event timing and ncu captures).
    python scripts/prof_layer.py conv  [B Cin Cout H W k stride transposed]   (bf16 tcgen05 path)
    python scripts/prof_layer.py nlspn [B]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdfc_gan_b200 import _cabi as C  # noqa: E402


def time_it(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def conv(B=8, Cin=64, Cout=64, H=228, W=304, k=3, stride=1, transposed=0, in_stride=None):
    in_stride = in_stride or Cin
    x = torch.randn(B, H, W, in_stride, device="cuda").bfloat16()
    Ho, Wo = (2 * H, 2 * W) if transposed else ((H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1)
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    CoutP = (Cout + 15) // 16 * 16
    w = (torch.randn(k * k, Cin // 8, CoutP, 8, device="cuda") / (Cin * k * k) ** 0.5).bfloat16()
    sc, sh = torch.ones(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    d = C.ConvDesc()
    d.B, d.Hi, d.Wi, d.Ho, d.Wo, d.kh, d.kw, d.stride, d.pad = B, H, W, Ho, Wo, k, k, stride, (1 if transposed else k // 2)
    d.transposed, d.act, d.path = transposed, 1, C.PATH_UMMA_BF16
    d.inp, d.in2, d.out, d.residual = C.view(x, Cin, 0), C.view(None), C.view(out), C.view(None)
    d.weight, d.scale, d.shift = w.data_ptr(), sc.data_ptr(), sh.data_ptr()
    s = C.stream_ptr()
    ms = time_it(lambda: C.check(C.lib.rdfc_conv_forward(ctypes.byref(d), s)))
    taps = k * k / (4 if transposed else 1)
    flops = 2.0 * B * Ho * Wo * Cin * Cout * taps
    if os.environ.get("RDFC_UMMA_DBG") and hasattr(C.lib, "rdfc_dev_umma_timers"):
        import numpy as np
        buf = (ctypes.c_longlong * (148 * 16))()
        C.lib.rdfc_dev_umma_timers(buf, 148 * 16)
        a = np.frombuffer(buf, dtype=np.int64).reshape(148, 16).astype(np.float64)
        a = a[a[:, 0] > 0]
        names = {0: "MMA warp lifetime", 1: "MMA wait ACC_EMPTY", 2: "MMA wait A_FULL", 3: "MMA wait B_FULL", 5: "Bload wait B_EMPTY",
                 6: "prod wait A_EMPTY", 7: "prod issue", 8: "prod wait_group", 9: "epi wait ACC_FULL", 10: "epi work"}
        print(f"role timers, cycles, median over {len(a)} CTAs (thread 0 of each role):")
        for i, nm in names.items():
            print(f"   {nm:20s} {np.median(a[:, i]):10.0f}  ({100*np.median(a[:, i])/np.median(a[:, 0]):5.1f}% of MMA warp lifetime)")
    print(f"conv B={B} {Cin}->{Cout} {H}x{W} k{k} s{stride} T{transposed}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s "
          f"env={ {k_: v for k_, v in os.environ.items() if k_.startswith('RDFC_')} }")


def nlspn(B=32, H=228, W=304, T=18):
    g = torch.Generator(device="cuda").manual_seed(0)
    off = 2.2 * torch.randn(B, 18, H, W, device="cuda", generator=g)
    off[:, 8:10] = 0
    aff = torch.rand(B, 9, H, W, device="cuda", generator=g)
    aff = aff / aff.sum(1, keepdim=True)
    f = torch.randn(B, 1, H, W, device="cuda", generator=g)
    out, scratch = torch.empty_like(f), torch.empty_like(f)
    s = C.stream_ptr()
    ms = time_it(lambda: C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(f), C.ptr(off), C.ptr(aff), None, 0, C.ptr(out),
                                                                     C.ptr(scratch), None, B, H, W, T, 0, None, s)))
    gbs = 116.0 * B * H * W * T / ms / 1e6
    print(f"nlspn B={B} T={T}: {ms*1e3:.1f} us total, {gbs:.0f} GB/s algorithmic ({gbs/6550.4:.3f} of measured HBM peak) "
          f"env={ {k_: v for k_, v in os.environ.items() if k_.startswith('RDFC_')} }")


if __name__ == "__main__":
    what = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    {"conv": conv, "nlspn": nlspn}[what](*args)
