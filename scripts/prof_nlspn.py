"""NLSPN micro-benchmark (development aid): the propagation kernels alone, both stream formats, knob sweeps.

    python scripts/prof_nlspn.py [B] [H W] [--sweep]

Per configuration: the T = 18 launches as a CUDA graph, L2 flushed between repetitions, median / min of 30 repetitions;
prints us per launch and GB/s on the 116 B per pixel and iteration algorithmic figure and on the bytes the kernel
actually streams (fp32 planes 106 B incl. feature read / write, packed fp16 56 B)."""
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rdfc_gan_b200 import _cabi as C  # noqa: E402

T = 18


def bench(fn, flush, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return statistics.median(ts), ts[0]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 32
    H, W = (int(args[1]), int(args[2])) if len(args) >= 3 else (228, 304)
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(0)
    guide = torch.randn(B, 8, H, W, device=dev, generator=gen)
    conf = torch.rand(B, 1, H, W, device=dev, generator=gen)
    init = torch.rand(B, 1, H, W, device=dev, generator=gen) * 2 - 1
    w = torch.cat([0.25 * torch.randn(16, 8, 3, 3, device=dev, generator=gen), 0.02 * torch.randn(8, 8, 3, 3, device=dev, generator=gen)])
    bias = torch.cat([torch.rand(16, device=dev, generator=gen) * 3 - 1.5, torch.rand(8, device=dev, generator=gen) * 1.7 + 0.3])
    w = w * float(os.environ.get("OFFSCALE", "0.32"))          # 0.32: sigma ~ 0.73 px like bench.py's weights; 1.0: sigma ~ 2.2 px
    scale = torch.tensor([4.0], device=dev)
    off, aff = torch.empty(B, 18, H, W, device=dev), torch.empty(B, 9, H, W, device=dev)
    packed = torch.empty(C.lib.rdfc_nlspn_packed_bytes(B, H, W), dtype=torch.uint8, device=dev)
    s = C.stream_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    aff_args = (C.ptr(guide), C.ptr(conf), C.ptr(w), C.ptr(bias), C.ptr(scale), 3, 1)
    f_aff32 = lambda: C.check(C.lib.rdfc_nlspn_affinity_forward(*aff_args, C.ptr(off), C.ptr(aff), B, H, W, C.stream_ptr()))
    f_aff16 = lambda: C.check(C.lib.rdfc_nlspn_affinity_forward_packed(*aff_args, C.ptr(packed), B, H, W, C.stream_ptr()))
    f_aff32(); f_aff16()
    print(f"B={B} {H}x{W}  offset sigma {float(off[:, :8].std()):.2f} px")
    for name, f in (("affinity fp32 planes", f_aff32), ("affinity packed fp16", f_aff16)):
        med, mn = bench(f, flush)
        print(f"{name:28s} {med * 1e3:8.1f} us (min {mn * 1e3:.1f})")
    out, scratch, pred = torch.empty_like(init), torch.empty_like(init), torch.empty_like(init)
    d1, c1 = torch.rand_like(init), torch.rand_like(init)
    fz = C.FuseOut(d1.data_ptr(), c1.data_ptr(), conf.data_ptr(), pred.data_ptr())
    f32 = lambda: C.check(C.lib.rdfc_nlspn_propagate_forward(C.ptr(init), C.ptr(off), C.ptr(aff), None, 0, C.ptr(out), C.ptr(scratch), None,
                                                             B, H, W, T, 0, ctypes.byref(fz), C.stream_ptr()))
    f16 = lambda: C.check(C.lib.rdfc_nlspn_propagate_forward_packed(C.ptr(init), C.ptr(packed), None, 0, C.ptr(out), C.ptr(scratch),
                                                                    B, H, W, T, 0, ctypes.byref(fz), C.stream_ptr()))
    P = B * H * W

    def report(name, f, actual):
        med, mn = bench(f, flush)
        us = med * 1e3 / T
        print(f"{name:44s} {us:7.2f} us/launch (min {mn * 1e3 / T:6.2f})  {116 * P / us / 1e3:7.0f} GB/s on 116 B/px  "
              f"{actual * P / us / 1e3:6.0f} GB/s streamed ({actual} B/px)")
    report("propagate fp32 planes (default)", f32, 108)
    report("propagate packed fp16 (default)", f16, 56)
    if "--sweep" in sys.argv:
        for per_sm in (2, 3, 4, 5, 6):
            for halo in (4, 6):
                C.set_knob("RDFC_NLSPN_PER_SM", per_sm)
                C.set_knob("RDFC_NLSPN_HALO", halo)
                report(f"packed per_sm={per_sm} halo={halo}", f16, 56)
        C.set_knob("RDFC_NLSPN_PER_SM", None)
        C.set_knob("RDFC_NLSPN_HALO", None)
        for pdl in (0, 1):
            C.set_knob("RDFC_NLSPN_PDL", pdl)
            report(f"packed pdl={pdl}", f16, 56)
        C.set_knob("RDFC_NLSPN_PDL", None)


if __name__ == "__main__":
    main()
