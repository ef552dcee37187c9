"""Development aid: the general DCNv2 forward at the reference's timing shape (deformconv/test.py:519-530: B=2, 64 -> 128, 128x128,
k3, deformable_groups 2): this repo's tcgen05 path and CUDA-core strip kernel against the reference's own extension
(baseline/_ref/build/DCN.so, im2col + cuBLAS)."""
import os, sys, importlib.util
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdfc_gan_b200 import _cabi as C
from rdfc_gan_b200.dcn import DCN

def timeit(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]

for B in (2, 16):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, 64, 128, 128, device="cuda", generator=g)
    w = torch.randn(128, 64, 3, 3, device="cuda", generator=g) * 0.05
    b = torch.zeros(128, device="cuda")
    off = torch.randn(B, 2 * 18, 128, 128, device="cuda", generator=g)
    m = torch.rand(B, 2 * 9, 128, 128, device="cuda", generator=g)
    geo = (3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 64)
    res = {}
    for tc in (1, 0):
        C.set_knob("RDFC_DCN_TC", tc)
        res[tc] = (timeit(lambda: DCN.modulated_deform_conv_forward(x, w, b, off, m, *geo)), DCN.modulated_deform_conv_forward(x, w, b, off, m, *geo))
    C.set_knob("RDFC_DCN_TC", None)
    line = f"DCNv2 forward B={B} 64->128 128x128 k3 dg2: tcgen05 path {res[1][0]:.3f} ms | strip kernel {res[0][0]:.3f} ms"
    so = os.path.join(ROOT, "baseline", "_ref", "build", "DCN.so")
    if os.path.exists(so):
        spec = importlib.util.spec_from_file_location("DCN", so); ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
        torch.backends.cuda.matmul.allow_tf32 = False
        t_ref = timeit(lambda: ref.modulated_deform_conv_forward(x, w, b, off, m, *geo))
        y_ref = ref.modulated_deform_conv_forward(x, w, b, off, m, *geo)
        line += f" | reference extension {t_ref:.3f} ms | max-abs vs reference: tcgen05 {float((res[1][1] - y_ref).abs().max()):.2e}, strip {float((res[0][1] - y_ref).abs().max()):.2e}"
    print(line)
    # backward (input / offset / mask / weight / bias gradients)
    go = torch.randn(B, 128, 128, 128, device="cuda", generator=g) * 0.01
    t_b = timeit(lambda: DCN.modulated_deform_conv_backward(x, w, b, off, m, go, *geo), reps=10)
    line = f"DCNv2 backward B={B}: this repo {t_b:.3f} ms"
    if os.path.exists(so):
        t_rb = timeit(lambda: ref.modulated_deform_conv_backward(x, w, b, off, m, go, *geo), reps=10)
        mine, theirs = DCN.modulated_deform_conv_backward(x, w, b, off, m, go, *geo), ref.modulated_deform_conv_backward(x, w, b, off, m, go, *geo)
        errs = ", ".join(f"{float((a - r).abs().max() / r.abs().max().clamp_min(1e-30)):.1e}" for a, r in zip(mine, theirs))
        line += f" | reference extension {t_rb:.3f} ms | relative max-abs differences (input, offset, mask, weight, bias): {errs}"
    print(line)
