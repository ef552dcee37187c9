"""Development aid: compare conv_umma pair mode (cta_group::2) against the single-CTA path on a few shapes."""
import ctypes, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdfc_gan_b200 import _cabi as C

def run(B, Cin, Cout, H, W, k, stride, transposed, pair):
    os.environ["RDFC_UMMA_PAIR"] = "1" if pair else "0"
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    Ho, Wo = (2 * H, 2 * W) if transposed else ((H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1)
    out = torch.zeros(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    CoutP = (Cout + 15) // 16 * 16
    w = (torch.randn(k * k, Cin // 8, CoutP, 8, device="cuda", generator=g) / (Cin * k * k) ** 0.5).bfloat16()
    sc, sh = torch.ones(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    d = C.ConvDesc()
    d.B, d.Hi, d.Wi, d.Ho, d.Wo, d.kh, d.kw, d.stride, d.pad = B, H, W, Ho, Wo, k, k, stride, (1 if transposed else k // 2)
    d.transposed, d.act, d.path = transposed, 1, C.PATH_UMMA_BF16
    d.inp, d.in2, d.out, d.residual = C.view(x, Cin, 0), C.view(None), C.view(out), C.view(None)
    d.weight, d.scale, d.shift = w.data_ptr(), sc.data_ptr(), sh.data_ptr()
    C.check(C.lib.rdfc_conv_forward(ctypes.byref(d), C.stream_ptr()))
    torch.cuda.synchronize()
    return out.float()

for cfg in [(1, 32, 64, 16, 16, 1, 1, 0), (1, 64, 64, 16, 16, 1, 1, 0), (1, 32, 64, 16, 32, 1, 1, 0), (1, 32, 128, 16, 16, 1, 1, 0), (2, 64, 128, 40, 40, 3, 2, 0), (1, 64, 128, 32, 32, 1, 2, 0)]:
    a, b = run(*cfg, pair=False), run(*cfg, pair=True)
    diff = (a - b).abs()
    bad = torch.nonzero(diff > 1e-2)
    if len(bad):
        ys, xs, cs = bad[:, 1], bad[:, 2], bad[:, 3]
        print("   bad y", sorted(set(ys.tolist()))[:40], "\n   bad x", sorted(set(xs.tolist()))[:40], "\n   bad c", sorted(set(cs.tolist()))[:70])
        print("   sample a", a[0, 0, 0, :8].tolist(), "\n   sample b", b[0, 0, 0, :8].tolist())
    print(cfg, "max diff", float(diff.max()), "nan" if torch.isnan(b).any() else "", "first bad", bad[0].tolist() if len(bad) else None, "n bad", len(bad), "of", a.numel())
