"""Development aid: repeat one small conv many times in both modes and compare against torch (detects sporadic errors)."""
import ctypes, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdfc_gan_b200 import _cabi as C

def run(B, Cin, Cout, H, W, k, pair, seed):
    os.environ["RDFC_UMMA_PAIR"] = "1" if pair else "0"
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    wt = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).bfloat16()
    w = wt.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin // 8, 8).permute(1, 2, 0, 3).contiguous()
    sc, sh = torch.ones(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    d = C.ConvDesc()
    d.B, d.Hi, d.Wi, d.Ho, d.Wo, d.kh, d.kw, d.stride, d.pad = B, H, W, H, W, k, k, 1, k // 2
    d.transposed, d.act, d.path = 0, 0, C.PATH_UMMA_BF16
    d.inp, d.in2, d.out, d.residual = C.view(x, Cin, 0), C.view(None), C.view(out), C.view(None)
    d.weight, d.scale, d.shift = w.data_ptr(), sc.data_ptr(), sh.data_ptr()
    C.check(C.lib.rdfc_conv_forward(ctypes.byref(d), C.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=k // 2).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    return float(err.max()), int((err > 0.05).sum())

for cfg in [(1, 32, 64, 16, 16, 1), (1, 32, 64, 16, 16, 3), (2, 64, 64, 40, 40, 3)]:
    for pair in (False, True):
        res = [run(*cfg, pair, seed) for seed in range(12)]
        print(cfg, "pair" if pair else "single", "bad runs:", sum(1 for e, n in res if n > 0), "of 12; worst", max(res))
