import ctypes, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RDFC_UMMA_PAIR"] = "1"
from rdfc_gan_b200 import _cabi as C
B, Cin, Cout, H, W, k = 1, 32, 64, 16, 16, 1
g = torch.Generator(device="cuda").manual_seed(0)
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
ref0 = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=k // 2).permute(0, 2, 3, 1).clone()
x0 = x.clone(); wt0 = wt.clone()
torch.cuda.synchronize()
C.check(C.lib.rdfc_conv_forward(ctypes.byref(d), C.stream_ptr()))
torch.cuda.synchronize()
ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=k // 2).permute(0, 2, 3, 1)
o = out.float()
print("x changed:", int((x0 != x).sum()), "wt changed:", int((wt0 != wt).sum()), "ref before vs after:", float((ref0 - ref).abs().max()), "ref0(0,0)[:4]", ref0[0,0,0,:4].tolist(), "x(0,0)[:4]", x[0,0,0,:4].tolist())
ref = ref0
ch = torch.nonzero((x0 != x).flatten()).flatten()
print("x changed flat idx min/max", int(ch.min()), int(ch.max()), "contiguous" if int(ch.max()-ch.min()+1)==len(ch) else "gaps", "ptr x", hex(x.data_ptr()), "out", hex(out.data_ptr()), "wt", hex(wt.data_ptr()), "w", hex(w.data_ptr()), "sc", hex(sc.data_ptr()))
print("values written into x:", x.flatten()[ch[:8]].tolist(), " w changed:", int((w != wt.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin // 8, 8).permute(1, 2, 0, 3)).sum()))
print("plain      bad:", int(((o - ref).abs() > 0.05).sum()))
print("chan halves swapped bad:", int(((torch.cat([o[..., 32:], o[..., :32]], -1) - ref).abs() > 0.05).sum()))
print("x halves swapped bad:", int(((torch.cat([o[:, :, 8:], o[:, :, :8]], 2) - ref).abs() > 0.05).sum()))
print("left tile (x<8) bad:", int(((o - ref)[:, :, :8].abs() > 0.05).sum()), " right tile bad:", int(((o - ref)[:, :, 8:].abs() > 0.05).sum()))
print("chan 0-31 bad:", int(((o - ref)[..., :32].abs() > 0.05).sum()), " chan 32-63 bad:", int(((o - ref)[..., 32:].abs() > 0.05).sum()))
# does any ref pixel row match out pixel (0,0)?
v = o[0, 0, 0]
dist = (ref[0] - v).abs().sum(-1)
print("out(0,0) closest ref pixel:", divmod(int(dist.argmin()), W), float(dist.min()))
v = o[0, 0, 8]
dist = (ref[0] - v).abs().sum(-1)
print("out(0,8) closest ref pixel:", divmod(int(dist.argmin()), W), float(dist.min()))
print("out(0,0)[:6]", o[0,0,0,:6].tolist(), "ref", ref[0,0,0,:6].tolist())
