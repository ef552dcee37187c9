"""Differentiable building blocks of the TRAINING step on the sm_100a kernels.

What the reference gets from autograd + cuDNN inside ``conv_bn_relu`` / ``convt_bn_relu`` / torchvision ``BasicBlock`` in
train() mode (encoder_decoder/common.py:29-61, encoder_decoder/encoder_decoder.py:39-59) becomes two autograd Functions over
bf16 NHWC tensors:

* ``conv2d_nhwc``   forward = ``rdfc_conv_forward`` (tcgen05 implicit GEMM); data gradient = the same kernel with a transformed
                    filter (flipped / transposed-conv form, conv_grad.py); filter gradient = ``rdfc_conv_wgrad``.
* ``bn_act``        BatchNorm2d with BATCH statistics (+ running-stat update) fused with the residual add and the
                    (Leaky)ReLU: ``rdfc_bn_stats`` + ``rdfc_affine_act_forward``; backward = ``rdfc_bn_act_backward``.

Everything else of the training forward (the 3 / 1-channel stems, the <= 8-channel ``*_dec0`` heads, the W-AdaIN modulation,
losses) is plain differentiable PyTorch on the same tensors -- < 1 % of the FLOPs (DESIGN.md section 7 lists what is not on
the repo's kernels yet).
"""
import ctypes

import torch

from . import _cabi as C
from .conv_grad import pack_filter


def _conv_umma(x, w_oihw, k, stride, transposed, out_hw):
    """x (B,H,W,Cin) bf16 contiguous; w_oihw (O, I, k, k) float -> (B, Ho, Wo, O) bf16, no bias / activation."""
    B, H, W, Cin = x.shape
    O = w_oihw.shape[0]
    if Cin % 32 or O % 8 or w_oihw.shape[1] != Cin:
        raise RuntimeError(f"tensor-core conv needs Cin % 32 == 0 and Cout % 8 == 0, got {Cin} -> {O}")
    out = torch.empty((B, out_hw[0], out_hw[1], O), dtype=torch.bfloat16, device=x.device)
    packed = pack_filter(w_oihw)
    d = C.ConvDesc()
    d.B, d.Hi, d.Wi, d.Ho, d.Wo = B, H, W, out_hw[0], out_hw[1]
    d.kh = d.kw = k
    d.stride, d.pad, d.transposed, d.act = stride, (k - 1) // 2, int(transposed), C.ACT_NONE
    d.path = C.PATH_UMMA_BF16
    d.inp, d.in2, d.out, d.residual = C.view(x), C.view(None), C.view(out), C.view(None)
    d.weight, d.scale, d.shift = packed.data_ptr(), None, None
    with torch.cuda.device(x.device):
        C.check(C.lib.rdfc_conv_forward(ctypes.byref(d), C.stream_ptr(x.device)))
    return out


def _dense(t):
    """The tensor itself when the kernels can read it as a (pointer, channels, pixel stride) view -- dense NHWC or a channel slice of a
    dense NHWC tensor, which is what autograd returns for the operands of a channel concatenation -- else a contiguous copy."""
    return t if C.nhwc_viewable(t) else t.contiguous()


def _wgrad(grad_out, inp, k, stride):
    """sum_p grad_out[p, o] * inp[stride * p - pad + tap, i] -> (O, I, k, k) fp32."""
    B, Hg, Wg, O = grad_out.shape
    _, Hi, Wi, I = inp.shape
    d = C.WgradDesc()
    d.B, d.Hg, d.Wg, d.Hi, d.Wi, d.k, d.stride, d.pad = B, Hg, Wg, Hi, Wi, k, stride, (k - 1) // 2
    d.grad_out, d.input = C.view(grad_out), C.view(inp)
    n = C.lib.rdfc_conv_wgrad_workspace_floats(ctypes.byref(d))
    if n < 0:
        raise RuntimeError(C.lib.rdfc_last_error().decode("utf-8", "replace"))
    ws = torch.empty(n, dtype=torch.float32, device=inp.device)
    gw = torch.empty((O, I, k, k), dtype=torch.float32, device=inp.device)
    with torch.cuda.device(inp.device):
        C.check(C.lib.rdfc_conv_wgrad(ctypes.byref(d), C.ptr(gw), C.ptr(ws), C.stream_ptr(inp.device)))
    return gw


class _Conv2dNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, k, stride, transposed):
        C.require_cuda(x, weight)
        x = _dense(x)
        B, H, W, _ = x.shape
        if transposed:                      # ConvTranspose2d(k3, s2, p1, op1): weight (Cin, Cout, 3, 3), output exactly 2x
            out_hw = (2 * H, 2 * W)
            w_oihw = weight.detach().permute(1, 0, 2, 3)
        else:
            p = (k - 1) // 2
            out_hw = ((H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1)
            w_oihw = weight.detach()
        ctx.save_for_backward(x, weight)
        ctx.cfg = (k, stride, transposed)
        return _conv_umma(x, w_oihw, k, stride, transposed, out_hw)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        k, stride, transposed = ctx.cfg
        gy = _dense(gy)
        B, H, W, Cin = x.shape
        w = weight.detach()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if transposed:                  # dgrad of the transposed conv = the stride-2 conv with the SAME filter (O = Cin, I = Cout)
                gx = _conv_umma(gy, w, 3, 2, False, (H, W))
            elif k == 3 and stride == 1:    # flipped filter, (Cout, Cin) swapped
                gx = _conv_umma(gy, w.flip(2, 3).permute(1, 0, 2, 3), 3, 1, False, (H, W))
            elif k == 3:                    # stride 2: the transposed conv with the same filter, cropped to the input size
                gx = _conv_umma(gy, w.permute(1, 0, 2, 3), 3, 2, True, (H, W))
            else:                           # 1x1: W^T per pixel; stride 2 scatters onto the even grid
                t = _conv_umma(gy, w.permute(1, 0, 2, 3), 1, 1, False, tuple(gy.shape[1:3]))
                if stride == 1:
                    gx = t
                else:
                    gx = torch.zeros_like(x)
                    gx[:, ::2, ::2] = t
        if ctx.needs_input_grad[1]:
            if transposed:                  # roles exchanged: "gradient" = the layer's input, "input" = d(output), stride-2 form
                gw = _wgrad(x, gy, 3, 2)                       # (Cin, Cout, 3, 3) = ConvTranspose2d's weight layout
            else:
                gw = _wgrad(gy, x, k, stride)                  # (Cout, Cin, k, k)
            gw = gw.to(weight.dtype)
        return gx, gw, None, None, None


def conv2d_nhwc(x, weight, k=3, stride=1, transposed=False):
    """bf16 NHWC convolution (3x3 / 1x1, stride 1 / 2, padding (k-1)/2) or ConvTranspose2d(k3, s2, p1, op1) without bias."""
    return _Conv2dNHWC.apply(x, weight, k, stride, transposed)


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, residual, act, eps, bn):
        y = _dense(y)
        B, H, W, Cc = y.shape
        npix = B * H * W
        dev = y.device
        ws = torch.empty(C.lib.rdfc_bn_workspace_floats(npix, Cc), dtype=torch.float32, device=dev)
        mean, var = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
        vy = C.view(y)
        with torch.cuda.device(dev):
            C.check(C.lib.rdfc_bn_stats(ctypes.byref(vy), npix, C.ptr(ws), C.ptr(mean), C.ptr(var), C.stream_ptr(dev)))
        rstd = torch.rsqrt(var + eps)
        scale = (gamma.detach().float() * rstd).contiguous()
        shift = (beta.detach().float() - mean * scale).contiguous()
        out = torch.empty(y.shape, dtype=y.dtype, device=y.device)
        res = _dense(residual) if residual is not None else None
        vo, vr = C.view(out), C.view(res)
        with torch.cuda.device(dev):
            C.check(C.lib.rdfc_affine_act_forward(ctypes.byref(vy), C.ptr(scale), C.ptr(shift), ctypes.byref(vr) if res is not None else None,
                                                  act, ctypes.byref(vo), npix, C.stream_ptr(dev)))
        if bn is not None and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():               # nn.BatchNorm2d: momentum 0.1, UNBIASED variance in the running estimate
                m = bn.momentum if bn.momentum is not None else 0.1
                bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
                bn.running_var.mul_(1 - m).add_(var * (npix / max(npix - 1, 1)), alpha=m)
                bn.num_batches_tracked += 1
        ctx.save_for_backward(y, out, mean, rstd, gamma)
        ctx.cfg = (act, residual is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        y, out, mean, rstd, gamma = ctx.saved_tensors
        act, has_res = ctx.cfg
        gout = _dense(gout)
        B, H, W, Cc = y.shape
        npix = B * H * W
        dev = y.device
        dy = torch.empty(y.shape, dtype=y.dtype, device=y.device)
        dres = torch.empty(y.shape, dtype=y.dtype, device=y.device) if (has_res and ctx.needs_input_grad[3]) else None
        ws = torch.empty(C.lib.rdfc_bn_workspace_floats(npix, Cc), dtype=torch.float32, device=dev)
        s_dz, s_dzx = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
        g32 = gamma.detach().float().contiguous()
        vg, vo, vy, vdy, vdr = C.view(gout), C.view(out), C.view(y), C.view(dy), C.view(dres)
        with torch.cuda.device(dev):
            C.check(C.lib.rdfc_bn_act_backward(ctypes.byref(vg), ctypes.byref(vo), ctypes.byref(vy), C.ptr(mean), C.ptr(rstd), C.ptr(g32), act,
                                               ctypes.byref(vdy), ctypes.byref(vdr) if dres is not None else None, C.ptr(ws), C.ptr(s_dz),
                                               C.ptr(s_dzx), npix, C.stream_ptr(dev)))
        return dy, s_dzx.to(gamma.dtype), s_dz.to(gamma.dtype), dres, None, None, None


def bn_act(y, bn, residual=None, act=C.ACT_NONE):
    """act(BatchNorm2d_train(y) + residual) over a bf16 NHWC tensor; ``bn`` is the nn.BatchNorm2d holding weight / bias and the
    running statistics (updated in place, as nn.BatchNorm2d.forward does in train mode)."""
    return _BNAct.apply(y, bn.weight, bn.bias, residual, act, bn.eps, bn)
