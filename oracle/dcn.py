"""ctypes wrappers over oracle/dcn_oracle.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Signatures mirror the reference's pybind module ``DCN`` (deformconv/src/vision.cpp:7-12,
deformconv/src/modulated_deform_conv.h:10-26,46-63, deformconv/src/deform_conv.h) but take / return
contiguous numpy arrays (float32 or float64).
"""
import ctypes

import numpy as np

from . import build

_lib = None


class _Shape(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("B", "Cin", "H", "W", "Cout", "kh", "kw", "sh", "sw", "ph", "pw", "dh", "dw", "group", "dg")]


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _np(x, dtype=None):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    x = np.ascontiguousarray(x)
    if dtype is not None and x.dtype != dtype:
        x = x.astype(dtype)
    return x


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def out_size(H, W, kh, kw, sh, sw, ph, pw, dh, dw):
    """deformconv/src/cuda/modulated_deform_conv_cuda.cu:75-76"""
    return ((H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1)


def _shape(inp, weight, kh, kw, sh, sw, ph, pw, dh, dw, group, dg):
    B, Cin, H, W = inp.shape
    assert weight.shape[2] == kh and weight.shape[3] == kw and weight.shape[1] * group == Cin
    return _Shape(B, Cin, H, W, weight.shape[0], kh, kw, sh, sw, ph, pw, dh, dw, group, dg)


def modulated_deform_conv_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw,
                                  group, deformable_group, im2col_step=64):
    """DCN.modulated_deform_conv_forward (mask=None gives DCN.deform_conv_forward)."""
    dt = np.float64 if _np(input).dtype == np.float64 else np.float32
    input, weight, bias, offset, mask = (_np(t, dt) for t in (input, weight, bias, offset, mask))
    s = _shape(input, weight, kh, kw, sh, sw, ph, pw, dh, dw, group, deformable_group)
    Ho, Wo = out_size(s.H, s.W, kh, kw, sh, sw, ph, pw, dh, dw)
    assert offset.shape == (s.B, deformable_group * 2 * kh * kw, Ho, Wo), offset.shape
    if mask is not None:
        assert mask.shape == (s.B, deformable_group * kh * kw, Ho, Wo), mask.shape
    out = np.empty((s.B, s.Cout, Ho, Wo), dt)
    fn = lib().orc_dcn_forward_f64 if dt == np.float64 else lib().orc_dcn_forward_f32
    rc = fn(_ptr(input), _ptr(weight), _ptr(bias), _ptr(offset), _ptr(mask), _ptr(out), ctypes.byref(s))
    assert rc == 0
    return out


def deform_conv_forward(input, weight, bias, offset, kh, kw, sh, sw, ph, pw, dh, dw, group, deformable_group,
                        im2col_step=64):
    return modulated_deform_conv_forward(input, weight, bias, offset, None, kh, kw, sh, sw, ph, pw, dh, dw,
                                         group, deformable_group, im2col_step)


def modulated_deform_conv_backward(input, weight, bias, offset, mask, grad_output, kh, kw, sh, sw, ph, pw, dh, dw,
                                   group, deformable_group, im2col_step=64):
    """DCN.modulated_deform_conv_backward -> (grad_input, grad_offset, grad_mask, grad_weight, grad_bias);
    mask=None gives DCN.deform_conv_backward (grad_mask is then None)."""
    dt = np.float64 if _np(input).dtype == np.float64 else np.float32
    input, weight, offset, mask, grad_output = (_np(t, dt) for t in (input, weight, offset, mask, grad_output))
    s = _shape(input, weight, kh, kw, sh, sw, ph, pw, dh, dw, group, deformable_group)
    gi, go, gw = np.empty_like(input), np.empty_like(offset), np.empty_like(weight)
    gm = np.empty((s.B, deformable_group * kh * kw) + offset.shape[2:], dt)
    gb = np.empty((s.Cout,), dt)
    fn = lib().orc_dcn_backward_f64 if dt == np.float64 else lib().orc_dcn_backward_f32
    rc = fn(_ptr(input), _ptr(weight), _ptr(offset), _ptr(mask), _ptr(grad_output), _ptr(gi), _ptr(go), _ptr(gm),
            _ptr(gw), _ptr(gb), ctypes.byref(s))
    assert rc == 0
    return gi, go, (gm if mask is not None else None), gw, gb


def nlspn_propagate(feat_init, offset, aff, feat_fix=None, preserve_input=False, k_f=3, prop_time=18):
    """nlspn_model.py:157-173 (the propagation loop only), float32."""
    feat_init, offset, aff, feat_fix = (_np(t, np.float32) for t in (feat_init, offset, aff, feat_fix))
    B, C, H, W = feat_init.shape
    assert C == 1 and offset.shape == (B, 2 * k_f * k_f, H, W) and aff.shape == (B, k_f * k_f, H, W)
    out = np.empty_like(feat_init)
    scratch = np.empty_like(feat_init)
    rc = lib().orc_nlspn_propagate_f32(_ptr(feat_init), _ptr(offset), _ptr(aff), _ptr(feat_fix),
                                       int(bool(preserve_input)), _ptr(out), _ptr(scratch),
                                       B, H, W, k_f, prop_time)
    assert rc == 0
    return out
