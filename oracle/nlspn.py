"""numpy restatement of the reference NLSPN (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows /root/reference/RDFC-GAN/lib/models/generator/rdf_generator/nlspn/nlspn_model.py line by line:
``get_offset_affinity`` = NLPSN._get_offset_affinity (:68-138), ``nlspn_forward`` = NLPSN.forward (:146-175).
Every ModulatedDeformConvFunction.apply of the reference becomes one call into oracle/dcn.py with the same
arguments (weights of ones, zero bias, the same stride / padding / groups).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import dcn


def get_offset_affinity(guidance, confidence, conv_w, conv_b, aff_scale_const, *, k_f=3, affinity="TGASS",
                        conf_prop=True):
    """guidance (B, k_f^2-1, H, W), confidence (B,1,H,W) or None -> offset (B, 2 k_f^2, H, W), aff (B, k_f^2, H, W)."""
    guidance = np.asarray(guidance, np.float32)
    B, _, H, W = guidance.shape
    num = k_f * k_f - 1
    idx_ref = num // 2
    pad_g = (conv_w.shape[-1] - 1) // 2
    # :72 conv_offset_aff (k_g x k_g conv, CPU fp32)
    offset_aff = F.conv2d(torch.from_numpy(guidance), torch.as_tensor(conv_w, dtype=torch.float32),
                          torch.as_tensor(conv_b, dtype=torch.float32), stride=1, padding=pad_g).numpy()
    o1, o2, aff = np.split(offset_aff, 3, axis=1)                                      # :73
    offset = np.concatenate((o1, o2), axis=1).reshape(B, num, 2, H, W)                 # :76 (consecutive pairs)
    offset = np.concatenate((offset[:, :idx_ref], np.zeros((B, 1, 2, H, W), np.float32), offset[:, idx_ref:]),
                            axis=1).reshape(B, -1, H, W)                               # :77-80
    scale = np.float32(np.asarray(aff_scale_const, np.float32).reshape(-1)[0])
    if affinity in ("AS", "ASS"):                                                      # :82-83
        pass
    elif affinity == "TC":                                                             # :84-85
        aff = np.tanh(aff) / scale
    elif affinity == "TGASS":                                                          # :86-87
        aff = np.tanh(aff) / (scale + np.float32(1e-8))
    else:
        raise NotImplementedError(affinity)
    aff = aff.astype(np.float32)

    if conf_prop:                                                                      # :96-119
        ones_w = np.ones((1, 1, 1, 1), np.float32)
        zero_b = np.zeros((1,), np.float32)
        modulation_dummy = np.ones((B, 1, H, W), np.float32)
        list_conf = []
        for idx_off in range(num + 1):
            ww, hh = idx_off % k_f, idx_off // k_f
            if ww == (k_f - 1) / 2 and hh == (k_f - 1) / 2:
                continue
            # intended (contiguous) semantics of offset_each[idx_off] -- SURVEY.md section 8 quirk 5
            offset_tmp = np.ascontiguousarray(offset[:, 2 * idx_off:2 * idx_off + 2])
            list_conf.append(dcn.modulated_deform_conv_forward(
                confidence, ones_w, zero_b, offset_tmp, modulation_dummy, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 64))
        aff = aff * np.concatenate(list_conf, axis=1)

    aff_abs_sum = np.sum(np.abs(aff), axis=1, keepdims=True, dtype=np.float32) + np.float32(1e-4)   # :122-123
    if affinity in ("ASS", "TGASS"):                                                   # :125-126
        aff_abs_sum[aff_abs_sum < 1.0] = 1.0
    if affinity in ("AS", "ASS", "TGASS"):                                             # :128-129
        aff = aff / aff_abs_sum
    aff_ref = np.float32(1.0) - np.sum(aff, axis=1, keepdims=True, dtype=np.float32)   # :131-132
    aff = np.concatenate((aff[:, :idx_ref], aff_ref, aff[:, idx_ref:]), axis=1).astype(np.float32)  # :134-136
    return np.ascontiguousarray(offset, np.float32), np.ascontiguousarray(aff)


def nlspn_forward(feat_init, guidance, confidence, feat_fix, conv_w, conv_b, aff_scale_const, *, k_f=3,
                  prop_time=18, affinity="TGASS", conf_prop=True, preserve_input=False, return_inter=False):
    """NLPSN.forward (:146-175) -> (feat_result, offset, aff[, list_feat])."""
    offset, aff = get_offset_affinity(guidance, confidence if conf_prop else None, conv_w, conv_b, aff_scale_const,
                                      k_f=k_f, affinity=affinity, conf_prop=conf_prop)
    feat_init = np.asarray(feat_init, np.float32)
    if not return_inter:
        out = dcn.nlspn_propagate(feat_init, offset, aff, feat_fix, preserve_input, k_f, prop_time)
        return out, offset, aff
    pad = (k_f - 1) // 2
    ones_w = np.ones((1, 1, k_f, k_f), np.float32)
    zero_b = np.zeros((1,), np.float32)
    feat = feat_init
    if preserve_input:                                                                 # :157-160
        mask_fix = (np.asarray(feat_fix) > 0.0).astype(np.float32)
    inter = []
    for _ in range(prop_time):                                                         # :166-173
        if preserve_input:
            feat = (np.float32(1.0) - mask_fix) * feat + mask_fix * np.asarray(feat_fix, np.float32)
        feat = dcn.modulated_deform_conv_forward(feat, ones_w, zero_b, offset, aff, k_f, k_f, 1, 1, pad, pad, 1, 1,
                                                 1, 1, 64)                             # :140-144
        inter.append(feat)
    return feat, offset, aff, inter


def nlspn_propagate_backward(grad_out, feat_init, offset, aff, feat_fix=None, preserve_input=False, prop_time=18,
                             grad_inter=None):
    """Reverse-mode of the propagation loop of nlspn_model.py:157-173 exactly as autograd runs it: per iteration one
    DCN.modulated_deform_conv_backward (modulated_deform_conv_cuda.cu:124-280) with weight = ones(1,1,3,3), bias = 0,
    stride 1, pad 1; the blend feat = (1 - m) * feat + m * fix (:169) passes (1 - m) * grad to the previous iteration.
    Returns (grad_feat_init, grad_offset, grad_aff), float64 accumulation of the per-iteration fp32 results.
    Test infrastructure only."""
    from . import dcn as odcn
    f32 = np.float32
    feat_init, offset, aff = (np.ascontiguousarray(t, f32) for t in (feat_init, offset, aff))
    w, b = np.ones((1, 1, 3, 3), f32), np.zeros((1,), f32)
    m = (np.asarray(feat_fix) > 0).astype(f32) if preserve_input else None
    xs, x = [], feat_init                    # x'_{t-1}: the (blended) input of iteration t
    for _ in range(prop_time):
        if preserve_input:
            x = (1.0 - m) * x + m * np.asarray(feat_fix, f32)
        xs.append(np.ascontiguousarray(x, f32))
        x = odcn.modulated_deform_conv_forward(xs[-1], w, b, offset, aff, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    g = np.array(grad_out, np.float64)
    g_off, g_aff = np.zeros(offset.shape, np.float64), np.zeros(aff.shape, np.float64)
    for t in range(prop_time, 0, -1):
        if grad_inter is not None:
            g = g + np.asarray(grad_inter[t - 1], np.float64)
        gi, go, gm, _, _ = odcn.modulated_deform_conv_backward(xs[t - 1], w, b, offset, aff, g.astype(f32), 3, 3, 1, 1, 1, 1, 1, 1,
                                                               1, 1)
        g_off += go
        g_aff += gm
        g = gi.astype(np.float64)
        if preserve_input:
            g = (1.0 - m) * g
    return g.astype(f32), g_off.astype(f32), g_aff.astype(f32)
