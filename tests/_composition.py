"""Test infrastructure: NLSPN as a composition of per-call DCN Functions on the GPU (what the reference's nlspn_model.py does
through its extension), built on this repo's general DCN kernels (rdfc_dcn_forward / backward via
rdfc_gan_b200.dcn.functions).  The product no longer contains such a path -- it runs the fused kernels only -- so the tests that
pin the fused forward / backward pair against "26 separate DCN calls + ATen glue" build the composition here."""
import torch

from rdfc_gan_b200.dcn.functions import ModulatedDeformConvFunction


def offset_affinity(pl, guidance, confidence):
    """pl: an NLPSN module (parameters only).  -> offset (B,18,H,W), aff (B,9,H,W) with autograd through every step."""
    B, _, H, W = guidance.shape
    raw = pl.conv_offset_aff(guidance)
    # neighbour j takes conv channels (2j, 2j+1) as (dy, dx); the centre tap (index 4) has zero offset
    pairs = raw[:, :16].reshape(B, 8, 2, H, W)
    offset = torch.cat([pairs[:, :4], pairs.new_zeros(B, 1, 2, H, W), pairs[:, 4:]], 1).reshape(B, 18, H, W)
    aff = raw[:, 16:]
    if pl.affinity == 'TC':
        aff = torch.tanh(aff) / pl.aff_scale_const
    elif pl.affinity == 'TGASS':
        aff = torch.tanh(aff) / (pl.aff_scale_const + 1e-8)
    if pl.conf_prop:
        ones = torch.ones(B, 1, H, W, device=guidance.device)
        taps = [k for k in range(9) if k != 4]
        gathered = [ModulatedDeformConvFunction.apply(confidence, offset[:, 2 * k:2 * k + 2].detach().contiguous(), ones, pl.w_conf, pl.b,
                                                      1, 0, 1, 1, 1, 64) for k in taps]
        aff = aff * torch.cat(gathered, 1)
    total = aff.abs().sum(1, keepdim=True) + 1e-4
    if pl.affinity in ('ASS', 'TGASS'):
        total = total.clamp(min=1.0)
    if pl.affinity != 'TC':
        aff = aff / total
    centre = 1.0 - aff.sum(1, keepdim=True)
    return offset, torch.cat([aff[:, :4], centre, aff[:, 4:]], 1)


def refine(mod, pred_init, guidance, confidence, feat_fix):
    """NLSPNRefineModule.forward as a composition -> (result, [every iteration's result])."""
    pl = mod.prop_layer
    offset, aff = offset_affinity(pl, guidance, confidence)
    feat, steps = pred_init, []
    keep = (feat_fix > 0).float() if pl.preserve_input else None
    for _ in range(pl.prop_time):
        if keep is not None:
            feat = (1.0 - keep) * feat + keep * feat_fix
        feat = ModulatedDeformConvFunction.apply(feat, offset, aff, pl.w, pl.b, 1, 1, 1, 1, 1, 64)
        steps.append(feat)
    return feat, steps
