"""NLSPN refinement -- drop-in for nlspn/nlspn_model.py (NLPSN :6-175, NLSPNRefineModule :178-197).

Same constructor arguments, parameter names / shapes (state_dict compatible: ``aff_scale_const``, ``w``, ``b``,
``w_conf``, ``conv_offset_aff.{weight,bias}``), asserts and return tuples.  The forward pass runs two fused sm_100a
kernels through the C ABI (rdfc_nlspn_affinity_forward, rdfc_nlspn_propagate_forward) instead of the reference's
26 DCN calls + ~25 ATen kernels.  When gradients are required it falls back to the reference's *composition*
(still on the GPU, through ModulatedDeformConvFunction -> rdfc_dcn_forward/backward); there is no CPU path.
"""
import torch
import torch.nn as nn

from . import _cabi as C
from .dcn.functions import ModulatedDeformConvFunction


class _PropagateFused(torch.autograd.Function):
    """The T propagation iterations as one differentiable op: forward = rdfc_nlspn_propagate_forward (keeping every
    iteration's result), backward = rdfc_nlspn_propagate_backward (the transposed gather as a scatter per iteration, then ONE
    pass for grad_offset / grad_aff).  Replaces the reference's prop_time ModulatedDeformConvFunction calls and their
    backwards (nlspn_model.py:140-175); feat_fix gets no gradient path worth keeping (the reference detaches the mask only,
    the m * fix term's gradient is returned as well)."""

    @staticmethod
    def forward(ctx, feat_init, offset, aff, feat_fix, prop_time, preserve_input):
        B, _, H, W = feat_init.shape
        feat_init, offset, aff = feat_init.contiguous(), offset.contiguous(), aff.contiguous()
        fix = feat_fix.contiguous() if (preserve_input and feat_fix is not None) else None
        out = torch.empty_like(feat_init)
        scratch = torch.empty_like(feat_init)
        inter = torch.empty((max(prop_time, 1),) + tuple(feat_init.shape), dtype=torch.float32, device=feat_init.device)
        with torch.cuda.device(feat_init.device):
            C.check(C.lib.rdfc_nlspn_propagate_forward(
                C.ptr(feat_init), C.ptr(offset), C.ptr(aff), C.ptr(fix), int(bool(preserve_input)), C.ptr(out),
                C.ptr(scratch), C.ptr(inter), B, H, W, prop_time, 0, C.stream_ptr(feat_init.device)))
        ctx.save_for_backward(feat_init, offset, aff, fix if fix is not None else feat_init.new_empty(0), inter)
        ctx.cfg = (prop_time, bool(preserve_input), fix is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat_init, offset, aff, fix, inter = ctx.saved_tensors
        prop_time, preserve, has_fix = ctx.cfg
        B, _, H, W = feat_init.shape
        grad_out = grad_out.contiguous()
        g_feat, g_off, g_aff = torch.empty_like(feat_init), torch.empty_like(offset), torch.empty_like(aff)
        scratch = torch.empty((prop_time + 1,) + tuple(feat_init.shape), dtype=torch.float32, device=feat_init.device)
        with torch.cuda.device(feat_init.device):
            C.check(C.lib.rdfc_nlspn_propagate_backward(
                C.ptr(grad_out), None, C.ptr(feat_init), C.ptr(inter), C.ptr(offset), C.ptr(aff),
                C.ptr(fix) if has_fix else None, int(preserve), C.ptr(g_feat), C.ptr(g_off), C.ptr(g_aff), C.ptr(scratch),
                B, H, W, prop_time, C.stream_ptr(feat_init.device)))
        g_fix = None
        if has_fix and ctx.needs_input_grad[3]:
            raise NotImplementedError("gradient w.r.t. feat_fix (the sparse input depth) is not provided by the fused NLSPN backward")
        return g_feat, g_off, g_aff, g_fix, None, None


class _AffinityFused(torch.autograd.Function):
    """Offsets + affinities (nlspn_model.py:68-138) as one differentiable op: forward = rdfc_nlspn_affinity_forward, backward =
    rdfc_nlspn_affinity_backward (tanh / scale, confidence product with its transposed 1x1 DCN gathers, abs-sum normalisation,
    aff_ref) followed by the backward of the 8 -> 24 channel conv_offset_aff, which is left to cuDNN (0.24 GFLOP per image)."""

    @staticmethod
    def forward(ctx, guidance, confidence, weight, bias, aff_scale, affinity, conf_prop):
        B, _, H, W = guidance.shape
        guidance, weight, bias = guidance.contiguous(), weight.contiguous(), bias.contiguous()
        conf = confidence.contiguous() if (conf_prop and confidence is not None) else None
        offset = torch.empty((B, 18, H, W), dtype=torch.float32, device=guidance.device)
        aff = torch.empty((B, 9, H, W), dtype=torch.float32, device=guidance.device)
        scale = aff_scale.detach().contiguous()
        with torch.cuda.device(guidance.device):
            C.check(C.lib.rdfc_nlspn_affinity_forward(
                C.ptr(guidance), C.ptr(conf), C.ptr(weight), C.ptr(bias), C.ptr(scale), C.AFFINITY[affinity], int(bool(conf_prop)),
                C.ptr(offset), C.ptr(aff), B, H, W, C.stream_ptr(guidance.device)))
        ctx.save_for_backward(guidance, conf if conf is not None else guidance.new_empty(0), weight, bias, scale, offset)
        ctx.cfg = (affinity, bool(conf_prop), conf is not None)
        return offset, aff

    @staticmethod
    def backward(ctx, g_offset, g_aff):
        guidance, conf, weight, bias, scale, offset = ctx.saved_tensors
        affinity, conf_prop, has_conf = ctx.cfg
        B, _, H, W = guidance.shape
        g_conv = torch.empty((B, 24, H, W), dtype=torch.float32, device=guidance.device)
        g_conf = torch.empty((B, 1, H, W), dtype=torch.float32, device=guidance.device) if has_conf else None
        g_scale = torch.empty((1,), dtype=torch.float32, device=guidance.device) if affinity == 'TGASS' else None
        with torch.cuda.device(guidance.device):
            C.check(C.lib.rdfc_nlspn_affinity_backward(
                C.ptr(guidance), C.ptr(conf) if has_conf else None, C.ptr(weight), C.ptr(bias), C.ptr(scale), C.AFFINITY[affinity],
                int(conf_prop and has_conf), C.ptr(offset), C.ptr(g_offset.contiguous()), C.ptr(g_aff.contiguous()), C.ptr(g_conv),
                C.ptr(g_conf), C.ptr(g_scale), B, H, W, C.stream_ptr(guidance.device)))
        need = ctx.needs_input_grad
        g_guid = torch.nn.grad.conv2d_input(guidance.shape, weight, g_conv, padding=1) if need[0] else None
        g_w = torch.nn.grad.conv2d_weight(guidance, weight.shape, g_conv, padding=1) if need[2] else None
        g_b = g_conv.sum((0, 2, 3)) if need[3] else None
        return (g_guid, g_conf if (has_conf and need[1]) else None, g_w, g_b,
                g_scale if (g_scale is not None and need[4]) else None, None, None)


class NLPSN(nn.Module):
    def __init__(self, channels_g, channels_f, k_g, k_f, prop_time=1, affinity=None, affinity_gamma=0.5,
                 conf_prop=True, preserve_input=False):
        super().__init__()
        self.conf_prop = conf_prop
        self.preserve_input = preserve_input
        assert channels_f == 1, 'only tested with channels_f == 1 but {}'.format(channels_g)
        assert (k_g % 2) == 1, 'only odd kernel is supported but k_g = {}'.format(k_g)
        assert (k_f % 2) == 1, 'only odd kernel is supported but k_g = {}'.format(k_f)
        pad_g = int((k_g - 1) / 2)
        pad_f = int((k_f - 1) / 2)
        self.prop_time = prop_time
        self.affinity = affinity
        self.channels_g = channels_g
        self.channels_f = channels_f
        self.k_g = k_g
        self.k_f = k_f
        self.num = self.k_f * self.k_f - 1          # nlspn_model.py:32
        self.idx_ref = self.num // 2
        if self.affinity in ['AS', 'ASS', 'TC', 'TGASS']:
            self.conv_offset_aff = nn.Conv2d(self.channels_g, 3 * self.num, kernel_size=self.k_g, stride=1,
                                             padding=pad_g, bias=True)
            self.conv_offset_aff.weight.data.zero_()    # :37-38
            self.conv_offset_aff.bias.data.zero_()
            if self.affinity == 'TC':
                self.aff_scale_const = nn.Parameter(self.num * torch.ones(1))
                self.aff_scale_const.requires_grad = False
            elif self.affinity == 'TGASS':
                self.aff_scale_const = nn.Parameter(affinity_gamma * self.num * torch.ones(1))
            else:
                self.aff_scale_const = nn.Parameter(torch.ones(1))
                self.aff_scale_const.requires_grad = False
        else:
            raise NotImplementedError
        # dummy gather parameters, kept for state_dict compatibility (:52-59)
        self.w = nn.Parameter(torch.ones((self.channels_f, 1, self.k_f, self.k_f)))
        self.b = nn.Parameter(torch.zeros(self.channels_f))
        self.w.requires_grad = False
        self.b.requires_grad = False
        self.w_conf = nn.Parameter(torch.ones(1, 1, 1, 1))
        self.w_conf.requires_grad = False
        self.stride = 1
        self.padding = pad_f
        self.dilation = 1
        self.groups = self.channels_f
        self.deformable_groups = 1
        self.im2col_step = 64
        self.return_intermediates = True
        self.fused_backward = True          # False: always differentiate through the reference's composition

    # ---- fused inference path --------------------------------------------------------------------------------
    def _fused_ok(self, *tensors):
        if self.k_f != 3 or self.k_g != 3:
            return False
        if torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors) or
                                        any(p.requires_grad for p in self.conv_offset_aff.parameters())):
            return False
        return all(t is None or t.dtype == torch.float32 for t in tensors)

    def _fused_train_ok(self, feat_init, offset, aff, feat_fix):
        """Gradients are needed: the fused forward + backward pair covers k_f = 3, fp32, no list_feat consumer and no
        gradient into feat_fix; everything else takes the reference's composition below."""
        if self.k_f != 3 or self.return_intermediates or not self.fused_backward:
            return False
        if self.preserve_input and feat_fix is not None and feat_fix.requires_grad:
            return False
        return all(t is None or t.dtype == torch.float32 for t in (feat_init, offset, aff, feat_fix))

    def _get_offset_affinity_fused(self, guidance, confidence):
        B, _, H, W = guidance.shape
        guidance = guidance.contiguous()
        confidence = None if confidence is None else confidence.contiguous()
        offset = torch.empty((B, 2 * (self.num + 1), H, W), dtype=torch.float32, device=guidance.device)
        aff = torch.empty((B, self.num + 1, H, W), dtype=torch.float32, device=guidance.device)
        with torch.cuda.device(guidance.device):
            C.check(C.lib.rdfc_nlspn_affinity_forward(
                C.ptr(guidance), C.ptr(confidence), C.ptr(self.conv_offset_aff.weight.detach().contiguous()),
                C.ptr(self.conv_offset_aff.bias.detach().contiguous()), C.ptr(self.aff_scale_const.detach()),
                C.AFFINITY[self.affinity], int(bool(self.conf_prop)), C.ptr(offset), C.ptr(aff), B, H, W,
                C.stream_ptr(guidance.device)))
        return offset, aff

    def _propagate_fused(self, feat_init, offset, aff, feat_fix, want_inter):
        B, _, H, W = feat_init.shape
        feat_init = feat_init.contiguous()
        out = torch.empty_like(feat_init)
        scratch = torch.empty_like(feat_init)
        inter = (torch.empty((self.prop_time,) + tuple(feat_init.shape), dtype=torch.float32, device=feat_init.device)
                 if want_inter and self.prop_time > 0 else None)
        fix = feat_fix.contiguous() if (self.preserve_input and feat_fix is not None) else None
        with torch.cuda.device(feat_init.device):
            C.check(C.lib.rdfc_nlspn_propagate_forward(
                C.ptr(feat_init), C.ptr(offset), C.ptr(aff), C.ptr(fix), int(bool(self.preserve_input)), C.ptr(out),
                C.ptr(scratch), C.ptr(inter), B, H, W, self.prop_time, 0, C.stream_ptr(feat_init.device)))
        return out, ([] if inter is None else list(inter.unbind(0)))

    # ---- reference composition (autograd) ----------------------------------------------------------------------
    def _get_offset_affinity(self, guidance, confidence=None, rgb=None):
        """nlspn_model.py:68-138"""
        C.require_cuda(guidance, confidence)
        if self._fused_ok(guidance, confidence):
            return self._get_offset_affinity_fused(guidance, confidence if self.conf_prop else None)
        if (self.fused_backward and self.k_f == 3 and self.k_g == 3 and guidance.dtype == torch.float32 and
                (confidence is None or confidence.dtype == torch.float32)):
            # training: one differentiable op (see _AffinityFused)
            return _AffinityFused.apply(guidance, confidence if self.conf_prop else None, self.conv_offset_aff.weight,
                                        self.conv_offset_aff.bias, self.aff_scale_const, self.affinity, self.conf_prop)
        B, _, H, W = guidance.shape
        offset_aff = self.conv_offset_aff(guidance)
        o1, o2, aff = torch.chunk(offset_aff, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1).view(B, self.num, 2, H, W)
        list_offset = list(torch.chunk(offset, self.num, dim=1))
        list_offset.insert(self.idx_ref, torch.zeros((B, 1, 2, H, W)).type_as(offset))
        offset = torch.cat(list_offset, dim=1).view(B, -1, H, W)
        if self.affinity in ['AS', 'ASS']:
            pass
        elif self.affinity == 'TC':
            aff = torch.tanh(aff) / self.aff_scale_const
        elif self.affinity == 'TGASS':
            aff = torch.tanh(aff) / (self.aff_scale_const + 1e-8)
        else:
            raise NotImplementedError
        if self.conf_prop:
            list_conf = []
            offset_each = torch.chunk(offset, self.num + 1, dim=1)
            modulation_dummy = torch.ones((B, 1, H, W)).type_as(offset).detach()
            for idx_off in range(0, self.num + 1):
                ww = idx_off % self.k_f
                hh = idx_off // self.k_f
                if ww == (self.k_f - 1) / 2 and hh == (self.k_f - 1) / 2:
                    continue
                offset_tmp = offset_each[idx_off].detach()
                conf_tmp = ModulatedDeformConvFunction.apply(confidence, offset_tmp, modulation_dummy, self.w_conf,
                                                             self.b, self.stride, 0, self.dilation, self.groups,
                                                             self.deformable_groups, self.im2col_step)
                list_conf.append(conf_tmp)
            conf_aff = torch.cat(list_conf, dim=1)
            aff = aff * conf_aff.contiguous()
        aff_abs = torch.abs(aff)
        aff_abs_sum = torch.sum(aff_abs, dim=1, keepdim=True) + 1e-4
        if self.affinity in ['ASS', 'TGASS']:
            aff_abs_sum = torch.clamp(aff_abs_sum, min=1.0)     # == aff_abs_sum[aff_abs_sum < 1.0] = 1.0 (:126)
        if self.affinity in ['AS', 'ASS', 'TGASS']:
            aff = aff / aff_abs_sum
        aff_sum = torch.sum(aff, dim=1, keepdim=True)
        aff_ref = 1.0 - aff_sum
        list_aff = list(torch.chunk(aff, self.num, dim=1))
        list_aff.insert(self.idx_ref, aff_ref)
        aff = torch.cat(list_aff, dim=1)
        return offset, aff

    def _propagate_once(self, feat, offset, aff):
        """nlspn_model.py:140-144"""
        return ModulatedDeformConvFunction.apply(feat, offset, aff, self.w, self.b, self.stride, self.padding,
                                                 self.dilation, self.groups, self.deformable_groups, self.im2col_step)

    def forward(self, feat_init, guidance, confidence=None, feat_fix=None, rgb=None):
        """nlspn_model.py:146-175 -> (feat_result, list_feat, offset, aff, aff_scale_const.data)"""
        assert self.channels_g == guidance.shape[1]
        assert self.channels_f == feat_init.shape[1]
        C.require_cuda(feat_init, guidance, confidence)
        if self.conf_prop:
            assert confidence is not None
            offset, aff = self._get_offset_affinity(guidance, confidence, rgb)
        else:
            offset, aff = self._get_offset_affinity(guidance, None, rgb)
        if self.preserve_input:
            assert feat_init.shape == feat_fix.shape
        if self._fused_ok(feat_init, offset, aff):
            feat_result, list_feat = self._propagate_fused(feat_init, offset.contiguous(), aff.contiguous(), feat_fix,
                                                           self.return_intermediates)
            return feat_result, list_feat, offset, aff, self.aff_scale_const.data
        if self._fused_train_ok(feat_init, offset, aff, feat_fix):
            # training: the propagation as ONE differentiable op with a fused backward (list_feat is not returned, as
            # NLSPNRefineModule drops it anyway, nlspn_model.py:193-197)
            feat_result = _PropagateFused.apply(feat_init, offset, aff, feat_fix if self.preserve_input else None,
                                                self.prop_time, self.preserve_input)
            return feat_result, [], offset, aff, self.aff_scale_const.data
        if self.preserve_input:
            mask_fix = torch.sum(feat_fix > 0.0, dim=1, keepdim=True).detach()
            mask_fix = (mask_fix > 0.0).type_as(feat_fix)
        feat_result = feat_init
        list_feat = []
        for k in range(1, self.prop_time + 1):
            if self.preserve_input:
                feat_result = (1.0 - mask_fix) * feat_result + mask_fix * feat_fix
            feat_result = self._propagate_once(feat_result, offset, aff)
            list_feat.append(feat_result)
        return feat_result, list_feat, offset, aff, self.aff_scale_const.data


class NLSPNRefineModule(nn.Module):
    """nlspn_model.py:178-197"""

    def __init__(self, prop_kernel=3, prop_time=18, affinity='TGASS', affinity_gamma=0.5, conf_prop=True,
                 preserve_input=False):
        super().__init__()
        self.num_neighbors = prop_kernel * prop_kernel - 1
        self.conf_prop = conf_prop
        self.prop_layer = NLPSN(channels_g=self.num_neighbors, channels_f=1, k_g=3, k_f=prop_kernel,
                                prop_time=prop_time, affinity=affinity, affinity_gamma=affinity_gamma,
                                conf_prop=conf_prop, preserve_input=preserve_input)
        self.prop_layer.return_intermediates = False    # the wrapper drops y_inter (:193-197)

    def forward(self, init_pred_depth, guide, confidence, origin_depth, origin_rgb=None):
        y, y_inter, offset, aff, aff_const = self.prop_layer(init_pred_depth, guide, confidence, origin_depth,
                                                             origin_rgb)
        return y, confidence
