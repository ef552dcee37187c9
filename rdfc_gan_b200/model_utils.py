"""Parameter containers of the RGB<-depth fusion layers -- state_dict-compatible with model_utils.py:7-129
(EqualLR / EqualLinear: ``style.linear.{bias,weight_orig}``; AdaptiveInstanceNorm = W-AdaIN; AdaIN; IN).
The arithmetic runs in rdfc_gan_b200.engine (rdfc_conv_forward for the per-pixel Linear, rdfc_instnorm_stats,
rdfc_wadain_apply / rdfc_adain_apply / rdfc_norm_apply)."""
from math import sqrt

import torch
import torch.nn as nn

from .encoder_decoder import _no_forward


class _EqualLinearParams(nn.Module):
    """nn.Linear after EqualLR.apply (model_utils.py:17-26): ``bias`` first, then ``weight_orig``."""
    forward = _no_forward

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.in_features, self.out_features = in_dim, out_dim
        self.bias = nn.Parameter(torch.zeros(out_dim))
        self.weight_orig = nn.Parameter(torch.randn(out_dim, in_dim))    # linear.weight.data.normal_() (:45)

    def effective_weight(self):
        """EqualLR.compute_weight (:11-15): weight_orig * sqrt(2 / fan_in)"""
        return self.weight_orig * sqrt(2 / self.in_features)


class EqualLinear(nn.Module):
    forward = _no_forward

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear = _EqualLinearParams(in_dim, out_dim)


class AdaptiveInstanceNorm(nn.Module):
    """W-AdaIN, model_utils.py:53-90"""
    forward = _no_forward

    def __init__(self, in_channel, style_dim, weighting=False):
        super().__init__()
        self.norm = nn.InstanceNorm2d(in_channel)
        self.style = EqualLinear(style_dim, in_channel * 2)
        self.style.linear.bias.data[:in_channel] = 1
        self.style.linear.bias.data[in_channel:] = 0
        self.weighting = bool(weighting)
        if weighting:
            self.gamma_weight_layer = nn.Conv2d(in_channel, in_channel, (1, 1))
            self.beta_weight_layer = nn.Conv2d(in_channel, in_channel, (1, 1))
        self.in_channel, self.style_dim = in_channel, style_dim


class AdaIN(nn.Module):
    """model_utils.py:102-116 (no parameters)"""
    forward = _no_forward


class IN(nn.Module):
    """model_utils.py:119-129"""
    forward = _no_forward

    def __init__(self, in_channel, style_dim):
        super().__init__()
        self.down_channel = nn.Conv2d(in_channel + style_dim, in_channel, (1, 1))
        self.norm = nn.InstanceNorm2d(in_channel + style_dim)
        self.in_channel, self.style_dim = in_channel, style_dim
