"""nn.Module API of the vendored DCNv2 package -- mirror of deformconv/modules/modulated_deform_conv.py:14-103 and
deformconv/modules/deform_conv.py:14-99 (same ctor arguments, parameter names, init, zero-initialised offset convs,
``lr_mult`` attribute; ``bias=False`` only freezes the always-allocated bias, as in the reference)."""
import math

import torch
from torch import nn
from torch.nn import init
from torch.nn.modules.utils import _pair

from .functions import DeformConvFunction, ModulatedDeformConvFunction


class _DeformBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels {} must be divisible by groups {}'.format(in_channels, groups))
        if out_channels % groups != 0:
            raise ValueError('out_channels {} must be divisible by groups {}'.format(out_channels, groups))
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.im2col_step = im2col_step
        self.use_bias = bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.Tensor(out_channels))
        self.reset_parameters()
        if not self.use_bias:
            self.bias.requires_grad = False

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def _taps(self):
        return self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]


class ModulatedDeformConv(_DeformBase):
    def forward(self, input, offset, mask):
        assert 2 * self._taps() == offset.shape[1]
        assert self._taps() == mask.shape[1]
        return ModulatedDeformConvFunction.apply(input, offset, mask, self.weight, self.bias, self.stride,
                                                 self.padding, self.dilation, self.groups, self.deformable_groups,
                                                 self.im2col_step)


_ModulatedDeformConv = ModulatedDeformConvFunction.apply


class ModulatedDeformConvPack(ModulatedDeformConv):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True, lr_mult=0.1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         deformable_groups, im2col_step, bias)
        self.conv_offset_mask = nn.Conv2d(self.in_channels, 3 * self._taps(), kernel_size=self.kernel_size,
                                          stride=self.stride, padding=self.padding, bias=True)
        self.conv_offset_mask.lr_mult = lr_mult
        self.init_offset()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, input):
        out = self.conv_offset_mask(input)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return ModulatedDeformConvFunction.apply(input, offset, mask, self.weight, self.bias, self.stride,
                                                 self.padding, self.dilation, self.groups, self.deformable_groups,
                                                 self.im2col_step)


class DeformConv(_DeformBase):
    def forward(self, input, offset):
        assert 2 * self._taps() == offset.shape[1]
        return DeformConvFunction.apply(input, offset, self.weight, self.bias, self.stride, self.padding,
                                        self.dilation, self.groups, self.deformable_groups, self.im2col_step)


_DeformConv = DeformConvFunction.apply


class DeformConvPack(DeformConv):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True, lr_mult=0.1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         deformable_groups, im2col_step, bias)
        self.conv_offset = nn.Conv2d(self.in_channels, 2 * self._taps(), kernel_size=self.kernel_size,
                                     stride=self.stride, padding=self.padding, bias=True)
        self.conv_offset.lr_mult = lr_mult
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, input):
        offset = self.conv_offset(input)
        return DeformConvFunction.apply(input, offset, self.weight, self.bias, self.stride, self.padding,
                                        self.dilation, self.groups, self.deformable_groups, self.im2col_step)
