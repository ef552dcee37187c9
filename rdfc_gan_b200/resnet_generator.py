"""``ResnetGenerator`` (RDFC-GAN's G_B2A, depth -> RGB of the cycle) -- drop-in for
C/lib/models/generator/resnet_generator.py:6-98 (C = /root/reference/RDFC-GAN): same constructor, the same ``self.model``
Sequential of standard modules (state_dict keys ``model.<i>...`` load into / from the reference class with strict=True).

``forward`` walks that Sequential: the 3x3 convolutions (stride-2 down-sampling, the 2 x n_blocks residual-block convs) and the
two ConvTranspose2d(k3, s2, p1, op1) layers -- all with >= 64 input channels, > 97 % of the FLOPs -- run on the tcgen05 conv
kernel through ``train_ops.conv2d_nhwc`` (forward, data gradient, filter gradient), BatchNorm2d in train() mode on the fused
batch-statistics kernels; activations stay bf16 NHWC in between.  The reflection pads, the two 7x7 convolutions (1 or 3
channels on one side), LeakyReLU(0.01) / PReLU / Tanh and eval-mode / instance normalisation are differentiable PyTorch ops on
the same tensors.  CUDA only: there is no CPU path.
"""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi as C
from .train_ops import bn_act, conv2d_nhwc


class ResnetBlock(nn.Module):
    """resnet_generator.py:61-98: x + conv_block(x), conv_block = [pad, conv3x3, norm, LeakyReLU(0.01), (dropout), pad, conv3x3, norm]"""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        blocks = []
        for half in range(2):
            p = 0
            if padding_type == 'reflect':
                blocks.append(nn.ReflectionPad2d(1))
            elif padding_type == 'replicate':
                blocks.append(nn.ReplicationPad2d(1))
            elif padding_type == 'zero':
                p = 1
            else:
                raise NotImplementedError('padding [%s] is not implemented' % padding_type)
            blocks += [nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim)]
            if half == 0:
                blocks.append(nn.LeakyReLU(negative_slope=0.01, inplace=True))
                if use_dropout:
                    blocks.append(nn.Dropout(0.5))
        self.conv_block = nn.Sequential(*blocks)

    def forward(self, x):
        raise RuntimeError("ResnetBlock is evaluated by ResnetGenerator.forward (bf16 NHWC on the sm_100a kernels)")


class ResnetGenerator(nn.Module):
    def __init__(self, input_channels, output_channels, ngf=64, norm_layer='BN2d', use_dropout=False, n_blocks=6, padding_type='reflect'):
        super().__init__()
        assert n_blocks >= 0
        norm_layer = nn.BatchNorm2d if norm_layer.lower() == 'bn2d' else nn.InstanceNorm2d
        use_bias = norm_layer == nn.InstanceNorm2d
        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_channels, ngf, kernel_size=7, padding=0, bias=use_bias), norm_layer(ngf),
                 nn.LeakyReLU(negative_slope=0.01, inplace=True)]
        for i in range(2):
            mult = 2 ** i
            model += [nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1, bias=use_bias), norm_layer(ngf * mult * 2),
                      nn.PReLU(init=0.25)]
        for _ in range(n_blocks):
            model.append(ResnetBlock(ngf * 4, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout, use_bias=use_bias))
        for i in range(2):
            mult = 2 ** (2 - i)
            model += [nn.ConvTranspose2d(ngf * mult, ngf * mult // 2, kernel_size=3, stride=2, padding=1, output_padding=1, bias=use_bias),
                      norm_layer(ngf * mult // 2), nn.PReLU(init=0.25)]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_channels, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*model)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _tc_ok(m, x):
        """a layer the tensor-core conv covers: 3x3, >= 32-aligned input channels, stride 1 / 2"""
        Cin = x.shape[3]
        if isinstance(m, nn.ConvTranspose2d):
            return (m.kernel_size == (3, 3) and m.stride == (2, 2) and m.padding == (1, 1) and m.output_padding == (1, 1) and
                    Cin % 32 == 0 and m.out_channels % 8 == 0 and m.groups == 1)
        return (m.kernel_size == (3, 3) and m.stride in ((1, 1), (2, 2)) and m.padding in ((0, 0), (1, 1)) and m.dilation == (1, 1) and
                m.groups == 1 and Cin % 32 == 0 and m.out_channels % 8 == 0)

    def _walk(self, seq, x):
        """x: bf16 NHWC.  Torch modules see it as a channels-last NCHW view (no copy)."""
        def via_torch(mod, t):
            return mod(t.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        for m in seq:
            if isinstance(m, ResnetBlock):
                x = x + self._walk(m.conv_block, x)
            elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and self._tc_ok(m, x):
                if isinstance(m, nn.ConvTranspose2d):
                    y = conv2d_nhwc(x, m.weight, 3, 2, transposed=True)
                elif m.padding == (0, 0):        # the input was padded explicitly (reflection): zero-pad conv, then drop the border
                    if m.stride != (1, 1):
                        x = via_torch(m, x.float()).to(torch.bfloat16)
                        continue
                    y = conv2d_nhwc(x, m.weight, 3, 1)[:, 1:-1, 1:-1]
                else:
                    y = conv2d_nhwc(x, m.weight, 3, m.stride[0])
                x = y if m.bias is None else (y.float() + m.bias).to(torch.bfloat16)
            elif isinstance(m, nn.BatchNorm2d) and self.training and x.shape[3] % 8 == 0:
                x = bn_act(x.contiguous(), m, None, C.ACT_NONE)
            elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d, nn.InstanceNorm2d)):
                x = via_torch(m, x.float()).to(torch.bfloat16)           # 7x7 / thin convs, eval-mode or instance normalisation: fp32 PyTorch
            elif isinstance(m, nn.LeakyReLU):                            # not in place: the producer saved its output for backward
                x = F.leaky_relu(x, m.negative_slope)
            elif isinstance(m, nn.PReLU):                                # fp32 slope parameter: evaluate in fp32
                x = via_torch(m, x.float()).to(torch.bfloat16)
            else:                                                        # pads, PReLU / Tanh, dropout: element-wise PyTorch
                x = via_torch(m, x)
        return x

    def forward(self, x):
        C.require_cuda(x)
        y = self._walk(self.model, x.permute(0, 2, 3, 1).to(torch.bfloat16))
        return y.permute(0, 3, 1, 2).float().contiguous()
