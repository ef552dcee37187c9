"""PatchGAN discriminator -- state_dict-compatible with C/lib/models/discriminator/patch_gan_discriminator.py:6-40 and its
ConvModule (C/lib/models/module/conv_norm_act.py: ``model.<i>.conv``, ``model.<i>.bn2d``).

Five 4x4 convolutions (stride 2, 2, 2, 1, 1; BatchNorm on the middle three; ReLU by default) over a one-channel depth map:
6.5 GFLOP per 228x304 image against the generator's 221, evaluated three times per step.  The 4x4 kernels are outside what
``rdfc_conv_forward`` covers (3x3 / 1x1), so this module runs on PyTorch's convolutions (cuDNN) in bf16 autocast-free fp32;
DESIGN.md section 7 lists it under "not on the repo's kernels yet"."""
import torch.nn as nn

_ACT = {'ReLU': lambda: nn.ReLU(inplace=True), 'LeakyReLU': lambda: nn.LeakyReLU(0.2, inplace=True)}


class ConvModule(nn.Module):
    """conv -> norm -> act with the reference's attribute names (conv_norm_act.py:9-120)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, norm=False, activation='ReLU'):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=not norm)
        self.with_norm = norm
        if norm:
            self.bn2d = nn.BatchNorm2d(out_channels, eps=1e-5)
        self.activation = activation
        if activation:
            self.act = _ACT[activation]()
        nn.init.kaiming_normal_(self.conv.weight, mode='fan_out', nonlinearity='leaky_relu' if activation == 'LeakyReLU' else 'relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn2d(x)
        if self.activation:
            x = self.act(x)
        return x


class PatchGANDiscriminator(nn.Module):
    def __init__(self, in_channels, out_channels=(64, 128, 256, 512, 1), kernel_size=(4, 4, 4, 4, 4), stride=(2, 2, 2, 1, 1),
                 padding=(1, 1, 1, 1, 1), conv_cfg=dict(type='Conv2d'), norm_cfg=dict(type='BN2d'), activation='ReLU'):
        super().__init__()
        assert out_channels[-1] == 1, f"The channel of the feature map obtained by the last conv layer must be 1, but got {out_channels[-1]}"
        if (conv_cfg or {}).get('type', 'Conv2d').lower() != 'conv2d' or (norm_cfg or {}).get('type', 'BN2d') != 'BN2d':
            raise NotImplementedError("PatchGANDiscriminator: Conv2d + BN2d only (the reference's default and only used configuration)")
        chans = [in_channels] + list(out_channels)
        n = len(chans) - 1
        self.model = nn.Sequential(*[
            ConvModule(chans[i], chans[i + 1], kernel_size[i], stride[i], padding[i], norm=(norm_cfg is not None and 0 < i < n - 1),
                       activation=None if i == n - 1 else activation) for i in range(n)])

    def forward(self, x):
        return self.model(x)
