"""Parameter containers of the two-branch encoder/decoder -- state_dict-compatible with
encoder_decoder/common.py:29-61 (conv_bn_relu / convt_bn_relu: ``<name>.0`` conv, ``<name>.1`` norm) and
encoder_decoder/encoder_decoder.py:5-61 (en2..en5 = torchvision ResNet-18/34 ``layer1..4`` BasicBlocks, en6, de5..de2).

These modules only HOLD parameters (and define the layer geometry); the arithmetic runs in rdfc_gan_b200.engine on the
sm_100a kernels.  Calling ``forward`` on them is an error on purpose: there is no PyTorch fallback path.
"""
import torch.nn as nn

RESNET_BLOCKS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}


def _no_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container; the forward pass runs in rdfc_gan_b200.engine")


def conv_bn_relu(channels_in, channels_out, kernel, stride=1, padding=0, bn=True, _in=False, relu=True):
    """common.py:29-43 (same Sequential indices, so the same state_dict keys)."""
    assert not (bn and _in)
    layers = [nn.Conv2d(channels_in, channels_out, kernel, stride, padding, bias=not bn)]
    if bn:
        layers.append(nn.BatchNorm2d(channels_out))
    if _in:
        layers.append(nn.InstanceNorm2d(channels_out))
    if relu:
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*layers)


def convt_bn_relu(ch_in, ch_out, kernel, stride=1, padding=0, output_padding=0, bn=True, relu=True):
    """common.py:46-61"""
    assert (kernel % 2) == 1, 'only odd kernel is supported but kernel = {}'.format(kernel)
    layers = [nn.ConvTranspose2d(ch_in, ch_out, kernel, stride, padding, output_padding, bias=not bn)]
    if bn:
        layers.append(nn.BatchNorm2d(ch_out))
    if relu:
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*layers)


class BasicBlock(nn.Module):
    """Same attribute names as torchvision.models.resnet.BasicBlock (conv1, bn1, conv2, bn2, downsample)."""
    forward = _no_forward

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


def _res_layer(inplanes, planes, blocks, stride):
    return nn.Sequential(*([BasicBlock(inplanes, planes, stride)] + [BasicBlock(planes, planes) for _ in range(1, blocks)]))


class EncoderDecoder(nn.Module):
    """encoder_decoder.py:5-61"""
    forward = _no_forward

    def __init__(self, encoder_type='resnet34', skip_type='concat', encoder_channels=[64, 128, 256, 512, 512],
                 decoder_channels=[256, 128, 64, 64], pretrained_on_imagenet=False):
        super().__init__()
        if encoder_type not in RESNET_BLOCKS:
            raise NotImplementedError
        if pretrained_on_imagenet:
            # common.py:5-17 loads pretrained_model/resnet/*.pth, which no checkout ships
            raise RuntimeError("pretrained_on_imagenet=True needs the reference's pretrained_model/resnet/*.pth; "
                               "load a checkpoint with load_state_dict instead")
        nb = RESNET_BLOCKS[encoder_type]
        cat = skip_type == 'concat'
        dec = [(encoder_channels[-1], decoder_channels[0]),
               (decoder_channels[0] + encoder_channels[-2] if cat else decoder_channels[0], decoder_channels[1]),
               (decoder_channels[1] + encoder_channels[-3] if cat else decoder_channels[1], decoder_channels[2]),
               (decoder_channels[2] + encoder_channels[-4] if cat else decoder_channels[2], decoder_channels[3])]
        self.en2 = _res_layer(64, 64, nb[0], 1)        # torchvision layer1, run at full resolution
        self.en3 = _res_layer(64, 128, nb[1], 2)
        self.en4 = _res_layer(128, 256, nb[2], 2)
        self.en5 = _res_layer(256, 512, nb[3], 2)
        self.en6 = conv_bn_relu(encoder_channels[-2], encoder_channels[-1], kernel=3, stride=2, padding=1)
        self.de5 = convt_bn_relu(dec[0][0], dec[0][1], kernel=3, stride=2, padding=1, output_padding=1)
        self.de4 = convt_bn_relu(dec[1][0], dec[1][1], kernel=3, stride=2, padding=1, output_padding=1)
        self.de3 = convt_bn_relu(dec[2][0], dec[2][1], kernel=3, stride=2, padding=1, output_padding=1)
        self.de2 = convt_bn_relu(dec[3][0], dec[3][1], kernel=3, stride=2, padding=1, output_padding=1)
