"""``init_weights`` -- the reference's definition of "random init" (C/lib/models/init_weights.py:5-33, called from
C/lib/models/rdfc_gan.py:120-123): every module whose class name contains ``Conv`` or ``Linear`` and that has a
``weight`` gets it redrawn (normal | xavier | kaiming | orthogonal) and its bias zeroed; ``BatchNorm2d`` gets
weight ~ N(1, gain), bias 0.  Modules are visited in ``nn.Module.apply`` order, so with the same RNG seed the drop-in
generators of this package end up with the same tensors as the reference's (tests/test_host_logic.py pins a checksum
generated from the reference).  Two consequences the reference relies on and this keeps: the zero-initialised
``conv_offset_aff`` of NLSPN is overwritten, and ``EqualLinear``'s inner linear (which only has ``weight_orig``) is
skipped."""
import torch.nn as nn
from torch.nn import init

_DRAW = {
    'normal': lambda w, gain: init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: init.xavier_normal_(w, gain=gain),
    'kaiming': lambda w, gain: init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: init.orthogonal_(w, gain=gain),
}


def init_weights(net: nn.Module, init_type: str = 'normal', init_gain: float = 0.02) -> nn.Module:
    if init_type not in _DRAW:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
    draw = _DRAW[init_type]

    def visit(m):
        name = type(m).__name__
        if hasattr(m, 'weight') and ('Conv' in name or 'Linear' in name):
            draw(m.weight.data, init_gain)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in name:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print('initialize network with %s' % init_type)
    return net.apply(visit)
