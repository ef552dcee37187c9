"""Losses of the training step -- same call signatures and arithmetic as C/lib/losses/gan_loss.py
(L1_loss :6-22, mse_loss :102-118, GANLoss :169-206; C = /root/reference/RDFC-GAN).  Scalar reductions over (B,1,H,W) maps:
plain differentiable PyTorch on CUDA tensors (they are outside the dense path, < 0.1 % of a step)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _weighted_sum(loss, pred, weight):
    if weight is not None:
        weight = weight.float()
    else:                                   # uniform weights that sum to one (gan_loss.py:13-15)
        weight = torch.ones_like(pred)
        weight = weight / (weight.sum() + 1e-6)
    if weight.dim() != loss.dim():
        weight = weight.unsqueeze(1)
    return (weight * loss).sum()


def L1_loss(pred, target, weight=None, reduction='sum'):
    assert reduction == 'sum'
    return _weighted_sum(F.l1_loss(pred, target, reduction='none'), pred, weight)


def mse_loss(pred, target, weight=None, reduction='sum'):
    assert reduction == 'sum'
    return _weighted_sum(F.mse_loss(pred, target, reduction='none'), pred, weight)


def binary_cross_entropy_loss(pred, target, weight=None, reduction='sum'):
    assert reduction == 'sum'
    return _weighted_sum(F.binary_cross_entropy_with_logits(pred, target.float(), reduction='none'), pred, weight)


class GANLoss(nn.Module):
    """gan_loss.py:169-206: 'lsgan' (weighted MSE against a constant target), 'vanilla' (BCE with logits), 'wgan' / 'wgangp'
    (+- mean of the critic)."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer('real_label', torch.tensor([target_real_label]))
        self.register_buffer('fake_label', torch.tensor([target_fake_label]))
        self.gan_mode = gan_mode
        if gan_mode == 'lsgan':
            self.criterion = mse_loss
        elif gan_mode == 'vanilla':
            self.criterion = binary_cross_entropy_loss
        elif gan_mode in ('wgangp', 'wgan'):
            self.criterion = None
        else:
            raise NotImplementedError('gan mode %s not implemented' % gan_mode)

    def get_target_tensor(self, prediction, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(prediction)

    def __call__(self, prediction, target_is_real, weight=None):
        if self.gan_mode in ('lsgan', 'vanilla'):
            return self.criterion(prediction, self.get_target_tensor(prediction, target_is_real), weight)
        return -prediction.mean() if target_is_real else prediction.mean()
