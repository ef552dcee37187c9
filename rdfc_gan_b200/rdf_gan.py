"""The GAN training step -- mirrors RDFGAN of C/lib/models/rdf_gan.py:18-207 (C = /root/reference/RDFC-GAN): ``set_input``,
``forward``, ``backward_D``, ``backward_G``, ``optimize_parameters`` with the same loss terms and the same update order
(D first, then G with D frozen), for ``gan_loss_type`` 'lsgan' / 'vanilla' / 'wgan'.

Differences by design (B200 data-parallel path): instead of wrapping the nets in DistributedDataParallel (:51-58)
every gradient lives in ONE pre-flattened buffer per net (parallel.GradientBucket) that is all-reduced over NCCL / NVLink once
per backward, and the loss scalars are reduced in one small all-reduce (parallel.reduce_losses) instead of one per key
(base.py:121-132).  The generator is called as RDFC-GAN's ``G(rgb, raw_depth, normal)``.
"""
import types

import torch

from .discriminator import PatchGANDiscriminator
from .init_weights import init_weights
from .losses import GANLoss, L1_loss
from .parallel import GradientBucket, reduce_losses

DEFAULT_ARGS = dict(gan_loss_type='lsgan', lambda_l1_rgb_branch=10.0, lambda_l1_depth_branch=10.0, lambda_l1_fusion=10.0,
                    optimizer='adam', lr=2e-4, beta1=0.5, beta2=0.999, wgan_clip_value=0.01)


class RDFGAN:
    def __init__(self, generator, discriminator=None, device='cuda', distributed=False, args=None, is_train=True):
        a = dict(DEFAULT_ARGS)
        a.update(args or {})
        self.args = types.SimpleNamespace(**a)
        self.device = torch.device(device)
        self.distributed = distributed
        self.is_train = is_train
        self.G = generator.to(self.device)
        self.D = (discriminator if discriminator is not None else PatchGANDiscriminator(in_channels=1)).to(self.device)
        self.models = dict(G=self.G, D=self.D)
        if is_train:
            self.init_optimizer()
            self.criterionGAN = GANLoss(self.args.gan_loss_type).to(self.device)
            self.criterionL1 = L1_loss
            # one flat gradient buffer per net: p.grad are views into it (no torch.cat / copy-back around the all-reduce)
            self.bucket_G = GradientBucket(self.G.parameters())
            self.bucket_D = GradientBucket(self.D.parameters())

    def init_weights(self):
        """rdf_gan.py:60-61"""
        init_weights(self.G)
        init_weights(self.D)

    def init_optimizer(self):
        kind = self.args.optimizer.lower()
        if kind == 'adam':
            mk = lambda params: torch.optim.Adam(params, lr=self.args.lr, betas=(self.args.beta1, self.args.beta2))
        elif kind == 'sgd':
            mk = lambda params: torch.optim.SGD(params, lr=self.args.lr)
        elif kind == 'rmsprop':
            mk = lambda params: torch.optim.RMSprop(params, lr=self.args.lr)
        else:
            raise NotImplementedError(f'Only Adam, SGD, RMSprop optimizers are supported, but got {kind}')
        self.optimizer_G = mk(self.G.parameters())
        self.optimizer_D = mk(self.D.parameters())
        self.optimizers = dict(G=self.optimizer_G, D=self.optimizer_D)

    def train(self):
        for m in self.models.values():
            m.train()

    def eval(self):
        for m in self.models.values():
            m.eval()

    @staticmethod
    def set_requires_grad(models, requires_grad=False):
        for m in (models if isinstance(models, list) else [models]):
            if m is not None:
                for p in m.parameters():
                    p.requires_grad = requires_grad

    # ------------------------------------------------------------------------------------------------------------
    def set_input(self, data):
        """rdf_gan.py:82-93 (+ the surface normals RDFC-GAN's generator reads)"""
        dev = self.device
        self.real_A, self.real_B = data['rgb'].to(dev), data['gt_depth'].to(dev)
        self.corrupted_B = data['raw_depth'].to(dev)
        self.normal = data['normal'].to(dev) if 'normal' in data else self.real_A
        self.mask = data['depth_mask'].to(dev) if 'depth_mask' in data else torch.ones_like(self.real_B)
        self.image_loss_weight = self.mask / (self.mask.sum() + 1e-6)

    def forward(self):
        ret = self.G(self.real_A, self.corrupted_B, self.normal)
        self.fake_B_rgb_branch, self.conf_map_rgb_branch = ret['depth_map_1'], ret['confidence_map_1']
        self.fake_B_depth_branch, self.conf_map_depth_branch = ret['depth_map_2'], ret['confidence_map_2']
        self.final_depth = ret['pred_depth']

    def backward_D(self):
        """rdf_gan.py:135-160"""
        pred_fake = self.D(self.fake_B_rgb_branch.detach())
        loss_D_fake = self.criterionGAN(pred_fake, False)
        pred_real = self.D(self.real_B)
        loss_D_real = self.criterionGAN(pred_real, True)
        loss_D = (loss_D_real + loss_D_fake) * 0.5
        loss_D.backward()
        ret = dict(loss_D=loss_D, loss_D_real=loss_D_real, loss_D_fake=loss_D_fake)
        if self.args.gan_loss_type == 'wgangp':
            raise NotImplementedError("wgangp needs double backward through the discriminator (rdf_gan.py:112-129)")
        return ret

    def backward_G(self):
        """rdf_gan.py:162-190"""
        a = self.args
        pred_fake = self.D(self.fake_B_rgb_branch)
        loss_G_GAN = self.criterionGAN(pred_fake, True)
        w = self.image_loss_weight
        loss_L1_rgb_branch = self.criterionL1(self.fake_B_rgb_branch, self.real_B, weight=w) * a.lambda_l1_rgb_branch
        loss_L1_depth_branch = self.criterionL1(self.fake_B_depth_branch, self.real_B, weight=w) * a.lambda_l1_depth_branch
        loss_L1_fusion = self.criterionL1(self.final_depth, self.real_B, weight=w) * a.lambda_l1_fusion
        loss_G = loss_G_GAN + loss_L1_rgb_branch + loss_L1_depth_branch + loss_L1_fusion
        loss_G.backward()
        return dict(loss_G_GAN=loss_G_GAN, loss_L1_rgb_branch=loss_L1_rgb_branch, loss_L1_depth_branch=loss_L1_depth_branch,
                    loss_L1_fusion=loss_L1_fusion)

    def optimize_parameters(self):
        """rdf_gan.py:192-207; the gradient all-reduce DDP hides inside backward() is the explicit bucket.allreduce()."""
        loss_stats = {}
        self.forward()
        self.set_requires_grad(self.D, True)
        self.bucket_D.zero()
        loss_stats.update(self.backward_D())
        self.bucket_D.allreduce()
        self.optimizer_D.step()
        if self.args.gan_loss_type == 'wgan':
            for p in self.D.parameters():
                p.data.clamp_(-self.args.wgan_clip_value, self.args.wgan_clip_value)
        self.set_requires_grad(self.D, False)
        self.bucket_G.zero()
        loss_stats.update(self.backward_G())
        self.bucket_G.allreduce()
        self.optimizer_G.step()
        return reduce_losses(loss_stats)
