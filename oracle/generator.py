"""torch-CPU functional restatement of the reference generator forward (TEST INFRASTRUCTURE ONLY).

Restates, over a plain ``state_dict`` (reference key names), the arithmetic of
  G/rdf_generator.py:280-414            RDFGenerator.forward           (G = /root/reference/RDFC-GAN/lib/models/generator/rdf_generator)
  F/.../rdf_gan_generator.py:233-361    DCVGANGenerator.forward        (same body, stems read global_guidance_module(rgb))
  G/encoder_decoder/common.py:29-61     conv_bn_relu / convt_bn_relu
  G/encoder_decoder/encoder_decoder.py  EncoderDecoder (torchvision BasicBlock layers as en2..en5)
  G/model_utils.py:7-129                EqualLR, AdaptiveInstanceNorm (W-AdaIN), AdaIN, IN
  G/nlspn/nlspn_model.py                via oracle/nlspn.py
BatchNorm is evaluated in eval mode (running statistics) unless ``train_bn=True``.
It is pinned against tests/golden/generator_*.npz (outputs of the reference's own Python).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import nlspn as onlspn


def _bn(sd, p, x, train_bn):
    if train_bn:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def _conv_bn_relu(sd, p, x, stride=1, padding=1, bn=True, relu=True, train_bn=False):
    """common.py:29-43"""
    x = F.conv2d(x, sd[p + ".0.weight"], sd.get(p + ".0.bias"), stride, padding)
    if bn:
        x = _bn(sd, p + ".1", x, train_bn)
    if relu:
        x = F.leaky_relu(x, 0.2)
    return x


def _convt_bn_relu(sd, p, x, train_bn=False):
    """common.py:46-61 with the EncoderDecoder arguments k3 s2 p1 op1 (encoder_decoder.py:53-59)"""
    x = F.conv_transpose2d(x, sd[p + ".0.weight"], sd.get(p + ".0.bias"), stride=2, padding=1, output_padding=1)
    x = _bn(sd, p + ".1", x, train_bn)
    return F.leaky_relu(x, 0.2)


def _basic_block(sd, p, x, stride, train_bn):
    """torchvision.models.resnet.BasicBlock (call sites common.py:11-26)"""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    out = F.relu(_bn(sd, p + ".bn1", out, train_bn))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn(sd, p + ".bn2", out, train_bn)
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), train_bn)
    return F.relu(out + x)


def _res_layer(sd, p, x, stride, train_bn):
    i = 0
    while (p + f".{i}.conv1.weight") in sd:
        x = _basic_block(sd, p + f".{i}", x, stride if i == 0 else 1, train_bn)
        i += 1
    return x


def _crop_cat(fd, fe):
    """rdf_generator.py:244-260"""
    fd = fd[:, :, :fe.shape[2], :fe.shape[3]]
    return torch.cat((fd, fe), dim=1)


def _wadain(sd, p, x, style, weighting):
    """model_utils.py:53-90 (+ EqualLR :11-15)"""
    w_orig = sd[p + ".style.linear.weight_orig"]
    w = w_orig * math.sqrt(2.0 / w_orig.shape[1])
    st = F.linear(style.permute(0, 2, 3, 1), w, sd[p + ".style.linear.bias"]).permute(0, 3, 1, 2)
    gamma, beta = st.chunk(2, 1)
    out = F.instance_norm(x, eps=1e-5)
    if not weighting:
        return gamma * out + beta
    gw = F.conv2d(x, sd[p + ".gamma_weight_layer.weight"], sd[p + ".gamma_weight_layer.bias"])
    bw = F.conv2d(x, sd[p + ".beta_weight_layer.weight"], sd[p + ".beta_weight_layer.bias"])
    return gw * gamma * out + bw * beta


def _adain(x, style, eps=1e-5):
    """model_utils.py:92-116 (unbiased variance)"""
    N, C = x.shape[:2]
    s_var = style.reshape(N, C, -1).var(dim=2) + eps
    c_var = x.reshape(N, C, -1).var(dim=2) + eps
    s_mean = style.reshape(N, C, -1).mean(dim=2)
    c_mean = x.reshape(N, C, -1).mean(dim=2)
    normalized = (x - c_mean.view(N, C, 1, 1)) / c_var.sqrt().view(N, C, 1, 1)
    return normalized * s_var.sqrt().view(N, C, 1, 1) + s_mean.view(N, C, 1, 1)


def _in_fuse(sd, p, x, style):
    """model_utils.py:119-129"""
    out = F.instance_norm(torch.cat([x, style], dim=1), eps=1e-5)
    return F.conv2d(out, sd[p + ".down_channel.weight"], sd[p + ".down_channel.bias"])


def generator_forward(sd, stem_in, depth, *, fuse="WAdaIN", adain_weighting=False, use_nlspn_refine=True,
                      nlspn_configs=None, train_bn=False, return_intermediates=False):
    """``stem_in`` is what the reference feeds both stems: ``normal`` for RDFGenerator (rdf_generator.py:286,289),
    ``global_guidance_module(rgb)`` for DCVGANGenerator.  Returns the dict of rdf_generator.py:409-413."""
    sd = {k: (v.detach().float().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}
    stem_in, depth = stem_in.float().cpu(), depth.float().cpu()
    tb = train_bn
    inter = {}

    def fuse_layer(i, x, style):
        p = f"fuse_layer{i}"
        if fuse == "WAdaIN":
            return _wadain(sd, p, x, style, adain_weighting)
        if fuse == "AdaIN":
            return _adain(x, style)
        if fuse == "IN":
            return _in_fuse(sd, p, x, style)
        raise NotImplementedError(fuse)

    with torch.no_grad():
        rgb_fe1 = _conv_bn_relu(sd, "rgb_branch_en1", stem_in, bn=False)                       # :286
        depth_fe1 = torch.cat([_conv_bn_relu(sd, "depth_branch_en1_rgb", stem_in, bn=False),   # :289-292
                               _conv_bn_relu(sd, "depth_branch_en1_depth", depth, bn=False)], dim=1)
        R, D = "rgb_branch_encoder_decoder", "depth_branch_encoder_decoder"
        rgb_fe, depth_fe = {1: rgb_fe1}, {1: depth_fe1}
        for i, stride in ((2, 1), (3, 2), (4, 2), (5, 2)):                                     # :295-308
            rgb_fe[i] = _res_layer(sd, f"{R}.en{i}", rgb_fe[i - 1], stride, tb)
            depth_fe[i] = _res_layer(sd, f"{D}.en{i}", depth_fe[i - 1], stride, tb)
        rgb_fe[6] = _conv_bn_relu(sd, f"{R}.en6", rgb_fe[5], stride=2, train_bn=tb)             # :311-312
        depth_fe[6] = _conv_bn_relu(sd, f"{D}.en6", depth_fe[5], stride=2, train_bn=tb)

        rgb_fd, depth_fd = rgb_fe[6], depth_fe[6]
        for n, lvl in enumerate((5, 4, 3, 2)):                                                 # :315-368
            fz = fuse_layer(n + 1, rgb_fd, depth_fd)
            inter[f"fuse{n + 1}"] = fz
            rgb_fd = _crop_cat(_convt_bn_relu(sd, f"{R}.de{lvl}", fz, tb), rgb_fe[lvl])
            depth_fd = _crop_cat(_convt_bn_relu(sd, f"{D}.de{lvl}", depth_fd, tb), depth_fe[lvl])
        inter["rgb_fd2"], inter["depth_fd2"] = rgb_fd, depth_fd

        depth_map_1 = torch.tanh(_conv_bn_relu(                                                # :374-376
            sd, "rgb_pred_dec0", _crop_cat(_conv_bn_relu(sd, "rgb_pred_dec1", rgb_fd, train_bn=tb), rgb_fe1),
            bn=False, relu=False))
        c1 = _crop_cat(_conv_bn_relu(sd, "rgb_conf_dec1", rgb_fd, train_bn=tb), rgb_fe1)       # :378-379
        confidence_map_1 = torch.sigmoid(F.conv2d(c1, sd["rgb_conf_dec0.0.weight"], sd["rgb_conf_dec0.0.bias"], 1, 1))

        pred_init = torch.tanh(_conv_bn_relu(                                                  # :385-387
            sd, "id_dec0", _crop_cat(_conv_bn_relu(sd, "id_dec1", depth_fd, train_bn=tb), depth_fe1),
            bn=False, relu=False))
        guide = None
        if use_nlspn_refine:                                                                   # :390-392
            guide = _conv_bn_relu(sd, "gd_dec0",
                                  _crop_cat(_conv_bn_relu(sd, "gd_dec1", depth_fd, train_bn=tb), depth_fe1),
                                  bn=False, relu=False)
        c2 = _crop_cat(_conv_bn_relu(sd, "cf_dec1", depth_fd, train_bn=tb), depth_fe1)         # :397-398
        confidence = torch.sigmoid(F.conv2d(c2, sd["cf_dec0.0.weight"], sd["cf_dec0.0.bias"], 1, 1))
        inter.update(pred_init=pred_init, guide=guide, confidence=confidence)

        if use_nlspn_refine:                                                                   # :400
            cfg = dict(prop_kernel=3, prop_time=18, affinity="TGASS", affinity_gamma=0.5, conf_prop=True,
                       preserve_input=False)
            cfg.update(nlspn_configs or {})
            P = "nlspn_refine_module.prop_layer"
            y, offset, aff = onlspn.nlspn_forward(
                pred_init.numpy(), guide.numpy(), confidence.numpy(), depth.numpy(),
                sd[P + ".conv_offset_aff.weight"].numpy(), sd[P + ".conv_offset_aff.bias"].numpy(),
                sd[P + ".aff_scale_const"].numpy(), k_f=cfg["prop_kernel"], prop_time=cfg["prop_time"],
                affinity=cfg["affinity"], conf_prop=cfg["conf_prop"], preserve_input=cfg["preserve_input"])
            depth_map_2 = torch.from_numpy(np.ascontiguousarray(y))
            inter.update(offset=torch.from_numpy(offset), aff=torch.from_numpy(aff))
        else:
            depth_map_2 = pred_init
        depth_map_2 = torch.clamp(depth_map_2, min=-1, max=1)                                  # :401
        score = F.softmax(torch.cat([confidence_map_1, confidence], dim=1), 1)                 # :403-404
        final = torch.sum(torch.cat([depth_map_1, depth_map_2], dim=1) * score, dim=1, keepdim=True)  # :405-406
    ret = dict(depth_map_1=depth_map_1, confidence_map_1=confidence_map_1, depth_map_2=depth_map_2,
               confidence_map_2=confidence, pred_depth=final)
    if return_intermediates:
        ret["_inter"] = inter
    return ret
