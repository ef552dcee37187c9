"""Depth-completion metrics on the device -- drop-in for lib/metrics/rdf_gan_metric.py (RDFGANMetric :6-151).

Same class name, constructor, ``metric_name`` list and ``evaluate_batch`` / ``evaluate_all`` signatures and return
values.  The reference masks, gathers and reduces with ~40 ATen kernels per image on the CPU; here one pass of
``rdfc_depth_metric_sums`` produces the nine sums per image (the de-normalisation of Eval.inference,
lib/evaluator/evaluator.py:27-29, can be fused in through ``evaluate_device``).  Inputs given as numpy arrays or CPU
tensors are moved to the GPU; there is no CPU implementation.
"""
import numpy as np
import torch

from . import _cabi as C


def _sums(pred, gt, evaluate_mask, std, mean, t_valid):
    """pred, gt: (B, ...) fp32 CUDA tensors -> (B, 9) float64 sums."""
    C.require_cuda(pred, gt)
    B = pred.shape[0]
    pred = pred.reshape(B, -1).float().contiguous()
    gt = gt.reshape(B, -1).float().contiguous()
    if pred.shape != gt.shape:
        raise RuntimeError(f"pred {tuple(pred.shape)} and gt {tuple(gt.shape)} differ")
    n = pred.shape[1]
    em = None
    if evaluate_mask is not None:
        em = evaluate_mask.to(pred.device).reshape(B, -1).to(torch.uint8).contiguous()
    sums = torch.empty((B, 9), dtype=torch.float64, device=pred.device)
    partial = torch.empty((B * C.lib.rdfc_depth_metric_nchunk(n) * 9,), dtype=torch.float64, device=pred.device)
    with torch.cuda.device(pred.device):
        C.check(C.lib.rdfc_depth_metric_sums(C.ptr(pred), C.ptr(gt), C.ptr(em), float(std), float(mean), float(t_valid),
                                             C.ptr(sums), C.ptr(partial), B, n, C.stream_ptr(pred.device)))
    return sums


def _metrics_from_sums(s):
    """(…, 9) sums -> (…, 8) [RMSE, MAE, iRMSE, iMAE, REL, D^1, D^2, D^3] (rdf_gan_metric.py:100-131)."""
    nv = s[..., 0] + 1e-8
    return torch.stack([torch.sqrt(s[..., 1] / nv), s[..., 2] / nv, torch.sqrt(s[..., 3] / nv), s[..., 4] / nv, s[..., 5] / nv,
                        s[..., 6] / nv, s[..., 7] / nv, s[..., 8] / nv], dim=-1)


def _to_cuda(x):
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.asarray(x))
    return x if x.is_cuda else x.cuda()


class RDFGANMetric:
    def __init__(self, t_valid=1e-4):
        self.t_valid = t_valid
        self.metric_name = ['RMSE', 'MAE', 'iRMSE', 'iMAE', 'REL', 'D^1', 'D^2', 'D^3']

    def evaluate_device(self, gt, pred, std=1.0, mean=0.0, evaluate_mask=None):
        """Per-image metrics (B, 8) float64 on the device; gt / pred (B, ...) normalised tensors, de-normalised as
        x * std + mean inside the kernel (evaluator.py:27-29)."""
        return _metrics_from_sums(_sums(_to_cuda(pred), _to_cuda(gt), evaluate_mask, std, mean, self.t_valid))

    def evaluate_batch(self, gt, pred):
        """rdf_gan_metric.py:17-57: the whole batch as one set of pixels -> (1, 6) [RMSE, MAE, REL, D^1, D^2, D^3]."""
        gt, pred = _to_cuda(gt), _to_cuda(pred)
        s = _sums(pred.reshape(1, -1), gt.reshape(1, -1), None, 1.0, 0.0, self.t_valid)
        m = _metrics_from_sums(s)[:, [0, 1, 4, 5, 6, 7]]
        return m.to(torch.float32).detach()

    def evaluate_all(self, results, logger=None):
        """rdf_gan_metric.py:59-151: results = [{'gt': image, 'pd': image, ['evaluate_mask': mask]}, ...]; the mean over
        images of the per-image metrics, returned as a dict and logged / printed like the reference."""
        per_image = []
        groups = {}
        for i, r in enumerate(results):
            groups.setdefault(tuple(np.shape(r['gt'])), []).append(i)
        out = [None] * len(results)
        for shape, idxs in groups.items():          # same-shaped images go through the kernel as one batch
            for j0 in range(0, len(idxs), 64):
                sel = idxs[j0:j0 + 64]
                gt = torch.stack([_to_cuda(results[i]['gt']).float() for i in sel])
                pd = torch.stack([_to_cuda(results[i]['pd']).float() for i in sel])
                em = None
                if any('evaluate_mask' in results[i] for i in sel):
                    em = torch.stack([_to_cuda(results[i].get('evaluate_mask', torch.ones(shape, dtype=torch.bool))).bool()
                                      for i in sel])
                m = _metrics_from_sums(_sums(pd, gt, em, 1.0, 0.0, self.t_valid)).to(torch.float32).cpu().numpy()
                for k, i in enumerate(sel):
                    out[i] = m[k:k + 1]
        per_image = np.concatenate(out, axis=0)
        metrics = np.mean(per_image, axis=0, keepdims=True)
        ret = {name: metrics[0, idx] for idx, name in enumerate(self.metric_name)}
        if logger is not None:
            logger.log(' ')
            for k, v in ret.items():
                logger.log(f'{k}: {v}')
        else:
            for k, v in ret.items():
                print(f'{k}: {v}')
        return ret
