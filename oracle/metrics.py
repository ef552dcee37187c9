"""CPU restatement of the reference's depth metrics -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Follows lib/metrics/rdf_gan_metric.py: evaluate_all :59-151 (per image: RMSE, MAE, iRMSE, iMAE, REL, D^1..D^3, then the
mean over images) and evaluate_batch :17-57 (whole batch as one pixel set, six metrics); fp32 per-pixel arithmetic and
fp32 sums like the torch ops it replaces.  Pinned by tests/golden/metric_golden.json (outputs of the reference class).
"""
import numpy as np

f32 = np.float32


def image_metrics(gt, pred, t_valid=1e-4, evaluate_mask=None):
    """rdf_gan_metric.py:66-134 for one image -> 8 float32 values."""
    gt, pred = np.asarray(gt, f32), np.asarray(pred, f32)
    pred_inv = f32(1.0) / (pred + f32(1e-8))
    gt_inv = f32(1.0) / (gt + f32(1e-8))
    mask = gt > f32(t_valid)
    if evaluate_mask is not None:
        mask = mask & np.asarray(evaluate_mask, bool)
    nv = f32(mask.sum()) + f32(1e-8)
    p, g, pi, gi = pred[mask], gt[mask], pred_inv[mask].copy(), gt_inv[mask].copy()
    pi[p <= f32(t_valid)] = 0
    gi[g <= f32(t_valid)] = 0
    d = p - g
    da = np.abs(d)
    di = pi - gi
    ratio = np.maximum(g / (p + f32(1e-8)), p / (g + f32(1e-8)))
    s = lambda v: np.sum(v, dtype=f32)
    return np.array([np.sqrt(s(d * d) / nv), s(da) / nv, np.sqrt(s(di * di) / nv), s(np.abs(di)) / nv, s(da / (g + f32(1e-8))) / nv,
                     s((ratio < f32(1.25)).astype(f32)) / nv, s((ratio < f32(1.25 ** 2)).astype(f32)) / nv,
                     s((ratio < f32(1.25 ** 3)).astype(f32)) / nv], f32)


def evaluate_all(results, t_valid=1e-4):
    """rdf_gan_metric.py:59-151 -> dict name -> float32 (mean over images)."""
    names = ['RMSE', 'MAE', 'iRMSE', 'iMAE', 'REL', 'D^1', 'D^2', 'D^3']
    m = np.stack([image_metrics(r['gt'], r['pd'], t_valid, r.get('evaluate_mask')) for r in results]).mean(axis=0)
    return {n: m[i] for i, n in enumerate(names)}


def evaluate_batch(gt, pred, t_valid=1e-4):
    """rdf_gan_metric.py:17-57 -> (1, 6) [RMSE, MAE, REL, D^1, D^2, D^3]."""
    m = image_metrics(np.asarray(gt).reshape(-1), np.asarray(pred).reshape(-1), t_valid)
    return m[[0, 1, 4, 5, 6, 7]][None]
