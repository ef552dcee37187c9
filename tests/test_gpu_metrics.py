"""-m gpu: the device depth metrics (rdfc_depth_metric_sums through the RDFGANMetric drop-in) against the golden outputs of
the reference's own class and the numpy oracle."""
import io
import json
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from _synth import metric_inputs

pytestmark = pytest.mark.gpu


def test_metrics_vs_reference_golden(golden_dir):
    from rdfc_gan_b200.metrics import RDFGANMetric
    gold = json.load(open(f"{golden_dir}/metric_golden.json"))
    for seed, g in gold.items():
        n_img, H, W, with_mask = g["cfg"]
        res = metric_inputs(int(seed), n_img, H, W, with_mask)
        m = RDFGANMetric()
        assert m.metric_name == list(g["evaluate_all"])
        with redirect_stdout(io.StringIO()) as out:
            ret = m.evaluate_all([{k: torch.from_numpy(v) for k, v in r.items()} for r in res])
        assert "RMSE:" in out.getvalue()                     # printed like the reference when no logger is given
        for k, v in g["evaluate_all"].items():
            assert abs(float(ret[k]) - v) <= 1e-5 * max(1.0, abs(v)), (seed, k, float(ret[k]), v)
        b = m.evaluate_batch(np.stack([r["gt"] for r in res]), np.stack([r["pd"] for r in res]))
        assert tuple(b.shape) == (1, 6) and b.dtype == torch.float32
        assert np.allclose(b.cpu().numpy()[0], np.array(g["evaluate_batch"], np.float32), rtol=1e-5, atol=1e-6)


def test_metrics_denormalise_mask_and_edge_cases():
    from oracle import metrics as om
    from rdfc_gan_b200.metrics import RDFGANMetric
    m = RDFGANMetric(t_valid=1e-4)
    # de-normalisation fused in (evaluator.py:27-29), ragged size (not a multiple of the chunk), an all-invalid image
    res = metric_inputs(7, 3, 131, 127, True)
    std, mean = 2.5, 3.0
    gt = np.stack([(r["gt"] - mean) / std for r in res]).astype(np.float32)
    pd = np.stack([(r["pd"] - mean) / std for r in res]).astype(np.float32)
    gt[2] = (0.0 - mean) / std                                    # image 2: nothing valid after de-normalisation
    em = np.stack([r["evaluate_mask"] for r in res])
    got = m.evaluate_device(torch.from_numpy(gt).cuda(), torch.from_numpy(pd).cuda(), std, mean, torch.from_numpy(em).cuda())
    for i in range(3):
        g = (gt[i] * np.float32(std) + np.float32(mean)).astype(np.float32)
        p = (pd[i] * np.float32(std) + np.float32(mean)).astype(np.float32)
        want = om.image_metrics(g, p, 1e-4, em[i])
        assert np.allclose(got[i].cpu().numpy(), want, rtol=2e-5, atol=1e-6), (i, got[i], want)
    assert torch.all(got[2] == 0)                                 # 0 / (0 + 1e-8), like the reference
    # bench-size batch: sums are additive over a split of the batch (deterministic reduction => exactly)
    big = metric_inputs(9, 8, 228, 304, False)
    G = torch.from_numpy(np.stack([r["gt"] for r in big])).cuda()
    P = torch.from_numpy(np.stack([r["pd"] for r in big])).cuda()
    a = m.evaluate_device(G, P)
    b = torch.cat([m.evaluate_device(G[:3], P[:3]), m.evaluate_device(G[3:], P[3:])])
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        m.evaluate_device(G, P[:, :100])
