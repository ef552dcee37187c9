"""Generates tests/golden/metric_golden.json from the reference's OWN RDFGANMetric (lib/metrics/rdf_gan_metric.py), imported
from /root/reference in the build container.  Run once:  python tests/golden/make_metric_golden.py
The inputs are re-created from the seeds by tests/_synth.metric_inputs, only the outputs are stored."""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _synth import metric_inputs  # noqa: E402

REF = os.path.join(os.environ.get("RDFC_REFERENCE", "/root/reference"), "RDFC-GAN", "lib", "metrics")


def ref_class():
    pkg = types.ModuleType("refmetrics")
    pkg.__path__ = [REF]
    sys.modules["refmetrics"] = pkg
    for name in ("base", "rdf_gan_metric"):
        spec = importlib.util.spec_from_file_location(f"refmetrics.{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refmetrics.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refmetrics.rdf_gan_metric"].RDFGANMetric


def main():
    Metric = ref_class()
    out = {}
    for seed, (n_img, H, W, with_mask) in {1: (5, 57, 76, False), 2: (3, 228, 304, True), 3: (4, 31, 45, False)}.items():
        res = metric_inputs(seed, n_img, H, W, with_mask)
        ref_results = [{k: torch.from_numpy(v) for k, v in r.items()} for r in res]
        with redirect_stdout(io.StringIO()):
            ret = Metric().evaluate_all(ref_results)
        gt = np.stack([r['gt'] for r in res])
        pd = np.stack([r['pd'] for r in res])
        batch = Metric().evaluate_batch(torch.from_numpy(gt), torch.from_numpy(pd)).numpy()
        out[str(seed)] = {"cfg": [n_img, H, W, with_mask], "evaluate_all": {k: float(v) for k, v in ret.items()},
                          "evaluate_batch": [float(v) for v in batch[0]]}
    json.dump(out, open(os.path.join(HERE, "metric_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:600])


if __name__ == "__main__":
    main()
