"""Import the reference's own Python generator in the BUILD container (needs /root/reference; never used on the
GPU box, never imported by tests at run time).  Recipe = SURVEY.md section 8c / Appendix F:

* ``DCN`` (the reference's CUDA-only pybind extension, imported by name at nlspn/modulated_deform_conv_func.py:13)
  is replaced by a stub whose forward is ``torchvision.ops.deform_conv2d`` -- the CPU stand-in BASELINE.json
  prescribes; backward flows through torchvision's autograd.
* package shells bypass ``lib/models/generator/__init__.py`` (it imports a ``build_generator`` file that the
  reference's own .gitignore swallowed).
"""
import os
import sys
import types

import torch
from torchvision.ops import deform_conv2d

REF = os.environ.get("RDFC_REFERENCE", "/root/reference")
C = os.path.join(REF, "RDFC-GAN")
F_ = os.path.join(REF, "RDF-GAN")


def _dcn_stub():
    m = types.ModuleType("DCN")

    def mdcf(i, w, b, off, msk, kh, kw, sh, sw, ph, pw, dh, dw, g, dg, step):
        return deform_conv2d(i, off.contiguous(), w, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw),
                             mask=msk.contiguous())

    def dcf(i, w, b, off, kh, kw, sh, sw, ph, pw, dh, dw, g, dg, step):
        return deform_conv2d(i, off.contiguous(), w, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))

    m.modulated_deform_conv_forward = mdcf
    m.deform_conv_forward = dcf
    return m


def _shell(name, path):
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    sys.modules[name] = pkg


def import_rdfc():
    """-> (RDFGenerator, NLSPNRefineModule, init_weights) of RDFC-GAN."""
    assert os.path.isdir(C), "reference checkout not present"
    sys.modules["DCN"] = _dcn_stub()
    for name, sub in (("lib", "lib"), ("lib.models", "lib/models"), ("lib.models.generator", "lib/models/generator")):
        _shell(name, os.path.join(C, sub))
    if C not in sys.path:
        sys.path.insert(0, C)
    from lib.models.generator.rdf_generator.rdf_generator import RDFGenerator
    from lib.models.generator.rdf_generator.nlspn.nlspn_model import NLSPNRefineModule
    from lib.models.init_weights import init_weights
    return RDFGenerator, NLSPNRefineModule, init_weights


class DeformFn(torch.autograd.Function):
    """Not used: the reference's own ModulatedDeformConvFunction (once_differentiable) wraps DCN.*; gradients for
    the golden files are taken through torchvision's autograd directly (see make_golden.py)."""
