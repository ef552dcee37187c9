"""Loads the UNMODIFIED reference Python (installed under the git-ignored baseline/_ref/, which travels to the GPU box) without
touching this repo's product package.  Used by ``bench.py --impl reference`` / ``cpu_baseline`` (CPU: the reference's ``DCN``
extension has no CPU kernel, so ``torchvision.ops.deform_conv2d`` serves it, as BASELINE.json prescribes) and by
baseline/time_ref_gpu.py (GPU: the reference's own CUDA extension built by baseline/build_ref_gpu.py).

``install()`` runs in the BUILD container only (it reads /root/reference); nothing under baseline/_ref is tracked by git.
"""
import importlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFD = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("RDFC_REFERENCE", "/root/reference")
_IGNORE = shutil.ignore_patterns("src", "*.cu", "*.cuh", "*.cpp", "*.h", "__pycache__", "*.sh", "*.so", "build")


def install():
    """Copy the reference's Python files for the hot path (and its training step) into baseline/_ref/.  Returns True if the
    reference checkout was present."""
    c, f = os.path.join(REFERENCE, "RDFC-GAN"), os.path.join(REFERENCE, "RDF-GAN")
    if not os.path.isdir(c):
        return False
    os.makedirs(REFD, exist_ok=True)
    jobs = [(os.path.join(c, "lib/models/generator/rdf_generator"), "rdf_generator"),
            (os.path.join(f, "lib/models"), "rdf_gan_lib/lib/models"),      # imported as `lib.models...`, like F/ does
            (os.path.join(c, "lib/models/module"), "rdfc_lib/lib/models/module"),             # the training step's pieces of RDFC-GAN
            (os.path.join(c, "lib/models/discriminator"), "rdfc_lib/lib/models/discriminator"),
            (os.path.join(c, "lib/losses"), "rdfc_lib/lib/losses")]
    for src, dst in jobs:
        if os.path.isdir(src):
            out = os.path.join(REFD, dst)
            shutil.rmtree(out, ignore_errors=True)
            shutil.copytree(src, out, ignore=_IGNORE)
    for d in ("rdf_gan_generator", "rdf_gan_segmentator", "rdf_gan_backbone", "rdf_gan_module"):      # older layout
        shutil.rmtree(os.path.join(REFD, d), ignore_errors=True)
    open(os.path.join(REFD, "rdf_gan_lib", "lib", "__init__.py"), "w").close()
    singles = [(os.path.join(c, "lib/models/discriminator/patch_gan_discriminator.py"), "ref_patch_gan_discriminator.py"),
               (os.path.join(c, "lib/losses/gan_loss.py"), "ref_gan_loss.py"),
               (os.path.join(c, "lib/models/generator/resnet_generator.py"), "ref_resnet_generator.py"),
               (os.path.join(c, "lib/models/init_weights.py"), "ref_init_weights.py")]
    for src, dst in singles:
        if os.path.isfile(src):
            shutil.copy(src, os.path.join(REFD, dst))
    return True


def available():
    return os.path.isfile(os.path.join(REFD, "rdf_generator", "rdf_generator.py"))


def _dcn_cpu_stub():
    from torchvision.ops import deform_conv2d
    m = types.ModuleType("DCN")

    def mdcf(i, w, b, off, msk, kh, kw, sh, sw, ph, pw, dh, dw, g, dg, step):
        return deform_conv2d(i, off.contiguous(), w, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw),
                             mask=msk.contiguous())

    def dcf(i, w, b, off, kh, kw, sh, sw, ph, pw, dh, dw, g, dg, step):
        return deform_conv2d(i, off.contiguous(), w, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))

    m.modulated_deform_conv_forward = mdcf
    m.deform_conv_forward = dcf
    return m


def _dcn_gpu_extension():
    so = os.path.join(REFD, "build", "DCN.so")
    spec = importlib.util.spec_from_file_location("DCN", so)
    dcn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dcn)
    return dcn


def _stub_esanet(pkg):
    # rdf_generator.py imports ESANet (never used by RDFGenerator.forward); its own imports need the rest of the reference
    # repo, so a stub module stands in for it -- the generator files themselves stay unmodified
    for name in (pkg + ".segmentator", pkg + ".segmentator.esa_net"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    stub = types.ModuleType(pkg + ".segmentator.esa_net.esa_net_one_modality")
    stub.ESANetOneModality = type("ESANetOneModality", (), {})
    sys.modules[stub.__name__] = stub


def load_rdfc(gpu=False):
    """-> the reference's RDFGenerator class (RDFC-GAN).  gpu=False: DCN served by torchvision (CPU); gpu=True: the reference's
    own CUDA extension (baseline/_ref/build/DCN.so)."""
    if not available():
        raise FileNotFoundError("baseline/_ref/rdf_generator is missing: run `python baseline/ref_loader.py` in the build container")
    sys.modules["DCN"] = _dcn_gpu_extension() if gpu else _dcn_cpu_stub()
    if REFD not in sys.path:
        sys.path.insert(0, REFD)
    _stub_esanet("rdf_generator")
    return importlib.import_module("rdf_generator.rdf_generator").RDFGenerator


def load_rdf_gan(gpu=False):
    """-> (DCVGANGenerator, ESANetOneModality) of RDF-GAN, imported as ``lib.models...`` from baseline/_ref/rdf_gan_lib.  The
    generator's ``nlspn`` sub-package is missing from the reference checkout; the RDFC-GAN copy of the
    same package is aliased in, as SURVEY 8c prescribes."""
    load_rdfc(gpu)
    root = os.path.join(REFD, "rdf_gan_lib")
    if root not in sys.path:
        sys.path.insert(0, root)
    importlib.import_module("lib.models.generator")                      # F/'s __init__ files are empty / harmless
    base = "lib.models.generator.rdf_gan_generator"
    pkg = types.ModuleType(base)                                         # bypass its __init__ (imports before the alias exists)
    pkg.__path__ = [os.path.join(root, "lib", "models", "generator", "rdf_gan_generator")]
    sys.modules[base] = pkg
    sys.modules[base + ".nlspn"] = importlib.import_module("rdf_generator.nlspn")
    sys.modules[base + ".nlspn.nlspn_model"] = importlib.import_module("rdf_generator.nlspn.nlspn_model")
    gen = importlib.import_module(base + ".rdf_gan_generator").DCVGANGenerator
    esa = importlib.import_module("lib.models.segmentator.esa_net.esa_net_one_modality").ESANetOneModality
    return gen, esa


def load_rdfc_training():
    """-> (PatchGANDiscriminator, GANLoss, L1_loss) of RDFC-GAN (C/lib/models/discriminator/patch_gan_discriminator.py,
    C/lib/losses/gan_loss.py), imported as ``lib...`` through package shells (the discriminator package's __init__ imports a
    build_discriminator file the checkout lacks).  Call in a process that has not imported RDF-GAN's ``lib``."""
    root = os.path.join(REFD, "rdfc_lib", "lib")
    for name, sub in (("lib", ""), ("lib.models", "models"), ("lib.models.module", "models/module"),
                      ("lib.models.discriminator", "models/discriminator"), ("lib.losses", "losses")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(root, sub)]
        sys.modules[name] = pkg
    D = importlib.import_module("lib.models.discriminator.patch_gan_discriminator").PatchGANDiscriminator
    gl = importlib.import_module("lib.losses.gan_loss")
    return D, gl.GANLoss, gl.L1_loss


def differentiable_dcn_on_cpu():
    """The reference's ModulatedDeformConvFunction calls the extension's backward, which has no CPU kernel: for a TRAINING step
    on the CPU, ``apply`` is served by torchvision.ops.deform_conv2d at that boundary (its autograd is the reference's col2im
    arithmetic).  Call after load_rdfc()."""
    from torchvision.ops import deform_conv2d
    nm = importlib.import_module("rdf_generator.nlspn.nlspn_model")

    class Shim:
        @staticmethod
        def apply(inp, offset, mask, weight, bias, stride, padding, dilation, groups, dg, step):
            return deform_conv2d(inp, offset.contiguous(), weight, bias, stride=stride, padding=padding, dilation=dilation,
                                 mask=mask.contiguous())
    nm.ModulatedDeformConvFunction = Shim


def load_single(fname, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFD, fname))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


if __name__ == "__main__":
    print("installed" if install() else "reference checkout not present", "->", REFD)
