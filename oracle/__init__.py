"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's generator hot path (RDFC-GAN / RDF-GAN generator forward, NLSPN,
DCNv2 boundary).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``rdfc_gan_b200/`` does (tests/test_layout.py
greps for it).  The product path has no CPU fallback.

Pieces
------
* ``dcn_oracle.c``   plain C (double accumulation) restatement of the DCN v1 / v2 forward + backward and of the
                     NLSPN propagation loop; built by ``oracle/Makefile`` / ``oracle.build()`` into
                     ``oracle/_build/libdcn_oracle.so``.
* ``dcn.py``         ctypes wrappers over that library (numpy / CPU torch tensors in and out).
* ``nlspn.py``       numpy restatement of ``nlspn/nlspn_model.py`` on top of ``dcn.py``.
* ``generator.py``   torch-CPU *functional* restatement of ``RDFGenerator.forward`` / ``DCVGANGenerator.forward``
                     over a plain ``state_dict`` (the reference itself is Python/PyTorch, so its conv / norm
                     arithmetic is restated with CPU ``torch.nn.functional`` calls; the DCN calls go to ``dcn.py``).

Parity pinning
--------------
The reference ships no golden vectors for this path; its only executable checks are the
known-answer properties in ``deformconv/test.py``.  The oracle is therefore pinned two ways:
(1) ``tests/golden/*.npz`` -- outputs of the reference's own Python (``/root/reference``), imported in the build
container by ``tests/golden/make_golden.py`` (its CUDA-only DCN extension replaced by
``torchvision.ops.deform_conv2d``, as BASELINE.json prescribes) -- ``tests/test_oracle_golden.py`` checks every
oracle function against them; (2) the ``deformconv/test.py`` properties, ported in ``tests/test_dcn_properties.py``.
``oracle/_ref`` does not exist: the reference has no CPU-buildable native source (``deformconv/src/cpu/*.cpp``
are ``AT_ERROR`` stubs) and its Python cannot travel to the GPU box.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libdcn_oracle.so")


def build(force: bool = False) -> str:
    """Compile dcn_oracle.c with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "dcn_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-fopenmp", "-std=c11", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH
