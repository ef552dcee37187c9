"""Builds the reference's OWN CUDA deformable-conv extension (module ``DCN``) for sm_100a into baseline/_ref/ (git-ignored;
it travels to the GPU box with gpurun), so that the reference's unmodified ``nlspn_model.py`` can be timed on the same B200
as this repo's kernels (SURVEY 8d "also time").  Runs in the BUILD container only (reads /root/reference).

Sources are copied to baseline/_ref/dcn_src and two tokens are patched for torch >= 2 (``.type()`` -> ``.scalar_type()``
in AT_DISPATCH); deform_psroi_pooling_cuda.cu is left out (it needs THC headers that no longer exist) and
vision.cpp's psroi bindings are stubbed.  Nothing from baseline/_ref is tracked by git."""
import os
import re
import shutil
import sys

REF = "/root/reference/RDFC-GAN/lib/models/generator/rdf_generator/nlspn"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    src = os.path.join(OUT, "dcn_src")
    shutil.rmtree(src, ignore_errors=True)
    shutil.copytree(os.path.join(REF, "deformconv", "src"), src)
    for f in ("nlspn_model.py", "modulated_deform_conv_func.py"):
        shutil.copy(os.path.join(REF, f), os.path.join(OUT, f))
    # the whole (unmodified) rdf_generator package, python files only, for the full-generator timing
    gen = os.path.join(OUT, "rdf_generator")
    shutil.rmtree(gen, ignore_errors=True)
    shutil.copytree(os.path.dirname(REF), gen, ignore=shutil.ignore_patterns("src", "*.cu", "*.cuh", "*.cpp", "*.h", "__pycache__", "*.sh"))
    os.remove(os.path.join(src, "cuda", "deform_psroi_pooling_cuda.cu"))
    for f in ("modulated_deform_conv_cuda.cu", "deform_conv_cuda.cu"):
        p = os.path.join(src, "cuda", f)
        s = open(p).read()
        s = re.sub(r"(AT_DISPATCH_FLOATING_TYPES\(\s*\w+)\.type\(\)", r"\1.scalar_type()", s)
        open(p, "w").write(s)
    # vision.cpp binds the psroi functions: keep the module buildable without that file
    p = os.path.join(src, "vision.cpp")
    s = open(p).read()
    s = re.sub(r'.*deform_psroi_pooling.*\n', '', s)
    open(p, "w").write(s)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    build = os.path.join(OUT, "build")
    os.makedirs(build, exist_ok=True)
    sources = [os.path.join(src, "vision.cpp"), os.path.join(src, "cuda", "modulated_deform_conv_cuda.cu"),
               os.path.join(src, "cuda", "deform_conv_cuda.cu")]
    sources += [os.path.join(src, "cpu", f) for f in os.listdir(os.path.join(src, "cpu")) if f.endswith(".cpp") and "psroi" not in f]
    load(name="DCN", sources=sources, extra_include_paths=[src], extra_cflags=["-DWITH_CUDA"],
         extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                            "-D__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=build, verbose=False, is_python_module=False)
    print("built", [f for f in os.listdir(build) if f.endswith(".so")])


if __name__ == "__main__":
    sys.exit(main())
