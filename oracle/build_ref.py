"""TEST INFRASTRUCTURE ONLY -- compile the reference's vendored CUDA op for sm_100a.

Sources are compiled where they lie under /root/reference/models/ops/src (never copied into the
repo); the only addition is a force-included compat header (oracle/ref_build/compat_shim.h).
Output: oracle/_ref/MultiScaleDeformableAttention.so (git-ignored, travels to the GPU box).
It serves as (a) a second parity oracle on the GPU ("matches the vendored CUDA op") and (b) the
head-to-head GPU baseline: the reference kernels recompiled for B200.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SNIPPER_REFERENCE", "/root/reference") + "/models/ops/src"
OUT = os.path.join(HERE, "_ref")
NAME = "MultiScaleDeformableAttention"


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(HERE, "ref_build", "compat_shim.h")
    srcs = [os.path.join(REF_SRC, "vision.cpp"),
            os.path.join(REF_SRC, "cpu", "ms_deform_attn_cpu.cpp"),
            os.path.join(REF_SRC, "cuda", "ms_deform_attn_cuda.cu")]
    load(name=NAME, sources=srcs, extra_include_paths=[REF_SRC],
         extra_cflags=["-DWITH_CUDA", "-include", shim, "-w"],
         extra_cuda_cflags=["-gencode=arch=compute_100a,code=sm_100a", "-include", shim, "-w",
                            "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                            "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=OUT, with_cuda=True, is_python_module=False, verbose=verbose)
    return os.path.join(OUT, NAME + ".so")


def load_ref():
    """Import the compiled reference extension (None if it was never built)."""
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
