"""Build libmsda_b200.so in-tree with nvcc for sm_100a (replaces the reference's
models/ops/setup.py + make.sh torch CUDAExtension build).  No torch headers are involved:
the library is plain CUDA behind a C ABI (include/msda_b200.h)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmsda_b200.so")
SOURCES = ["msda_capi.cu", "msda_percall.cu", "msda_snippet.cu", "msda_deterministic.cu", "msda_mask.cu", "msda_frames.cu", "msda_tail.cu", "msda_planar.cu"]
HEADERS = ["msda_common.cuh", "msda_fast.cuh", "msda_snippet_common.cuh", "msda_internal.h", os.path.join("..", "..", "include", "msda_b200.h")]
OBJ_DIR = os.path.join(PKG, "lib", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libmsda_b200.so")
    return exe


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile every .cu of the package into snipper_b200/lib/libmsda_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + os.environ.get("MSDA_NVCC_EXTRA", "").split() + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:  # the translation units compile side by side
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, res in results:
        if verbose:
            print(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, res.stdout + res.stderr))
    link = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + [obj for _, obj, _ in results]
    res = subprocess.run(link, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
