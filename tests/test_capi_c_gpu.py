"""The drop-in boundary is a plain C ABI: a C11 program (tests/capi/capi_smoke.c, no torch, no C++) links
libmsda_b200.so, runs forward + backward on a seeded problem and checks them against the C oracle."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "capi", "capi_smoke.c")


def _build(tmpdir):
    from oracle import c_oracle
    from snipper_b200.build import LIB_DIR, LIB_PATH
    assert os.path.exists(LIB_PATH), "libmsda_b200.so is not built"
    oracle_so = c_oracle.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(str(tmpdir), "capi_smoke")
    cmd = ["gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"), SRC,
           "-o", exe, "-L", LIB_DIR, "-lmsda_b200", "-L", os.path.dirname(oracle_so), "-lmsda_oracle",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + LIB_DIR, "-Wl,-rpath," + os.path.dirname(oracle_so), "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_program_compiles_and_links_against_the_c_abi(tmp_path):
    """CPU part: the header is valid C11 and every symbol the C caller uses resolves at link time."""
    _build(tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_program_matches_the_oracle_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK" in res.stdout
