"""Stage the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box
(where /root/reference does not exist) and can be executed there against the product:

  * as the subject of the drop-in tests (tests/test_reference_dropin_gpu.py): the reference's own
    ``build_model`` / ``deformable_transformer.py`` / ``MSDeformAttn`` / ``MSDeformAttnFunction`` with the
    extension shim (surface 1) or the fused module (surface 3) plugged in;
  * as the GPU baseline of bench.py (``gpu_baseline``): the reference's per-(t1,t2) loop calling its own
    vendored CUDA op (compiled in place by oracle/build_ref.py).

Only importable Python packages are staged, byte for byte; nothing under baseline/_ref/ is tracked by git
(.gitignore) and nothing in the product imports it.  The reference is not a pip package (no setup.py /
pyproject at its root; its only buildable artefact is the CUDA op), so ``pip install --target baseline/_ref``
does not apply -- this copy is the "install".
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["models", "util", "datasets", "main.py", "engine.py", "eval_utils.py", "dataset_class.py"]
SKIP_DIRS = {"src", "build", "dist", "__pycache__", "MultiScaleDeformableAttention.egg-info", "poseval_old"}


def stage(src="/root/reference", force=False):
    """Copy the reference's Python packages to baseline/_ref/ (returns the path, or None if ``src`` is absent)."""
    if not os.path.isdir(os.path.join(src, "models", "ops")):
        return DEST if os.path.isdir(os.path.join(DEST, "models", "ops")) else None
    if force and os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST, exist_ok=True)

    def ignore(d, names):
        return [n for n in names if n in SKIP_DIRS or n.endswith((".pyc", ".so", ".o", ".gif", ".jpg", ".png"))]

    for item in ITEMS:
        s, d = os.path.join(src, item), os.path.join(DEST, item)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=ignore, dirs_exist_ok=True)
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    return DEST


if __name__ == "__main__":
    print(stage(*(sys.argv[1:2] or ["/root/reference"]), force=True))
