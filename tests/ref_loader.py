"""Test utility: import the REAL, unmodified reference with the three shims of SURVEY.md Appendix A.

Search order: $SNIPPER_REFERENCE, /root/reference (the build container), baseline/_ref (the copy
`baseline/stage_reference.py` stages at build() time so the GPU box has one too).  Nothing of it is modified."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find():
    for cand in (os.environ.get("SNIPPER_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models", "ops", "modules")):
            return cand
    return None


REF = _find()


def available():
    return REF is not None


def import_reference():
    """Put the reference on sys.path with the torchvision-version and pretrained-download shims applied."""
    import torchvision
    torchvision.__version__ = "0.9.0"          # util/misc.py:20-22 parses "0.26" as < 0.5
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.backbone as bb
    bb.is_main_process = lambda: False          # no pretrained download (models/backbone.py:105-107)


def reference_args(**overrides):
    import_reference()
    import main as refmain
    args = refmain.get_args_parser().parse_args([])
    args.device = "cpu"
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def build_reference_model(**overrides):
    args = reference_args(**overrides)
    from models.model import build_model
    model, criterion, post = build_model(args)
    return model, args
