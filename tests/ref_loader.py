"""Test utility: import the REAL reference (only in the build container, where /root/reference
exists) with the three shims of SURVEY.md Appendix A.  Nothing is copied or modified."""
import os
import sys

REF = os.environ.get("SNIPPER_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models", "ops"))


def build_reference_model(**overrides):
    import torchvision
    torchvision.__version__ = "0.9.0"          # util/misc.py:20-22 parses "0.26" as < 0.5
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.backbone as bb
    bb.is_main_process = lambda: False          # no pretrained download (models/backbone.py:105-107)
    import main as refmain
    args = refmain.get_args_parser().parse_args([])
    args.device = "cpu"
    for k, v in overrides.items():
        setattr(args, k, v)
    from models.model import build_model
    model, criterion, post = build_model(args)
    return model, args
