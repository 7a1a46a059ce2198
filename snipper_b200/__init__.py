"""snipper_b200 -- B200-native (sm_100a) multi-scale deformable attention for Snipper.

Scope: ONE hot path of JimmyZou/Snipper -- ``MSDeformAttn`` forward + backward -- behind the
reference's own three surfaces:

  1. extension module  ``MultiScaleDeformableAttention``   -> snipper_b200/shim/
  2. autograd Function ``MSDeformAttnFunction``            -> snipper_b200.functions
  3. nn.Module         ``MSDeformAttn``                    -> snipper_b200.modules

All compute runs in hand-written CUDA kernels inside ``lib/libmsda_b200.so`` (C ABI in
``include/msda_b200.h``).  There is no CPU / PyTorch fallback: a missing library raises.
"""
import os
import sys

from . import capi, ops  # noqa: F401
from .functions import MSDeformAttnFunction, ms_deform_attn
from .modules import MSDeformAttn
from .ops import is_deterministic, set_deterministic
from .runtime import GraphRunner
from .layers import disable_fused_layer_tails, enable_fused_layer_tails

__all__ = ["MSDeformAttn", "MSDeformAttnFunction", "ms_deform_attn", "set_deterministic",
           "is_deterministic", "install_extension_shim", "install_module", "GraphRunner",
           "enable_fused_layer_tails", "disable_fused_layer_tails"]

_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def install_extension_shim():
    """Make ``import MultiScaleDeformableAttention`` resolve to the B200 kernels (surface 1).

    Also rebinds ``MSDA`` inside an already-imported reference
    ``models.ops.functions.ms_deform_attn_func`` (the reference looks the name up at call time,
    ms_deform_attn_func.py:28,38)."""
    if _SHIM_DIR not in sys.path:
        sys.path.insert(0, _SHIM_DIR)
    import MultiScaleDeformableAttention as shim
    ref_func = sys.modules.get("models.ops.functions.ms_deform_attn_func")
    if ref_func is not None:
        ref_func.MSDA = shim
    return shim


def install_module():
    """Swap the fused ``MSDeformAttn`` into an imported reference tree (surface 3): rebinds
    ``models.ops.modules.MSDeformAttn`` and ``models.deformable_transformer.MSDeformAttn`` so
    ``build_model`` constructs this class."""
    for name in ("models.ops.modules", "models.ops.modules.ms_deform_attn", "models.deformable_transformer"):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.MSDeformAttn = MSDeformAttn
