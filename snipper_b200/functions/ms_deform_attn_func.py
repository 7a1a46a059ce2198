"""Autograd boundary of the per-call op -- same name, argument order and gradient contract as the
reference's ``MSDeformAttnFunction`` (models/ops/functions/ms_deform_attn_func.py:24-42):
gradients flow to ``value``, ``sampling_locations`` and ``attention_weights`` only.

The reference file also carries a pure-PyTorch grid_sample implementation
(``ms_deform_attn_core_pytorch``, :45-65).  It is deliberately NOT provided here: this package
has no CPU / PyTorch fallback; that formulation lives in ``oracle/`` as test infrastructure.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import ops


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = torch.ops.snipper_b200.msda_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations,
            attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = torch.ops.snipper_b200.msda_backward(
            value, shapes, lsi, loc, attn, grad_output.contiguous(), ctx.im2col_step,
            ops.is_deterministic())
        return grad_value, None, None, grad_loc, grad_attn, None


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                   im2col_step=64):
    """Functional form; differentiable through the custom op's registered autograd formula."""
    return torch.ops.snipper_b200.msda_forward(value, spatial_shapes, level_start_index,
                                               sampling_locations, attention_weights, im2col_step)
