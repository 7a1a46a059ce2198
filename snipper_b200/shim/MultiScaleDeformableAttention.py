"""Drop-in for the reference's compiled extension module ``MultiScaleDeformableAttention``
(pybind11 module of models/ops/src/vision.cpp:13-16, built by models/ops/setup.py).

Put this directory on ``sys.path`` (or call ``snipper_b200.install_extension_shim()``) BEFORE the
reference's ``models.ops.functions.ms_deform_attn_func`` is imported; its guarded
``import MultiScaleDeformableAttention as MSDA`` (:18-21) then binds to these functions and the
unmodified reference runs the B200 kernels with ``--use_pytorch_deform 0``.

Same two functions, same argument order, same error behaviour as the reference host code
(models/ops/src/ms_deform_attn.h:20-61, models/ops/src/cuda/ms_deform_attn_cuda.cu:20-153):
every tensor must be a contiguous CUDA tensor; a CPU ``value`` raises "Not implemented on the CPU".
"""
import torch

import snipper_b200.ops  # noqa: F401  (registers torch.ops.snipper_b200.*)


def _strict(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output=None):
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    named = [(value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
             (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")]
    if grad_output is not None:
        named.append((grad_output, "grad_output"))
    for t, name in named:
        if not t.is_contiguous():
            raise RuntimeError("%s tensor has to be contiguous" % name)
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    _strict(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    return torch.ops.snipper_b200.msda_forward(value, spatial_shapes, level_start_index, sampling_loc,
                                               attn_weight, im2col_step)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, im2col_step):
    _strict(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output)
    gv, gl, ga = torch.ops.snipper_b200.msda_backward(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step,
        snipper_b200.ops.is_deterministic())
    return [gv, gl, ga]
