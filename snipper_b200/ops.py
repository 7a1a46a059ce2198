"""torch custom ops over the C ABI (libmsda_b200.so).

PyTorch is plumbing here: it owns device memory and the current stream; all arithmetic happens
in the hand-written kernels.  Ops (namespace ``snipper_b200``):

  msda_forward / msda_backward          per-call op, the reference's extension functions
                                        (models/ops/src/vision.cpp:13-16)
  snippet_forward / snippet_backward    fused per-layer Snipper attention
                                        (models/ops/modules/ms_deform_attn.py:126-225)

Both forwards carry ``register_autograd`` formulas, so they compose with autograd / DDP as plain
nodes (no host sync, no unused parameters).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import capi

_DTYPES = {torch.float32: capi.MSDA_DTYPE_F32, torch.float64: capi.MSDA_DTYPE_F64,
           torch.bfloat16: capi.MSDA_DTYPE_BF16}

# process-wide switch for the deterministic (atomics-free) grad_value path
_deterministic = False


def set_deterministic(flag: bool) -> None:
    """Select the bit-reproducible two-pass backward (north_star: 'deterministic two-pass mode')."""
    global _deterministic
    _deterministic = bool(flag)


def is_deterministic() -> bool:
    return _deterministic


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class LaunchStats:
    """Counts kernel launches issued through the C ABI and, when ``timing`` is on, brackets each
    launch with CUDA events on the launching stream (bench.py's roofline measurement)."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.events = []  # (tag, dims, start_event, end_event)

    def reset(self):
        self.launches = 0
        self.events = []

    def kernel_ms(self):
        """tag -> list of per-launch durations in ms (call after a device synchronize)."""
        out = {}
        for tag, dims, s, e in self.events:
            out.setdefault((tag, dims), []).append(s.elapsed_time(e))
        return out


STATS = LaunchStats()


class _Launch:
    """with _Launch(tag, dims, device, n_kernels): <C call>"""

    def __init__(self, tag, dims, device, n_kernels=1):
        self.tag, self.dims, self.device, self.n = tag, dims, device, n_kernels

    def __enter__(self):
        if STATS.timing:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        STATS.launches += self.n
        if STATS.timing:
            self.e.record(torch.cuda.current_stream(self.device))
            STATS.events.append((self.tag, self.dims, self.s, self.e))
        return False


def _require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        # reference models/ops/src/ms_deform_attn.h:38,60
        raise RuntimeError("Not implemented on the CPU" if name == "value" else "%s must be a CUDA tensor" % name)


def _require_contiguous(t: Tensor, name: str) -> None:
    if not t.is_contiguous():
        # reference models/ops/src/cuda/ms_deform_attn_cuda.cu:28-32
        raise RuntimeError("%s tensor has to be contiguous" % name)


def _value_batch_stride(value: Tensor) -> int:
    """value (N,S,M,D): the inner three dims must be dense; the batch stride is free."""
    N, S, M, D = value.shape
    if N * S * M * D == 0:
        return 0
    if value.stride(3) != 1 or value.stride(2) != D or (S > 1 and value.stride(1) != M * D):
        raise RuntimeError("value tensor has to be contiguous")
    return value.stride(0) if N > 1 else S * M * D


def _check_percall(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    for t, name in ((value, "value"), (spatial_shapes, "spatial_shapes"),
                    (level_start_index, "level_start_index"), (sampling_loc, "sampling_loc"),
                    (attn_weight, "attn_weight")):
        _require_cuda(t, name)
    for t, name in ((spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                    (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
        _require_contiguous(t, name)
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError("expected value (N,S,M,D), sampling_loc (N,Lq,M,L,P,2), attn_weight (N,Lq,M,L,P)")
    if value.dtype not in _DTYPES:
        raise RuntimeError("ms_deform_attn: unsupported dtype %s (float32 / float64 / bfloat16)" % value.dtype)
    if value.dtype == torch.bfloat16:
        # bf16 mode: value / output / grad_output are bf16; locations and weights are computed in fp32
        if sampling_loc.dtype != torch.float32 or attn_weight.dtype != torch.float32:
            raise RuntimeError("bfloat16 value needs float32 sampling_loc and attn_weight")
        if value.shape[3] % 16 != 0 or value.shape[3] > 128:
            raise RuntimeError("bfloat16 ms_deform_attn needs head channels D % 16 == 0 and D <= 128")
    elif sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("value, sampling_loc and attn_weight must share one dtype")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    N, S, M, D = value.shape
    Nl, Lq, Ml, L, P, two = sampling_loc.shape
    if (Nl, Ml, two) != (N, M, 2) or tuple(attn_weight.shape) != (N, Lq, M, L, P):
        raise RuntimeError("sampling_loc / attn_weight shapes do not match value")
    if tuple(spatial_shapes.shape) != (L, 2) or level_start_index.numel() != L:
        raise RuntimeError("spatial_shapes must be (L,2) and level_start_index (L,)")
    return N, S, M, D, L, Lq, P


@torch.library.custom_op("snipper_b200::msda_forward", mutates_args=())
def msda_forward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor,
                 sampling_loc: Tensor, attn_weight: Tensor, im2col_step: int) -> Tensor:
    N, S, M, D, L, Lq, P = _check_percall(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    vbs = _value_batch_stride(value)
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _Launch("msda_forward", (N, S, M, D, L, Lq, P), value.device):
        st = capi.lib().msda_forward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(),
            N, S, M, D, L, Lq, P, vbs, int(im2col_step), _DTYPES[value.dtype], _stream(value.device))
    capi.check(st, "ms_deform_attn_forward", N, im2col_step)
    return out


@msda_forward.register_fake
def _(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    N, S, M, D = value.shape
    return value.new_empty((N, sampling_loc.shape[1], M * D))


@torch.library.custom_op("snipper_b200::msda_backward", mutates_args=())
def msda_backward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor,
                  sampling_loc: Tensor, attn_weight: Tensor, grad_output: Tensor,
                  im2col_step: int, deterministic: bool) -> Tuple[Tensor, Tensor, Tensor]:
    N, S, M, D, L, Lq, P = _check_percall(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    _require_cuda(grad_output, "grad_output")
    _require_contiguous(grad_output, "grad_output")
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output must be (N,Lq,M*D) in value's dtype")
    vbs = _value_batch_stride(value)
    bf16 = value.dtype == torch.bfloat16
    if bf16 and deterministic:
        raise RuntimeError("the deterministic backward is float32 only")
    # bf16 mode accumulates grad_value in fp32 (include/msda_b200.h) and rounds once at the end
    grad_value = torch.empty((N, S, M, D), dtype=torch.float32 if bf16 else value.dtype, device=value.device)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    flags = capi.MSDA_FLAG_DETERMINISTIC if deterministic else 0
    dt = _DTYPES[value.dtype]
    ws, ws_bytes = None, 0
    with torch.cuda.device(value.device):
        if deterministic:
            ws_bytes = capi.lib().msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, dt, flags)
            ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=value.device)
        ws_ptr = 0 if ws is None else (ws.data_ptr() + 255) // 256 * 256
        STATS.launches += 1
        st = capi.lib().msda_backward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
            N, S, M, D, L, Lq, P, vbs, int(im2col_step), dt, flags, ws_ptr, ws_bytes,
            _stream(value.device))
    capi.check(st, "ms_deform_attn_backward", N, im2col_step)
    if bf16:
        grad_value = grad_value.to(torch.bfloat16)
    return grad_value, grad_loc, grad_attn


@msda_backward.register_fake
def _(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step, deterministic):
    return (value.new_empty(value.shape), torch.empty_like(sampling_loc), torch.empty_like(attn_weight))


def _msda_setup_context(ctx, inputs, output):
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step = inputs
    ctx.im2col_step = im2col_step
    ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def _msda_backward_formula(ctx, grad_output):
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight = ctx.saved_tensors
    gv, gl, ga = torch.ops.snipper_b200.msda_backward(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
        grad_output.contiguous(), ctx.im2col_step, _deterministic)
    return gv, None, None, gl, ga, None


msda_forward.register_autograd(_msda_backward_formula, setup_context=_msda_setup_context)


# ------------------------------------------------------------------------------------------
# in-place masked zero-fill (value production)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("snipper_b200::masked_zero_", mutates_args=("data",))
def masked_zero_(data: Tensor, mask: Tensor) -> None:
    """data[i] = 0 where mask[i]; both contiguous, same shape; data float32 / bfloat16, mask bool."""
    _require_cuda(data, "data")
    _require_cuda(mask, "mask")
    if data.shape != mask.shape or not data.is_contiguous() or not mask.is_contiguous():
        raise RuntimeError("masked_zero_: data and mask must be contiguous tensors of one shape")
    if mask.dtype != torch.bool or data.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("masked_zero_: data must be float32 / bfloat16 and mask bool")
    with torch.cuda.device(data.device), _Launch("masked_zero", (data.numel(),), data.device):
        st = capi.lib().msda_masked_zero(data.data_ptr(), mask.data_ptr(), data.numel(), _DTYPES[data.dtype],
                                         _stream(data.device))
    capi.check(st, "msda_masked_zero")


def masked_zero_supported(data: Tensor, mask: Tensor) -> bool:
    return (data.is_cuda and mask.is_cuda and data.dtype in (torch.float32, torch.bfloat16) and
            mask.dtype == torch.bool and data.shape == mask.shape and data.is_contiguous() and
            mask.is_contiguous() and mask.data_ptr() % 16 == 0)


class MaskedValue(torch.autograd.Function):
    """``value.masked_fill(mask, 0)`` done IN PLACE on the freshly produced projection output.

    The backward is the identity ON PURPOSE: the only consumer of the result must be
    ``snippet_forward(..., value_mask=mask)``, whose backward zeroes the masked elements of the
    grad_value buffer it has just produced (same kernel, in place, no extra pass over the tensor).
    Used only by the fused path of :class:`snipper_b200.modules.MSDeformAttn`."""

    @staticmethod
    def forward(ctx, value, mask):
        torch.ops.snipper_b200.masked_zero_(value, mask)
        ctx.mark_dirty(value)
        return value

    @staticmethod
    def backward(ctx, grad):
        return grad, None


# ------------------------------------------------------------------------------------------
# fused snippet op
# ------------------------------------------------------------------------------------------
def snippet_supported(n_heads: int, d_head: int, n_levels: int, n_points: int, dtype) -> bool:
    """Shapes the fused kernels cover (include/msda_b200.h, msda_snippet_forward)."""
    return (dtype in (torch.float32, torch.bfloat16) and d_head % 16 == 0 and d_head <= 128 and
            n_levels * n_points <= 32 and n_levels <= 64)


def _check_snippet(value, spatial_shapes, level_start_index, offsets, logits, ref, n_frame):
    for t, name in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                    (offsets, "offsets"), (logits, "logits"), (ref, "reference_points")):
        _require_cuda(t, name)
    if value.dim() != 5 or offsets.dim() != 7 or logits.dim() != 6 or ref.dim() != 5:
        raise RuntimeError("expected value (N,T2,S,M,D), offsets (N,T1,Lq,M,L,P,2), logits (N,T1,Lq,M,L,P), "
                           "reference_points (N,T1,Lq,L,2)")
    N, T2, S, M, D = value.shape
    No, T1, Lq, Mo, L, P, two = offsets.shape
    if (No, Mo, two) != (N, M, 2) or tuple(logits.shape) != (N, T1, Lq, M, L, P):
        raise RuntimeError("offsets / logits shapes do not match value")
    if tuple(ref.shape) != (N, T1, Lq, L, 2) or tuple(spatial_shapes.shape) != (L, 2):
        raise RuntimeError("reference_points must be (N,T1,Lq,L,2) and spatial_shapes (L,2)")
    if not snippet_supported(M, D, L, P, value.dtype):
        raise RuntimeError("fused snippet attention needs float32 / bfloat16 value, D % 16 == 0, D <= 128, L*P <= 32")
    if any(t.dtype != torch.float32 for t in (offsets, logits, ref)):
        raise RuntimeError("offsets, logits and reference_points must be float32")
    if not (0 < n_frame <= T2):
        raise RuntimeError("n_frame must be in (0, T2]")
    _require_contiguous(offsets, "offsets")
    _require_contiguous(logits, "logits")
    _require_contiguous(spatial_shapes, "spatial_shapes")
    _require_contiguous(level_start_index, "level_start_index")
    return N, T2, T1, S, M, D, L, Lq, P


def _value_strides5(value):
    N, T2, S, M, D = value.shape
    if value.stride(4) != 1 or value.stride(3) != D or (S > 1 and value.stride(2) != M * D):
        raise RuntimeError("value tensor has to be contiguous in its (S,M,D) dims")
    st = value.stride(1) if T2 > 1 else S * M * D
    sn = value.stride(0) if N > 1 else st * T2
    return sn, st


def _ref_strides(ref):
    """(N,T1,Lq,L,2): inner (Lq,L,2) dense; batch / frame strides free (0 = broadcast)."""
    N, T1, Lq, L, _ = ref.shape
    if ref.numel() and (ref.stride(4) != 1 or ref.stride(3) != 2 or (Lq > 1 and ref.stride(2) != 2 * L)):
        ref = ref.contiguous()
    return ref, (ref.stride(0) if N > 1 else 0), (ref.stride(1) if T1 > 1 else 0)


@torch.library.custom_op("snipper_b200::snippet_forward", mutates_args=())
def snippet_forward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor,
                    offsets: Tensor, logits: Tensor, reference_points: Tensor, n_frame: int,
                    value_mask: Optional[Tensor] = None) -> Tensor:
    """``value_mask`` (bool, value's shape, contiguous) is not read here -- ``value`` must already be zero
    where it is set (``MaskedValue``); it is carried to the backward, which zeroes those elements of grad_value."""
    N, T2, T1, S, M, D, L, Lq, P = _check_snippet(value, spatial_shapes, level_start_index, offsets,
                                                  logits, reference_points, n_frame)
    sn, st = _value_strides5(value)
    ref, rsn, rst = _ref_strides(reference_points)
    out = torch.empty((N, T1, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _Launch("snippet_forward", (N, T2, T1, S, M, D, L, Lq, P), value.device):
        status = capi.lib().msda_snippet_forward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            offsets.data_ptr(), logits.data_ptr(), ref.data_ptr(), out.data_ptr(),
            N, T2, T1, int(n_frame), S, M, D, L, Lq, P, sn, st, rsn, rst, 0, 0, None, None,
            _DTYPES[value.dtype], _stream(value.device))
    capi.check(status, "msda_snippet_forward")
    return out


@snippet_forward.register_fake
def _(value, spatial_shapes, level_start_index, offsets, logits, reference_points, n_frame, value_mask=None):
    N, T2, S, M, D = value.shape
    return value.new_empty((N, offsets.shape[1], offsets.shape[2], M * D))


@torch.library.custom_op("snipper_b200::snippet_backward", mutates_args=())
def snippet_backward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor,
                     offsets: Tensor, logits: Tensor, reference_points: Tensor, grad_output: Tensor,
                     n_frame: int, value_mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    N, T2, T1, S, M, D, L, Lq, P = _check_snippet(value, spatial_shapes, level_start_index, offsets,
                                                  logits, reference_points, n_frame)
    _require_cuda(grad_output, "grad_output")
    _require_contiguous(grad_output, "grad_output")
    sn, st = _value_strides5(value)
    ref, rsn, rst = _ref_strides(reference_points)
    grad_value = torch.empty((N, T2, S, M, D), dtype=torch.float32, device=value.device)  # fp32 accumulation
    grad_offsets = torch.empty_like(offsets)
    grad_logits = torch.empty_like(logits)
    with torch.cuda.device(value.device), _Launch("snippet_backward", (N, T2, T1, S, M, D, L, Lq, P), value.device):
        status = capi.lib().msda_snippet_backward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            offsets.data_ptr(), logits.data_ptr(), ref.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_offsets.data_ptr(), grad_logits.data_ptr(),
            N, T2, T1, int(n_frame), S, M, D, L, Lq, P, sn, st, rsn, rst, 0, 0, None, None,
            _DTYPES[value.dtype], 0, _stream(value.device))
    capi.check(status, "msda_snippet_backward")
    _mask_grad_value(grad_value, value_mask, value.device)
    if value.dtype != torch.float32:
        grad_value = grad_value.to(value.dtype)
    return grad_value, grad_offsets, grad_logits


def _mask_grad_value(grad_value, value_mask, device):
    if value_mask is not None:
        # d(masked_fill)/d(value) : no gradient reaches the masked elements (fresh buffer, in place)
        if value_mask.dtype != torch.bool or value_mask.numel() != grad_value.numel() or not value_mask.is_contiguous():
            raise RuntimeError("value_mask must be a contiguous bool tensor with value's number of elements")
        with torch.cuda.device(device), _Launch("masked_zero", (grad_value.numel(),), device):
            status = capi.lib().msda_masked_zero(grad_value.data_ptr(), value_mask.data_ptr(), grad_value.numel(),
                                                 capi.MSDA_DTYPE_F32, _stream(device))
        capi.check(status, "msda_masked_zero")


@snippet_backward.register_fake
def _(value, spatial_shapes, level_start_index, offsets, logits, reference_points, grad_output, n_frame, value_mask=None):
    return (value.new_empty(value.shape), torch.empty_like(offsets), torch.empty_like(logits))


def _snippet_setup_context(ctx, inputs, output):
    value, spatial_shapes, level_start_index, offsets, logits, reference_points, n_frame, value_mask = inputs
    ctx.n_frame = n_frame
    ctx.has_mask = value_mask is not None
    if ctx.has_mask:
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets, logits, reference_points, value_mask)
    else:
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets, logits, reference_points)


def _snippet_backward_formula(ctx, grad_output):
    if ctx.has_mask:
        value, spatial_shapes, level_start_index, offsets, logits, ref, value_mask = ctx.saved_tensors
    else:
        value, spatial_shapes, level_start_index, offsets, logits, ref = ctx.saved_tensors
        value_mask = None
    gv, goff, glog = torch.ops.snipper_b200.snippet_backward(
        value, spatial_shapes, level_start_index, offsets, logits, ref, grad_output.contiguous(), ctx.n_frame,
        value_mask)
    gref = None
    if ctx.needs_input_grad[5]:
        # loc = ref + off/(W,H)  =>  dL/dref = sum_{m,p} dL/dloc = sum_{m,p} dL/doff * (W,H)
        wh = torch.stack([spatial_shapes[:, 1], spatial_shapes[:, 0]], -1).to(goff.dtype)
        gref = (goff * wh[None, None, None, None, :, None, :]).sum(dim=(3, 5))
    return gv, None, None, goff, glog, gref, None, None


snippet_forward.register_autograd(_snippet_backward_formula, setup_context=_snippet_setup_context)


# ------------------------------------------------------------------------------------------
# fused snippet op, packed projection: offsets and logits are column blocks of ONE GEMM output and
# the two Linear biases are added in-kernel (no second GEMM over the queries, no epilogue passes)
# ------------------------------------------------------------------------------------------
def _check_packed(value, spatial_shapes, level_start_index, proj, offsets_bias, logits_bias, ref, n_frame):
    for t, name in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                    (proj, "proj"), (ref, "reference_points")):
        _require_cuda(t, name)
    if value.dim() != 5 or proj.dim() != 4 or ref.dim() != 5:
        raise RuntimeError("expected value (N,T2,S,M,D), proj (N,T1,Lq,3*M*L*P), reference_points (N,T1,Lq,L,2)")
    N, T2, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Np, T1, Lq, W = proj.shape
    if Np != N or W % (3 * M * L) != 0:
        raise RuntimeError("proj must be (N,T1,Lq,3*M*L*P): [offsets (M,L,P,2) | logits (M,L,P)] per query")
    P = W // (3 * M * L)
    if tuple(ref.shape) != (N, T1, Lq, L, 2) or tuple(spatial_shapes.shape) != (L, 2):
        raise RuntimeError("reference_points must be (N,T1,Lq,L,2) and spatial_shapes (L,2)")
    if not snippet_supported(M, D, L, P, value.dtype):
        raise RuntimeError("fused snippet attention needs float32 / bfloat16 value, D % 16 == 0, D <= 128, L*P <= 32")
    if proj.dtype != torch.float32 or ref.dtype != torch.float32:
        raise RuntimeError("proj and reference_points must be float32")
    for b, n in ((offsets_bias, 2 * M * L * P), (logits_bias, M * L * P)):
        if b is not None and (not b.is_cuda or b.dtype != torch.float32 or b.numel() != n or not b.is_contiguous()):
            raise RuntimeError("biases must be contiguous float32 CUDA tensors of 2*M*L*P / M*L*P elements")
    if not (0 < n_frame <= T2):
        raise RuntimeError("n_frame must be in (0, T2]")
    _require_contiguous(proj, "proj")
    _require_contiguous(spatial_shapes, "spatial_shapes")
    _require_contiguous(level_start_index, "level_start_index")
    return N, T2, T1, S, M, D, L, Lq, P


def _ptr(t):
    return None if t is None else t.data_ptr()


@torch.library.custom_op("snipper_b200::snippet_forward_packed", mutates_args=())
def snippet_forward_packed(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, proj: Tensor,
                           offsets_bias: Optional[Tensor], logits_bias: Optional[Tensor],
                           reference_points: Tensor, n_frame: int, value_mask: Optional[Tensor] = None) -> Tensor:
    N, T2, T1, S, M, D, L, Lq, P = _check_packed(value, spatial_shapes, level_start_index, proj, offsets_bias,
                                                 logits_bias, reference_points, n_frame)
    sn, st = _value_strides5(value)
    ref, rsn, rst = _ref_strides(reference_points)
    mlp = M * L * P
    out = torch.empty((N, T1, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device), _Launch("snippet_forward", (N, T2, T1, S, M, D, L, Lq, P), value.device):
        status = capi.lib().msda_snippet_forward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            proj.data_ptr(), proj.data_ptr() + 4 * 2 * mlp, ref.data_ptr(), out.data_ptr(),
            N, T2, T1, int(n_frame), S, M, D, L, Lq, P, sn, st, rsn, rst, 3 * mlp, 3 * mlp,
            _ptr(offsets_bias), _ptr(logits_bias), _DTYPES[value.dtype], _stream(value.device))
    capi.check(status, "msda_snippet_forward")
    return out


@snippet_forward_packed.register_fake
def _(value, spatial_shapes, level_start_index, proj, offsets_bias, logits_bias, reference_points, n_frame,
      value_mask=None):
    N, T2, S, M, D = value.shape
    return value.new_empty((N, proj.shape[1], proj.shape[2], M * D))


@torch.library.custom_op("snipper_b200::snippet_backward_packed", mutates_args=())
def snippet_backward_packed(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, proj: Tensor,
                            offsets_bias: Optional[Tensor], logits_bias: Optional[Tensor],
                            reference_points: Tensor, grad_output: Tensor, n_frame: int,
                            value_mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Returns (grad_value, grad_proj): grad_proj has proj's layout [grad_offsets | grad_logits]."""
    N, T2, T1, S, M, D, L, Lq, P = _check_packed(value, spatial_shapes, level_start_index, proj, offsets_bias,
                                                 logits_bias, reference_points, n_frame)
    _require_cuda(grad_output, "grad_output")
    _require_contiguous(grad_output, "grad_output")
    sn, st = _value_strides5(value)
    ref, rsn, rst = _ref_strides(reference_points)
    mlp = M * L * P
    grad_value = torch.empty((N, T2, S, M, D), dtype=torch.float32, device=value.device)  # fp32 accumulation
    grad_proj = torch.empty_like(proj)
    with torch.cuda.device(value.device), _Launch("snippet_backward", (N, T2, T1, S, M, D, L, Lq, P), value.device):
        status = capi.lib().msda_snippet_backward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            proj.data_ptr(), proj.data_ptr() + 4 * 2 * mlp, ref.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_proj.data_ptr(), grad_proj.data_ptr() + 4 * 2 * mlp,
            N, T2, T1, int(n_frame), S, M, D, L, Lq, P, sn, st, rsn, rst, 3 * mlp, 3 * mlp,
            _ptr(offsets_bias), _ptr(logits_bias), _DTYPES[value.dtype], 0, _stream(value.device))
    capi.check(status, "msda_snippet_backward")
    _mask_grad_value(grad_value, value_mask, value.device)
    if value.dtype != torch.float32:
        grad_value = grad_value.to(value.dtype)
    return grad_value, grad_proj


@snippet_backward_packed.register_fake
def _(value, spatial_shapes, level_start_index, proj, offsets_bias, logits_bias, reference_points, grad_output,
      n_frame, value_mask=None):
    return value.new_empty(value.shape), torch.empty_like(proj)


def _packed_setup_context(ctx, inputs, output):
    value, spatial_shapes, level_start_index, proj, ob, lb, reference_points, n_frame, value_mask = inputs
    ctx.n_frame = n_frame
    ctx.flags = (ob is not None, lb is not None, value_mask is not None)
    ctx.save_for_backward(*[t for t in (value, spatial_shapes, level_start_index, proj, reference_points, ob, lb,
                                        value_mask) if t is not None])


def _packed_backward_formula(ctx, grad_output):
    saved = list(ctx.saved_tensors)
    value, spatial_shapes, level_start_index, proj, ref = saved[:5]
    rest = saved[5:]
    ob = rest.pop(0) if ctx.flags[0] else None
    lb = rest.pop(0) if ctx.flags[1] else None
    value_mask = rest.pop(0) if ctx.flags[2] else None
    gv, gproj = torch.ops.snipper_b200.snippet_backward_packed(
        value, spatial_shapes, level_start_index, proj, ob, lb, ref, grad_output.contiguous(), ctx.n_frame, value_mask)
    N, T2, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    mlp = proj.shape[-1] // 3
    P = mlp // (M * L)
    gob = glb = gref = None
    if ctx.needs_input_grad[4] or ctx.needs_input_grad[5]:
        col = gproj.sum(dim=(0, 1, 2))                       # bias gradients = column sums of the projection gradient
        gob = col[:2 * mlp] if ctx.needs_input_grad[4] else None
        glb = col[2 * mlp:] if ctx.needs_input_grad[5] else None
    if ctx.needs_input_grad[6]:
        goff = gproj[..., :2 * mlp].view(proj.shape[0], proj.shape[1], proj.shape[2], M, L, P, 2)
        wh = torch.stack([spatial_shapes[:, 1], spatial_shapes[:, 0]], -1).to(goff.dtype)
        gref = (goff * wh[None, None, None, None, :, None, :]).sum(dim=(3, 5))
    return gv, None, None, gproj, gob, glb, gref, None, None


snippet_forward_packed.register_autograd(_packed_backward_formula, setup_context=_packed_setup_context)
