/* TEST INFRASTRUCTURE ONLY -- force-included (-include) when compiling the reference's vendored
 * CUDA op IN PLACE from /root/reference/models/ops/src, unmodified.
 *
 * The reference calls AT_DISPATCH_FLOATING_TYPES(value.type(), ...)
 * (models/ops/src/cuda/ms_deform_attn_cuda.cu:64,134).  torch >= 2.x dropped the
 * ::detail::scalar_type(const DeprecatedTypeProperties&) overload that made this compile;
 * restoring it here lets the untouched sources build against torch 2.11. */
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
