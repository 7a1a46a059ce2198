// In-place masked zero-fill: data[i] = 0 where mask[i] != 0  (sm_100a).
//
// Replaces the out-of-place `value.masked_fill(input_padding_mask, 0)` of the reference module
// (models/ops/modules/ms_deform_attn.py:116-117) and its autograd mirror on grad_value.  The
// reference's pass reads and re-writes the whole (N,T,S,C) tensor (2 x 60.7 MB per layer at the
// headline config) although padding masks are mostly false; this kernel reads only the mask
// (1 byte per element) and WRITES only where it is set, so its traffic is the mask plus the padded
// region.  16 elements per thread: one 128-bit load of mask bytes, then -- only if any byte is set --
// predicated 128-bit read-modify-writes of the corresponding data.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_internal.h"

namespace msda {

template <typename T>  // float or __nv_bfloat16 (any 4- or 2-byte type: only zeros are written)
__global__ void __launch_bounds__(256)
masked_zero_kernel(T *__restrict__ data, const uint8_t *__restrict__ mask, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 16 <= n) {
            const uint4 mk = __ldg(reinterpret_cast<const uint4 *>(mask + i));
            if ((mk.x | mk.y | mk.z | mk.w) == 0u) continue;
            const unsigned w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (w[k] == 0u) continue;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((w[k] >> (8 * b)) & 0xffu) data[i + 4 * k + b] = T(0.f);
            }
        } else {
            for (int64_t j = i; j < n; ++j)
                if (mask[j]) data[j] = T(0.f);
        }
    }
}

cudaError_t launch_masked_zero(void *data, const uint8_t *mask, int64_t n, int elem_bytes, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int64_t threads_needed = (n + 15) / 16;
    int64_t blocks = (threads_needed + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (elem_bytes == 4)
        masked_zero_kernel<float><<<(int)blocks, 256, 0, stream>>>(static_cast<float *>(data), mask, n);
    else if (elem_bytes == 2)
        masked_zero_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, stream>>>(static_cast<__nv_bfloat16 *>(data), mask, n);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace msda
