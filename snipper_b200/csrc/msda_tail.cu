// Layer tail of the deformable encoder / decoder layers around MSDeformAttn, sm_100a.
//
// The reference finishes every attention block and every FFN block with
//     x = x + dropout(Linear(...))      (Linear bias included)
//     x = LayerNorm(x)
// (models/deformable_transformer.py:204-205 and :194-198 encoder; :294-295, :286-287, :270-274 decoder) and starts the
// next attention block with `with_pos_embed(x, pos) = x + pos` (:188-190, :202, :292).  In stock PyTorch that is a
// bias epilogue, an add, a LayerNorm and another add: five passes over the (N*T*Lq, C) activations.  This kernel
// does it in ONE pass:
//
//     out      = LayerNorm(residual + y + bias) * gamma + beta
//     out_pos  = out + pos                         (optional second output: the next block's query)
//
// y is the raw GEMM output (bias NOT yet added).  Inference-only (dropout is the identity in eval mode); the
// opt-in layer forwards (snipper_b200/layers.py) keep the stock ops whenever autograd is recording.
//
// One warp per row, VPL float4 per lane (C = 128 * VPL), the row lives in registers: mean and variance are a
// two-pass computation over registers (sum, then sum of squared deviations), reduced with shuffles -- HBM traffic is
// exactly the algorithmic bytes: read y + residual (+ pos), write out (+ out_pos).
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_internal.h"

namespace msda {

namespace {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int VPL>
__global__ void __launch_bounds__(256)
layer_tail_kernel(const float *__restrict__ y, const float *__restrict__ bias, const float *__restrict__ residual,
                  const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ pos,
                  float *__restrict__ out, float *__restrict__ out_pos, int64_t rows, float eps)
{
    constexpr int C = 128 * VPL;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4 *y4 = reinterpret_cast<const float4 *>(y + row * C);
    const float4 *r4 = reinterpret_cast<const float4 *>(residual + row * C);
    float4 x[VPL];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        const float4 a = __ldg(y4 + c4), r = __ldg(r4 + c4);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        // reference order: (Linear output incl. bias) first, then the residual add
        x[k] = make_float4(r.x + (a.x + b.x), r.y + (a.y + b.y), r.z + (a.z + b.z), r.w + (a.w + b.w));
        sum += (x[k].x + x[k].y) + (x[k].z + x[k].w);
    }
    const float mean = warp_sum(sum) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const float dx = x[k].x - mean, dy = x[k].y - mean, dz = x[k].z - mean, dw = x[k].w - mean;
        sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / C) + eps);
    float4 *o4 = reinterpret_cast<float4 *>(out + row * C);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + c4);
        float4 o;
        o.x = (x[k].x - mean) * rstd * g.x + b.x;
        o.y = (x[k].y - mean) * rstd * g.y + b.y;
        o.z = (x[k].z - mean) * rstd * g.z + b.z;
        o.w = (x[k].w - mean) * rstd * g.w + b.w;
        o4[c4] = o;
        if (out_pos != nullptr) {
            const float4 p = __ldg(reinterpret_cast<const float4 *>(pos + row * C) + c4);
            reinterpret_cast<float4 *>(out_pos + row * C)[c4] = make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w);
        }
    }
}

}  // namespace

bool layer_tail_ok(int cols) { return cols > 0 && cols % 128 == 0 && cols <= 1024; }

cudaError_t launch_layer_tail(const float *y, const float *bias, const float *residual, const float *gamma,
                              const float *beta, const float *pos, float *out, float *out_pos, int64_t rows,
                              int cols, float eps, cudaStream_t stream)
{
    if (rows == 0) return cudaSuccess;
    constexpr int kWarps = 8;
    const int64_t blocks = (rows + kWarps - 1) / kWarps;
    if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
#define MSDA_TAIL(V)                                                                                                \
    case V:                                                                                                         \
        layer_tail_kernel<V><<<(unsigned)blocks, 32 * kWarps, 0, stream>>>(y, bias, residual, gamma, beta, pos, out, \
                                                                          out_pos, rows, eps);                      \
        break;
    switch (cols / 128) {
        MSDA_TAIL(1) MSDA_TAIL(2) MSDA_TAIL(3) MSDA_TAIL(4) MSDA_TAIL(5) MSDA_TAIL(6) MSDA_TAIL(7) MSDA_TAIL(8)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_TAIL
    return cudaGetLastError();
}

}  // namespace msda
