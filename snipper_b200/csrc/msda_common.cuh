// Device helpers shared by the per-call and the fused snippet kernels (sm_100a).
//
// Data layout in HBM (all row-major, as the reference op):
//   value   (N, S, M, D)          one "cell" = the D channels of head m at flattened pixel s
//   loc     (N, Lq, M, L, P, 2)   (x, y) normalised;  attn (N, Lq, M, L, P)
//   out     (N, Lq, M, D)
// A "pair" is one (n, q, m); it owns L*P "samples" (one per level x point); a sample touches
// up to four cells.  Math follows reference ms_deform_im2col_cuda.cuh:33-84 (forward),
// :87-159 (backward), :272-296 (pixel convention / range test).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_internal.h"

namespace msda {

// Level table staged once per CTA in shared memory (the int64 device tensors are never
// re-read per level/point as the reference does, ms_deform_im2col_cuda.cuh:272-281).
struct LevelTable {
    int H[kMaxLevels];
    int W[kMaxLevels];
    int start[kMaxLevels];
};

// A level whose slab [start, start + H*W) does not lie inside the S cells of a batch item is given an
// empty extent, so every corner on it is invalid: inconsistent spatial_shapes / level_start_index (the
// reference asserts (H*W).sum() == Len_in on the host, ms_deform_attn.py:112 -- a device sync per call)
// can never turn into an out-of-bounds gather or reduction.
__device__ __forceinline__ void load_level_table(LevelTable &t, const int64_t *__restrict__ shapes,
                                                 const int64_t *__restrict__ lsi, int L, int S)
{
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const int64_t H = shapes[2 * l], W = shapes[2 * l + 1], start = lsi[l];
        const bool ok = H >= 0 && W >= 0 && start >= 0 && H <= S && W <= S && start + H * W <= (int64_t)S;
        t.H[l] = ok ? (int)H : 0;
        t.W[l] = ok ? (int)W : 0;
        t.start[l] = ok ? (int)start : 0;
    }
}

// One bilinear sample: fractional parts and the four cell indices (start + y*W + x) or -1
// for a corner that falls outside the level / a sample outside (-1,W)x(-1,H).
template <typename real>
struct Sample {
    int cell[4];
    int base;   // start + y0*W + x0 (may point outside the level when x0 or y0 is -1)
    int mask;   // bit k set <=> corner k valid; 0 for an inactive sample
    real lx, ly;
};

template <typename real>
__device__ __forceinline__ Sample<real> make_sample(real u, real v, int H, int W, int start)
{
    Sample<real> s;
    // explicit single-rounding FMA (what nvcc makes of the reference's `loc_w * spatial_w - 0.5`,
    // ms_deform_im2col_cuda.cuh:285-286); the CPU oracle does the same, so floor() agrees bit for bit
    const real x = fma(u, (real)W, (real)-0.5);
    const real y = fma(v, (real)H, (real)-0.5);
    const bool active = (x > (real)-1) && (y > (real)-1) && (x < (real)W) && (y < (real)H);
    const real fx = floor(x), fy = floor(y);
    s.lx = x - fx;
    s.ly = y - fy;
    const int x0 = active ? (int)fx : 0, y0 = active ? (int)fy : 0;
    const bool vx0 = active && x0 >= 0, vx1 = active && x0 + 1 <= W - 1;
    const bool vy0 = y0 >= 0, vy1 = y0 + 1 <= H - 1;
    const int base = start + y0 * W + x0;
    s.cell[0] = (vy0 && vx0) ? base : -1;
    s.cell[1] = (vy0 && vx1) ? base + 1 : -1;
    s.cell[2] = (vy1 && vx0) ? base + W : -1;
    s.cell[3] = (vy1 && vx1) ? base + W + 1 : -1;
    s.mask = (int)(vy0 && vx0) | ((int)(vy0 && vx1) << 1) | ((int)(vy1 && vx0) << 2) | ((int)(vy1 && vx1) << 3);
    s.base = active ? base : 0;
    if (!active) { s.lx = (real)0; s.ly = (real)0; }
    return s;
}

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &v)
{
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// 128-bit vector reduction to global memory (sm_90+): one L2 atomic op per 16 bytes
// instead of four scalar REDs (the reference issues scalar atomicAdd per channel,
// ms_deform_im2col_cuda.cuh:125-152).
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :
                 : "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b)
{
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

}  // namespace msda
