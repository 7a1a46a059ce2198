// Shared machinery of the fast fp32 kernels (per-call and fused snippet), sm_100a.
//
// CTA = TILE of PAIRS consecutive queries x ONE head; thread = (query-in-tile, 16-byte channel
// chunk) -- LANES = D/4 lanes per query, so one warp-wide LDG.128 covers whole 4*D-byte head
// slices and every gathered cell is consumed as full 32-byte sectors.
//
//   phase 1  one thread per SAMPLE (query, level, point): bilinear set-up done once and parked
//            in shared memory as
//                SampleMeta {byte offset of the (y0,x0) cell, row stride | corner mask}  (8 B)
//                float4     per-corner weights (forward) or {lx, ly, A, -} (backward)   (16 B)
//            (the reference redoes this set-up in every one of the D channel threads)
//   phase 2  each lane walks its query's samples: LDS.64 + LDS.128 broadcast, then -- when all
//            four corners are valid, the common case -- 4 unpredicated LDG.128 whose "+1 cell"
//            addresses are immediate offsets (the cell stride is a template constant for
//            Snipper's M*D = 384), 16 FFMA.  Border samples take a predicated slow path.
//
// Why this shape: ncu on the first versions (profiles/r01_run3_*, r01_run6_*) showed the gather
// is limited first by INSTRUCTION ISSUE (64-bit address arithmetic, zero-filling registers for
// predicated loads, per-corner predicate tests: ~93 instructions per lane-point for 16 useful
// FFMA + 4 LDG) and then by the L1 data pipe (two 128-byte wavefronts per 192-byte head slice);
// HBM traffic is only the compulsory ~42 MB per call.
#pragma once

#include "msda_common.cuh"

namespace msda {

struct __align__(8) SampleMeta {
    int off;        // byte offset of cell (y0,x0) relative to (batch base + head slice), may be "virtual"
    unsigned wm;    // (row stride in bytes) | corner mask << 28
};

constexpr unsigned kAllCorners = 0xF0000000u;

// exact i / d for 0 <= i*d < 2^24 with magic = ceil(2^24 / d)  (host: fast_magic)
__device__ __forceinline__ int fast_div(int i, unsigned magic) { return (int)(((unsigned)i * magic) >> 24); }
inline unsigned fast_magic(int d) { return (unsigned)(((1u << 24) + (unsigned)d - 1u) / (unsigned)d); }
__device__ __forceinline__ unsigned fast_magic_dev(int d) { return ((1u << 24) + (unsigned)d - 1u) / (unsigned)d; }

// Build the 8-byte meta word of a sample. cell_bytes = M*D*4 (bytes between consecutive cells).
__device__ __forceinline__ SampleMeta make_meta(const Sample<float> &s, int W, int cell_bytes)
{
    SampleMeta m;
    m.off = s.base * cell_bytes;
    m.wm = (unsigned)(W * cell_bytes) | ((unsigned)s.mask << 28);
    return m;
}

__device__ __forceinline__ SampleMeta empty_meta()
{
    SampleMeta m;
    m.off = 0; m.wm = 0u;
    return m;
}

template <int CSB>
__device__ __forceinline__ int cell_stride_bytes(int runtime_csb) { return CSB > 0 ? CSB : runtime_csb; }

// Forward gather of one sample for this lane: acc += sum_k w_k * V[corner_k]
template <int CSB>
__device__ __forceinline__ void gather_fma(float4 &acc, const SampleMeta mt, const float4 w,
                                           const char *__restrict__ p0, int runtime_csb)
{
    const int csb = cell_stride_bytes<CSB>(runtime_csb);
    const char *a0 = p0 + (ptrdiff_t)mt.off;
    const char *a2 = a0 + (mt.wm & 0x0fffffffu);
    if (mt.wm >= kAllCorners) {
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(a0));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(a0 + csb));
        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(a2));
        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(a2 + csb));
        fma4(acc, w.x, v0);
        fma4(acc, w.y, v1);
        fma4(acc, w.z, v2);
        fma4(acc, w.w, v3);
    } else {
        const unsigned mask = mt.wm >> 28;
        if (mask & 1u) fma4(acc, w.x, __ldg(reinterpret_cast<const float4 *>(a0)));
        if (mask & 2u) fma4(acc, w.y, __ldg(reinterpret_cast<const float4 *>(a0 + csb)));
        if (mask & 4u) fma4(acc, w.z, __ldg(reinterpret_cast<const float4 *>(a2)));
        if (mask & 8u) fma4(acc, w.w, __ldg(reinterpret_cast<const float4 *>(a2 + csb)));
    }
}

// Two samples per step: when both have all four corners, the eight gathers are issued before any
// of the 32 FFMAs so twice as many loads are in flight per warp.
template <int CSB>
__device__ __forceinline__ void gather_fma2(float4 &acc, const SampleMeta m0, const float4 w0, const SampleMeta m1,
                                            const float4 w1, const char *__restrict__ p0, int runtime_csb)
{
    const int csb = cell_stride_bytes<CSB>(runtime_csb);
    if (m0.wm >= kAllCorners && m1.wm >= kAllCorners) {
        const char *a0 = p0 + (ptrdiff_t)m0.off;
        const char *a2 = a0 + (m0.wm & 0x0fffffffu);
        const char *b0 = p0 + (ptrdiff_t)m1.off;
        const char *b2 = b0 + (m1.wm & 0x0fffffffu);
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(a0));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(a0 + csb));
        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(a2));
        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(a2 + csb));
        const float4 u0 = __ldg(reinterpret_cast<const float4 *>(b0));
        const float4 u1 = __ldg(reinterpret_cast<const float4 *>(b0 + csb));
        const float4 u2 = __ldg(reinterpret_cast<const float4 *>(b2));
        const float4 u3 = __ldg(reinterpret_cast<const float4 *>(b2 + csb));
        fma4(acc, w0.x, v0); fma4(acc, w0.y, v1); fma4(acc, w0.z, v2); fma4(acc, w0.w, v3);
        fma4(acc, w1.x, u0); fma4(acc, w1.y, u1); fma4(acc, w1.z, u2); fma4(acc, w1.w, u3);
    } else {
        gather_fma<CSB>(acc, m0, w0, p0, runtime_csb);
        gather_fma<CSB>(acc, m1, w1, p0, runtime_csb);
    }
}

// Same, over `nf` consecutive value frames (fused snippet kernel): the sample set-up is shared by all
// neighbour frames, so the fast/slow decision is taken once and the frame loop is branch-free,
// which lets the loads of two frames overlap.
template <int CSB>
__device__ __forceinline__ void gather_fma_frames(float4 &acc, const SampleMeta mt, const float4 w,
                                                  const char *__restrict__ pf, int64_t frame_bytes, int nf,
                                                  int runtime_csb)
{
    const int csb = cell_stride_bytes<CSB>(runtime_csb);
    const char *a0 = pf + (ptrdiff_t)mt.off;
    const ptrdiff_t row = (ptrdiff_t)(mt.wm & 0x0fffffffu);
    if (mt.wm >= kAllCorners) {
#pragma unroll 2
        for (int f = 0; f < nf; ++f, a0 += frame_bytes) {
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(a0));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(a0 + csb));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(a0 + row));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(a0 + row + csb));
            fma4(acc, w.x, v0);
            fma4(acc, w.y, v1);
            fma4(acc, w.z, v2);
            fma4(acc, w.w, v3);
        }
    } else {
        const unsigned mask = mt.wm >> 28;
        if (mask == 0u) return;
        for (int f = 0; f < nf; ++f, a0 += frame_bytes) {
            if (mask & 1u) fma4(acc, w.x, __ldg(reinterpret_cast<const float4 *>(a0)));
            if (mask & 2u) fma4(acc, w.y, __ldg(reinterpret_cast<const float4 *>(a0 + csb)));
            if (mask & 4u) fma4(acc, w.z, __ldg(reinterpret_cast<const float4 *>(a0 + row)));
            if (mask & 8u) fma4(acc, w.w, __ldg(reinterpret_cast<const float4 *>(a0 + row + csb)));
        }
    }
}

// Per-sample scalars a backward lane needs, derived once per sample from {lx, ly, A}
// (outside the neighbour-frame loop of the fused kernel).
struct BwdWeights {
    float lx, ly, hx, hy;
    float w0, w1, w2, w3;      // bilinear corner weights
    float a0, a1, a2, a3;      // w_k * A  (what is scattered into grad_value, times G)
};

__device__ __forceinline__ BwdWeights make_bwd_weights(float lx, float ly, float at)
{
    BwdWeights b;
    b.lx = lx; b.ly = ly; b.hx = 1.f - lx; b.hy = 1.f - ly;
    b.w0 = b.hy * b.hx; b.w1 = b.hy * lx; b.w2 = ly * b.hx; b.w3 = ly * lx;
    b.a0 = b.w0 * at; b.a1 = b.w1 * at; b.a2 = b.w2 * at; b.a3 = b.w3 * at;
    return b;
}

// Backward work of one sample for this lane on one value frame: scatter w_k*A*G into grad_value
// (vector reductions) and accumulate the three per-sample partial dot products.
// "Dot first": d_k = <G, V_k> over this lane's 4 channels, then
//     <G, val>      = sum_k w_k d_k
//     <G, dval/dx>  = hy (d1 - d0) + ly (d3 - d2)
//     <G, dval/dy>  = hx (d2 - d0) + lx (d3 - d1)
// -- 28 FP instructions instead of the 60 of forming val / dval per channel and dotting after.
template <int CSB, bool SCATTER>
__device__ __forceinline__ void gather_scatter(const SampleMeta mt, const BwdWeights &b, const float4 g,
                                               const char *__restrict__ p0, char *gp0,
                                               int runtime_csb, float &pa, float &px, float &py)
{
    const int csb = cell_stride_bytes<CSB>(runtime_csb);
    const ptrdiff_t o0 = (ptrdiff_t)mt.off;
    const ptrdiff_t o2 = o0 + (mt.wm & 0x0fffffffu);
    float d0, d1, d2, d3;
    if (mt.wm >= kAllCorners) {
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p0 + o0));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p0 + o0 + csb));
        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(p0 + o2));
        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(p0 + o2 + csb));
        if (SCATTER) {
            red_add_v4(reinterpret_cast<float *>(gp0 + o0), b.a0 * g.x, b.a0 * g.y, b.a0 * g.z, b.a0 * g.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o0 + csb), b.a1 * g.x, b.a1 * g.y, b.a1 * g.z, b.a1 * g.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o2), b.a2 * g.x, b.a2 * g.y, b.a2 * g.z, b.a2 * g.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o2 + csb), b.a3 * g.x, b.a3 * g.y, b.a3 * g.z, b.a3 * g.w);
        }
        d0 = dot4(g, v0); d1 = dot4(g, v1); d2 = dot4(g, v2); d3 = dot4(g, v3);
    } else {
        const unsigned mask = mt.wm >> 28;
        if (mask == 0u) return;  // inactive sample contributes nothing anywhere
        d0 = 0.f; d1 = 0.f; d2 = 0.f; d3 = 0.f;
        if (mask & 1u) d0 = dot4(g, __ldg(reinterpret_cast<const float4 *>(p0 + o0)));
        if (mask & 2u) d1 = dot4(g, __ldg(reinterpret_cast<const float4 *>(p0 + o0 + csb)));
        if (mask & 4u) d2 = dot4(g, __ldg(reinterpret_cast<const float4 *>(p0 + o2)));
        if (mask & 8u) d3 = dot4(g, __ldg(reinterpret_cast<const float4 *>(p0 + o2 + csb)));
        if (SCATTER) {
            if (mask & 1u) red_add_v4(reinterpret_cast<float *>(gp0 + o0), b.a0 * g.x, b.a0 * g.y, b.a0 * g.z, b.a0 * g.w);
            if (mask & 2u) red_add_v4(reinterpret_cast<float *>(gp0 + o0 + csb), b.a1 * g.x, b.a1 * g.y, b.a1 * g.z, b.a1 * g.w);
            if (mask & 4u) red_add_v4(reinterpret_cast<float *>(gp0 + o2), b.a2 * g.x, b.a2 * g.y, b.a2 * g.z, b.a2 * g.w);
            if (mask & 8u) red_add_v4(reinterpret_cast<float *>(gp0 + o2 + csb), b.a3 * g.x, b.a3 * g.y, b.a3 * g.z, b.a3 * g.w);
        }
    }
    pa = fmaf(b.w0, d0, fmaf(b.w1, d1, fmaf(b.w2, d2, fmaf(b.w3, d3, pa))));
    px = fmaf(b.hy, d1 - d0, fmaf(b.ly, d3 - d2, px));
    py = fmaf(b.hx, d2 - d0, fmaf(b.lx, d3 - d1, py));
}

// Queries per CTA tile for D = 48.  The tuned default is 16; when that would leave the GPU with
// fewer than ~3 CTAs per SM (decoder: Lq = 60) halve the tile so twice as many CTAs are in flight
// -- those launches are latency-bound, not bandwidth-bound.
inline int pick_pairs_d48(int configured, int Lq, int M, int nz)
{
    if (configured != 16) return configured;
    const long long ctas = (long long)((Lq + 15) / 16) * M * nz;
    return ctas < 148 * 3 ? 8 : 16;
}

// sum over the 4 lanes of a shuffle sub-group (full-warp participation required)
__device__ __forceinline__ void subgroup_sum3(float &a, float &b, float &c)
{
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
}

}  // namespace msda
