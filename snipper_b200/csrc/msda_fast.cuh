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

// Backward work of one sample for this lane on one value frame: scatter w_k*A*G into grad_value
// (vector reductions) and accumulate the three per-sample partial dot products.
template <int CSB, bool SCATTER>
__device__ __forceinline__ void gather_scatter(const SampleMeta mt, float lx, float ly, const float4 ga,
                                               const float4 g, const char *__restrict__ p0, char *gp0,
                                               int runtime_csb, float &pa, float &px, float &py)
{
    const int csb = cell_stride_bytes<CSB>(runtime_csb);
    const ptrdiff_t o0 = (ptrdiff_t)mt.off;
    const ptrdiff_t o2 = o0 + (mt.wm & 0x0fffffffu);
    const float hx = 1.f - lx, hy = 1.f - ly;
    const float w0 = hy * hx, w1 = hy * lx, w2 = ly * hx, w3 = ly * lx;
    float4 v0, v1, v2, v3;
    if (mt.wm >= kAllCorners) {
        v0 = __ldg(reinterpret_cast<const float4 *>(p0 + o0));
        v1 = __ldg(reinterpret_cast<const float4 *>(p0 + o0 + csb));
        v2 = __ldg(reinterpret_cast<const float4 *>(p0 + o2));
        v3 = __ldg(reinterpret_cast<const float4 *>(p0 + o2 + csb));
        if (SCATTER) {
            red_add_v4(reinterpret_cast<float *>(gp0 + o0), w0 * ga.x, w0 * ga.y, w0 * ga.z, w0 * ga.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o0 + csb), w1 * ga.x, w1 * ga.y, w1 * ga.z, w1 * ga.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o2), w2 * ga.x, w2 * ga.y, w2 * ga.z, w2 * ga.w);
            red_add_v4(reinterpret_cast<float *>(gp0 + o2 + csb), w3 * ga.x, w3 * ga.y, w3 * ga.z, w3 * ga.w);
        }
    } else {
        const unsigned mask = mt.wm >> 28;
        if (mask == 0u) return;  // inactive sample contributes nothing anywhere
        v0 = make_float4(0.f, 0.f, 0.f, 0.f); v1 = v0; v2 = v0; v3 = v0;
        if (mask & 1u) v0 = __ldg(reinterpret_cast<const float4 *>(p0 + o0));
        if (mask & 2u) v1 = __ldg(reinterpret_cast<const float4 *>(p0 + o0 + csb));
        if (mask & 4u) v2 = __ldg(reinterpret_cast<const float4 *>(p0 + o2));
        if (mask & 8u) v3 = __ldg(reinterpret_cast<const float4 *>(p0 + o2 + csb));
        if (SCATTER) {
            if (mask & 1u) red_add_v4(reinterpret_cast<float *>(gp0 + o0), w0 * ga.x, w0 * ga.y, w0 * ga.z, w0 * ga.w);
            if (mask & 2u) red_add_v4(reinterpret_cast<float *>(gp0 + o0 + csb), w1 * ga.x, w1 * ga.y, w1 * ga.z, w1 * ga.w);
            if (mask & 4u) red_add_v4(reinterpret_cast<float *>(gp0 + o2), w2 * ga.x, w2 * ga.y, w2 * ga.z, w2 * ga.w);
            if (mask & 8u) red_add_v4(reinterpret_cast<float *>(gp0 + o2 + csb), w3 * ga.x, w3 * ga.y, w3 * ga.z, w3 * ga.w);
        }
    }
    // val = sum w_k v_k ; dval/dx = hy (v1 - v0) + ly (v3 - v2) ; dval/dy = hx (v2 - v0) + lx (v3 - v1)
    float4 val, dxv, dyv;
    val.x = w0 * v0.x + w1 * v1.x + w2 * v2.x + w3 * v3.x;
    val.y = w0 * v0.y + w1 * v1.y + w2 * v2.y + w3 * v3.y;
    val.z = w0 * v0.z + w1 * v1.z + w2 * v2.z + w3 * v3.z;
    val.w = w0 * v0.w + w1 * v1.w + w2 * v2.w + w3 * v3.w;
    dxv.x = hy * (v1.x - v0.x) + ly * (v3.x - v2.x);
    dxv.y = hy * (v1.y - v0.y) + ly * (v3.y - v2.y);
    dxv.z = hy * (v1.z - v0.z) + ly * (v3.z - v2.z);
    dxv.w = hy * (v1.w - v0.w) + ly * (v3.w - v2.w);
    dyv.x = hx * (v2.x - v0.x) + lx * (v3.x - v1.x);
    dyv.y = hx * (v2.y - v0.y) + lx * (v3.y - v1.y);
    dyv.z = hx * (v2.z - v0.z) + lx * (v3.z - v1.z);
    dyv.w = hx * (v2.w - v0.w) + lx * (v3.w - v1.w);
    pa += dot4(g, val);
    px += dot4(g, dxv);
    py += dot4(g, dyv);
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
