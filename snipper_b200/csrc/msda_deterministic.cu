// Deterministic (atomics-free accumulation, bit-reproducible) backward for grad_value, sm_100a.
//
// The fast backward scatters w_k*A*G into grad_value with floating-point reductions, whose
// order -- and therefore rounding -- changes run to run (the reference has the same property,
// ms_deform_im2col_cuda.cuh:125-152).  This mode turns the scatter into a gather in two passes:
//
//   pass 1  count   one thread per sample: integer-count the valid corners per destination
//                   cell (n, s, m)                         -> count[N*S*M]   (integer atomics)
//           scan    exclusive prefix sum                   -> start[N*S*M+1]
//           fill    one thread per sample: append (corner id, w_k*A) to its cell's list
//                   (slot order inside a list is arbitrary)
//   pass 2  reduce  16 lanes per destination cell: rank-sort the cell's list by corner id
//                   (a canonical order), then sum w * grad_out[pair, :] in that order and WRITE
//                   grad_value -- every cell is written exactly once by one thread group.
//
// grad_sampling_loc / grad_attn_weight never needed atomics; they come from the regular kernels
// run with the scatter switched off.  Workspace: 3 ints per cell + 8 bytes per corner record.
#include "msda_common.cuh"
#include "msda_internal.h"

namespace msda {

namespace {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

struct __align__(8) CornerRec {
    int id;    // (sample index) * 4 + corner
    float w;   // bilinear weight * attention weight
};

struct DetLayout {
    size_t count, cursor, long_cells, start, blocksums, records, total;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// frames of grad_value: one per batch item, or the fused layer's slots (see OpDims)
inline int64_t dest_frames(const OpDims &d) { return d.frame_q > 0 ? (int64_t)(d.N / d.frame_q) * d.frame_slots : d.N; }

DetLayout det_layout(const OpDims &d)
{
    const size_t cells = (size_t)dest_frames(d) * d.S * d.M;
    const size_t corners = (size_t)d.N * d.Lq * d.M * d.L * d.P * 4;
    const size_t nblocks = (cells + kScanTile - 1) / kScanTile;
    DetLayout l;
    size_t off = 0;
    l.count = off; off = align256(off + sizeof(int) * cells);
    l.cursor = off; off = align256(off + sizeof(int) * cells);
    l.long_cells = off; off = align256(off + sizeof(int) * (cells + 1));  // queue of over-long lists
    l.start = off; off = align256(off + sizeof(int) * (cells + 1));
    l.blocksums = off; off = align256(off + sizeof(int) * (nblocks + 1));
    l.records = off; off = align256(off + sizeof(CornerRec) * corners);
    l.total = off;
    return l;
}

// FILL == false: count corners per cell.  FILL == true: append records.
template <bool FILL>
__global__ void __launch_bounds__(256)
det_count_fill_kernel(const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                      const float *__restrict__ loc, const float *__restrict__ attn,
                      int *__restrict__ count, int *__restrict__ cursor, const int *__restrict__ start,
                      CornerRec *__restrict__ records, int S, int M, int L, int P, int Lq, int64_t total_samples,
                      int frame_q, int frame_local, int frame_slots)
{
    __shared__ LevelTable lv;
    load_level_table(lv, shapes, lsi, L, S);
    __syncthreads();
    const int LP = L * P;
    for (int64_t si = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; si < total_samples;
         si += (int64_t)gridDim.x * blockDim.x) {
        const int lp = (int)(si % LP);
        const int64_t pair = si / LP;
        const int m = (int)(pair % M);
        const int64_t n = pair / ((int64_t)Lq * M);
        const int l = lp / P;
        const int64_t frame = frame_q > 0 ? (n / frame_q) * frame_slots + min((int)(n % frame_q), frame_local) : n;
        const Sample<float> s = make_sample<float>(loc[2 * si], loc[2 * si + 1], lv.H[l], lv.W[l], lv.start[l]);
        float w[4];
        if (FILL) {
            const float a = attn[si], hx = 1.f - s.lx, hy = 1.f - s.ly;
            w[0] = hy * hx * a; w[1] = hy * s.lx * a; w[2] = s.ly * hx * a; w[3] = s.ly * s.lx * a;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (s.cell[k] < 0) continue;
            const int64_t dest = (frame * S + s.cell[k]) * M + m;
            if (!FILL) {
                atomicAdd(count + dest, 1);
            } else {
                const int slot = start[dest] + atomicAdd(cursor + dest, 1);
                CornerRec r;
                r.id = (int)(si * 4 + k);
                r.w = w[k];
                records[slot] = r;
            }
        }
    }
}

// ---- three-kernel exclusive scan over `n` ints ----
__global__ void __launch_bounds__(kScanBlock)
scan_tiles_kernel(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ blocksums, int64_t n)
{
    __shared__ int warp_sums[kScanBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    // block-wide exclusive scan of the per-thread sums
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int ws = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        if (lane < kScanBlock / 32) warp_sums[lane] = ws;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - sum + (wid > 0 ? warp_sums[wid - 1] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == kScanBlock - 1) blocksums[blockIdx.x] = run;
}

__global__ void scan_blocksums_kernel(int *__restrict__ blocksums, int nblocks)
{
    // one thread: nblocks is tiny (cells / 2048)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < nblocks; ++i) { const int t = blocksums[i]; blocksums[i] = run; run += t; }
        blocksums[nblocks] = run;
    }
}

__global__ void __launch_bounds__(kScanBlock)
scan_add_kernel(int *__restrict__ out, const int *__restrict__ blocksums, int64_t n, int nblocks)
{
    const int add = blocksums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    for (int i = threadIdx.x; i < kScanTile; i += kScanBlock)
        if (base + i < n) out[base + i] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = blocksums[nblocks];
}

// ---- pass 2: per-cell ordered sum ----
// Every destination cell (n, s, m) owns a list of (corner id, weight) records in arbitrary order.
// The list is put into a canonical order -- ascending corner id, by a rank sort in shared memory --
// and summed in that order with a fixed reduction shape, so the bits of grad_value do not depend
// on how the atomics of pass 1 interleaved.
//
//   det_reduce_kernel       one WARP per cell, lists up to kWarpList records; longer lists are
//                           queued (their cell index appended to `long_cells`)
//   det_reduce_long_kernel  one CTA per queued cell, lists up to kBlockList records in dynamic
//                           shared memory; beyond that a (slow, still deterministic) selection
//                           straight from global memory
//
// Summation shape: the D channels are owned by D/4 float4 lanes (scalar lanes when D % 4 != 0);
// when 2 * D/4 <= 32 two lane groups take the even / odd positions of the ordered list and are
// combined at the end -- a fixed tree, hence reproducible.
struct ListSmem {
    int *id;         // corner id (sort key)
    int *row;        // pair index = id / (4 * LP): row of grad_out
    float *w;
    unsigned short *order;  // order[rank] = position in the unsorted list
};

__device__ __forceinline__ void rank_sort(const ListSmem &l, int cnt, int tid, int nthreads)
{
    for (int i = tid; i < cnt; i += nthreads) {
        const int key = l.id[i];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) rank += l.id[j] < key;  // ids are unique
        l.order[rank] = (unsigned short)i;
    }
}

// Ordered sum of one cell by one warp.  Returns through `dst` (global).
__device__ __forceinline__ void ordered_sum_warp(const ListSmem &l, int cnt, const float *__restrict__ grad_out,
                                                 float *__restrict__ dst, int D, int lane, bool accumulate, bool vec)
{
    const int q4 = D >> 2;
    if (vec && (D & 3) == 0 && q4 <= 16) {  // vec: grad_out and grad_value are 16-byte aligned
        // two interleaved partial sums: lanes [0,q4) take even list positions, [q4,2*q4) odd ones
        const int half = lane / q4, c4 = lane - half * q4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (half < 2) {
            for (int t = half; t < cnt; t += 2) {
                const int i = l.order[t];
                const float4 g = __ldg(reinterpret_cast<const float4 *>(grad_out + (int64_t)l.row[i] * D) + c4);
                fma4(acc, l.w[i], g);
            }
        }
        // lane c4 (first group) adds the second group's partial sum
        const int src = lane + q4 < 32 ? lane + q4 : lane;
        const float ox = __shfl_sync(0xffffffffu, acc.x, src), oy = __shfl_sync(0xffffffffu, acc.y, src);
        const float oz = __shfl_sync(0xffffffffu, acc.z, src), ow = __shfl_sync(0xffffffffu, acc.w, src);
        if (half == 0) {
            float4 r = make_float4(acc.x + ox, acc.y + oy, acc.z + oz, acc.w + ow);
            float4 *d4 = reinterpret_cast<float4 *>(dst) + c4;
            if (accumulate) { const float4 o = *d4; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
            *d4 = r;
        }
    } else {
        for (int c = lane; c < D; c += 32) {
            float acc = 0.f;
            for (int t = 0; t < cnt; ++t) {
                const int i = l.order[t];
                acc = fmaf(l.w[i], __ldg(grad_out + (int64_t)l.row[i] * D + c), acc);
            }
            dst[c] = accumulate ? dst[c] + acc : acc;
        }
    }
}

constexpr int kWarpList = 640;         // records per warp-owned list (14 B each in smem)
constexpr int kReduceWarps = 4;
constexpr int kBlockList = 8192;       // records per CTA-owned list
constexpr int kLongThreads = 256;

__global__ void __launch_bounds__(32 * kReduceWarps)
det_reduce_kernel(const float *__restrict__ grad_out, const int *__restrict__ start,
                  const CornerRec *__restrict__ records, float *__restrict__ grad_value,
                  int *__restrict__ long_cells /* [0] = count, then cell indices */,
                  int D, int LP, int64_t cells, int accumulate, int vec)
{
    __shared__ int s_id[kReduceWarps][kWarpList];
    __shared__ int s_row[kReduceWarps][kWarpList];
    __shared__ float s_w[kReduceWarps][kWarpList];
    __shared__ unsigned short s_order[kReduceWarps][kWarpList];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ListSmem l{s_id[wid], s_row[wid], s_w[wid], s_order[wid]};
    const int lp4 = 4 * LP;
    for (int64_t cell = (int64_t)blockIdx.x * kReduceWarps + wid; cell < cells;
         cell += (int64_t)gridDim.x * kReduceWarps) {
        const int lo = start[cell];
        const int cnt = start[cell + 1] - lo;
        float *dst = grad_value + cell * D;
        if (cnt > kWarpList) {
            if (lane == 0) long_cells[1 + atomicAdd(long_cells, 1)] = (int)cell;
            continue;
        }
        for (int i = lane; i < cnt; i += 32) {
            const CornerRec r = records[lo + i];
            l.id[i] = r.id;
            l.row[i] = r.id / lp4;
            l.w[i] = r.w;
        }
        __syncwarp();
        rank_sort(l, cnt, lane, 32);
        __syncwarp();
        ordered_sum_warp(l, cnt, grad_out, dst, D, lane, accumulate != 0, vec != 0);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kLongThreads)
det_reduce_long_kernel(const float *__restrict__ grad_out, const int *__restrict__ start,
                       const CornerRec *__restrict__ records, float *__restrict__ grad_value,
                       const int *__restrict__ long_cells, int D, int LP, int accumulate, int vec)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    ListSmem l;
    l.id = reinterpret_cast<int *>(dyn);
    l.row = l.id + kBlockList;
    l.w = reinterpret_cast<float *>(l.row + kBlockList);
    l.order = reinterpret_cast<unsigned short *>(l.w + kBlockList);
    const int n_long = long_cells[0];
    const int lp4 = 4 * LP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = blockIdx.x; k < n_long; k += gridDim.x) {
        const int64_t cell = long_cells[1 + k];
        const int lo = start[cell];
        const int cnt = start[cell + 1] - lo;
        float *dst = grad_value + cell * D;
        if (cnt <= kBlockList) {
            for (int i = tid; i < cnt; i += kLongThreads) {
                const CornerRec r = records[lo + i];
                l.id[i] = r.id;
                l.row[i] = r.id / lp4;
                l.w[i] = r.w;
            }
            __syncthreads();
            rank_sort(l, cnt, tid, kLongThreads);
            __syncthreads();
            // the ordered sum itself is one warp's job (fixed reduction shape, same as the short path)
            if (wid == 0) ordered_sum_warp(l, cnt, grad_out, dst, D, lane, accumulate != 0, vec != 0);
            __syncthreads();
        } else {
            // pathological pile-up: selection in id order straight from global memory
            for (int c = tid; c < D; c += kLongThreads) {
                float acc = 0.f;
                int last = -1;
                for (int t = 0; t < cnt; ++t) {
                    int best = 0x7fffffff;
                    float bw = 0.f;
                    for (int j = 0; j < cnt; ++j) {
                        const CornerRec r = records[lo + j];
                        if (r.id > last && r.id < best) { best = r.id; bw = r.w; }
                    }
                    last = best;
                    acc = fmaf(bw, __ldg(grad_out + (int64_t)(best / lp4) * D + c), acc);
                }
                dst[c] = accumulate ? dst[c] + acc : acc;
            }
        }
    }
}

}  // namespace

size_t deterministic_workspace_bytes(const OpDims &d) { return det_layout(d).total; }

// implemented in msda_percall.cu: the regular backward kernels with the grad_value scatter off
cudaError_t launch_backward_no_scatter_f32(const float *value, const int64_t *shapes, const int64_t *lsi,
                                           const float *loc, const float *attn, const float *grad_out,
                                           float *grad_loc, float *grad_attn, const OpDims &d,
                                           cudaStream_t stream);

cudaError_t launch_backward_deterministic_f32(const float *value, const int64_t *shapes,
                                              const int64_t *lsi, const float *loc,
                                              const float *attn, const float *grad_out,
                                              float *grad_value, float *grad_loc, float *grad_attn,
                                              const OpDims &d, void *workspace, cudaStream_t stream,
                                              bool accumulate)
{
    cudaError_t e = launch_backward_no_scatter_f32(value, shapes, lsi, loc, attn, grad_out, grad_loc, grad_attn, d, stream);
    if (e != cudaSuccess) return e;
    return launch_deterministic_grad_value_f32(shapes, lsi, loc, attn, grad_out, grad_value, d, workspace, stream, accumulate);
}

cudaError_t launch_deterministic_grad_value_f32(const int64_t *shapes, const int64_t *lsi, const float *loc,
                                                const float *attn, const float *grad_out, float *grad_value,
                                                const OpDims &d, void *workspace, cudaStream_t stream,
                                                bool accumulate)
{
    const int64_t cells = dest_frames(d) * d.S * d.M;
    const int64_t samples = (int64_t)d.N * d.Lq * d.M * d.L * d.P;
    if (samples * 4 >= (int64_t)INT32_MAX || cells >= (int64_t)INT32_MAX) return cudaErrorInvalidValue;  // 32-bit corner / cell ids
    const DetLayout lay = det_layout(d);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    int *count = reinterpret_cast<int *>(ws + lay.count);
    int *cursor = reinterpret_cast<int *>(ws + lay.cursor);
    int *start = reinterpret_cast<int *>(ws + lay.start);
    int *blocksums = reinterpret_cast<int *>(ws + lay.blocksums);
    CornerRec *records = reinterpret_cast<CornerRec *>(ws + lay.records);
    cudaError_t e;
    if (cells == 0) return cudaSuccess;
    int *long_cells = reinterpret_cast<int *>(ws + lay.long_cells);
    // count, cursor and the long-list queue are adjacent regions: one memset clears all three
    e = cudaMemsetAsync(count, 0, lay.start - lay.count, stream);
    if (e != cudaSuccess) return e;
    const int sample_blocks = (int)((samples + 255) / 256 < 148 * 32 ? (samples + 255) / 256 : 148 * 32);
    if (samples > 0)
        det_count_fill_kernel<false><<<sample_blocks, 256, 0, stream>>>(shapes, lsi, loc, attn, count, cursor, start,
                                                                      records, d.S, d.M, d.L, d.P, d.Lq, samples,
                                                                      d.frame_q, d.frame_local, d.frame_slots);
    const int nblocks = (int)((cells + kScanTile - 1) / kScanTile);
    scan_tiles_kernel<<<nblocks, kScanBlock, 0, stream>>>(count, start, blocksums, cells);
    scan_blocksums_kernel<<<1, 32, 0, stream>>>(blocksums, nblocks);
    scan_add_kernel<<<nblocks, kScanBlock, 0, stream>>>(start, blocksums, cells, nblocks);
    if (samples > 0)
        det_count_fill_kernel<true><<<sample_blocks, 256, 0, stream>>>(shapes, lsi, loc, attn, count, cursor, start,
                                                                     records, d.S, d.M, d.L, d.P, d.Lq, samples,
                                                                      d.frame_q, d.frame_local, d.frame_slots);
    // 128-bit loads / stores of the ordered sum need 16-byte aligned rows (a view with a storage offset, or a C
    // caller's 4-byte aligned buffers, take the scalar branch instead of faulting)
    const int vec = ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_value)) & 15u) == 0 ? 1 : 0;
    const int64_t rblocks = (cells + kReduceWarps - 1) / kReduceWarps;
    det_reduce_kernel<<<(int)(rblocks < 148 * 32 ? rblocks : 148 * 32), 32 * kReduceWarps, 0, stream>>>(
        grad_out, start, records, grad_value, long_cells, d.D, d.L * d.P, cells, accumulate ? 1 : 0, vec);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // over-long lists (none at Snipper's sizes unless many queries pile onto a coarse level)
    const size_t long_smem = (size_t)kBlockList * (3 * sizeof(int) + sizeof(unsigned short));
    e = cudaFuncSetAttribute(det_reduce_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)long_smem);
    if (e != cudaSuccess) return e;
    det_reduce_long_kernel<<<148, kLongThreads, long_smem, stream>>>(grad_out, start, records, grad_value, long_cells,
                                                                    d.D, d.L * d.P, accumulate ? 1 : 0, vec);
    return cudaGetLastError();
}

}  // namespace msda
