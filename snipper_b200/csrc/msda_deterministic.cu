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
constexpr int kReduceLanes = 16;        // lanes per destination cell
constexpr int kReduceCellsPerBlock = 8; // 128 threads
constexpr int kMaxListInSmem = 192;     // records per cell staged in shared memory

struct __align__(8) CornerRec {
    int id;    // (sample index) * 4 + corner
    float w;   // bilinear weight * attention weight
};

struct DetLayout {
    size_t count, cursor, start, blocksums, records, total;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

DetLayout det_layout(const OpDims &d)
{
    const size_t cells = (size_t)d.N * d.S * d.M;
    const size_t corners = (size_t)d.N * d.Lq * d.M * d.L * d.P * 4;
    const size_t nblocks = (cells + kScanTile - 1) / kScanTile;
    DetLayout l;
    size_t off = 0;
    l.count = off; off = align256(off + sizeof(int) * cells);
    l.cursor = off; off = align256(off + sizeof(int) * cells);
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
                      CornerRec *__restrict__ records, int S, int M, int L, int P, int Lq, int64_t total_samples)
{
    __shared__ LevelTable lv;
    load_level_table(lv, shapes, lsi, L);
    __syncthreads();
    const int LP = L * P;
    for (int64_t si = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; si < total_samples;
         si += (int64_t)gridDim.x * blockDim.x) {
        const int lp = (int)(si % LP);
        const int64_t pair = si / LP;
        const int m = (int)(pair % M);
        const int64_t n = pair / ((int64_t)Lq * M);
        const int l = lp / P;
        const Sample<float> s = make_sample<float>(loc[2 * si], loc[2 * si + 1], lv.H[l], lv.W[l], lv.start[l]);
        float w[4];
        if (FILL) {
            const float a = attn[si], hx = 1.f - s.lx, hy = 1.f - s.ly;
            w[0] = hy * hx * a; w[1] = hy * s.lx * a; w[2] = s.ly * hx * a; w[3] = s.ly * s.lx * a;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (s.cell[k] < 0) continue;
            const int64_t dest = (n * S + s.cell[k]) * M + m;
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
__global__ void __launch_bounds__(kReduceLanes * kReduceCellsPerBlock)
det_reduce_kernel(const float *__restrict__ grad_out, const int *__restrict__ start,
                  const CornerRec *__restrict__ records, float *__restrict__ grad_value,
                  int D, int LP, int64_t cells, int accumulate)
{
    __shared__ CornerRec raw[kReduceCellsPerBlock][kMaxListInSmem];
    __shared__ CornerRec sorted[kReduceCellsPerBlock][kMaxListInSmem];
    const int g = threadIdx.x / kReduceLanes;
    const int lane = threadIdx.x % kReduceLanes;
    const unsigned gmask = 0xffffu << ((threadIdx.x & 31) / kReduceLanes * kReduceLanes);
    for (int64_t cell0 = (int64_t)blockIdx.x * kReduceCellsPerBlock; cell0 < cells;
         cell0 += (int64_t)gridDim.x * kReduceCellsPerBlock) {
        const int64_t cell = cell0 + g;
        const bool live = cell < cells;
        const int lo = live ? start[cell] : 0;
        const int cnt = live ? start[cell + 1] - lo : 0;
        float *dst = grad_value + cell * D;
        if (cnt <= kMaxListInSmem) {
            for (int i = lane; i < cnt; i += kReduceLanes) raw[g][i] = records[lo + i];
            __syncwarp(gmask);
            // rank sort by id (ids are unique): canonical order independent of the fill order
            for (int i = lane; i < cnt; i += kReduceLanes) {
                const CornerRec r = raw[g][i];
                int rank = 0;
                for (int j = 0; j < cnt; ++j) rank += raw[g][j].id < r.id;
                sorted[g][rank] = r;
            }
            __syncwarp(gmask);
            for (int c = lane; c < D; c += kReduceLanes) {
                float acc = 0.f;
                for (int t = 0; t < cnt; ++t) {
                    const CornerRec r = sorted[g][t];
                    const int64_t pair = (int64_t)(r.id >> 2) / LP;
                    acc = fmaf(r.w, __ldg(grad_out + pair * D + c), acc);
                }
                if (live) dst[c] = accumulate ? dst[c] + acc : acc;
            }
            __syncwarp(gmask);
        } else {
            // very long list (pathological pile-up on one cell): selection in id order from global
            for (int c = lane; c < D; c += kReduceLanes) {
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
                    const int64_t pair = (int64_t)(best >> 2) / LP;
                    acc = fmaf(bw, __ldg(grad_out + pair * D + c), acc);
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
    const int64_t cells = (int64_t)d.N * d.S * d.M;
    const int64_t samples = (int64_t)d.N * d.Lq * d.M * d.L * d.P;
    if (samples * 4 >= (int64_t)INT32_MAX) return cudaErrorInvalidValue;  // corner ids are 32-bit
    const DetLayout lay = det_layout(d);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    int *count = reinterpret_cast<int *>(ws + lay.count);
    int *cursor = reinterpret_cast<int *>(ws + lay.cursor);
    int *start = reinterpret_cast<int *>(ws + lay.start);
    int *blocksums = reinterpret_cast<int *>(ws + lay.blocksums);
    CornerRec *records = reinterpret_cast<CornerRec *>(ws + lay.records);

    cudaError_t e = launch_backward_no_scatter_f32(value, shapes, lsi, loc, attn, grad_out, grad_loc, grad_attn, d, stream);
    if (e != cudaSuccess) return e;
    if (cells == 0) return cudaSuccess;
    // count and cursor are adjacent regions: one memset clears both
    e = cudaMemsetAsync(count, 0, lay.start - lay.count, stream);
    if (e != cudaSuccess) return e;
    const int sample_blocks = (int)((samples + 255) / 256 < 148 * 32 ? (samples + 255) / 256 : 148 * 32);
    if (samples > 0)
        det_count_fill_kernel<false><<<sample_blocks, 256, 0, stream>>>(shapes, lsi, loc, attn, count, cursor, start,
                                                                      records, d.S, d.M, d.L, d.P, d.Lq, samples);
    const int nblocks = (int)((cells + kScanTile - 1) / kScanTile);
    scan_tiles_kernel<<<nblocks, kScanBlock, 0, stream>>>(count, start, blocksums, cells);
    scan_blocksums_kernel<<<1, 32, 0, stream>>>(blocksums, nblocks);
    scan_add_kernel<<<nblocks, kScanBlock, 0, stream>>>(start, blocksums, cells, nblocks);
    if (samples > 0)
        det_count_fill_kernel<true><<<sample_blocks, 256, 0, stream>>>(shapes, lsi, loc, attn, count, cursor, start,
                                                                     records, d.S, d.M, d.L, d.P, d.Lq, samples);
    const int64_t rblocks = (cells + kReduceCellsPerBlock - 1) / kReduceCellsPerBlock;
    det_reduce_kernel<<<(int)(rblocks < 148 * 64 ? rblocks : 148 * 64), kReduceLanes * kReduceCellsPerBlock, 0, stream>>>(
        grad_out, start, records, grad_value, d.D, d.L * d.P, cells, accumulate ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace msda
