// Micro-benchmark of the ACCESS SHAPES the fused forward could use ("test the access-shape hypotheses"):
// pure gathers of 16 bytes per lane from an L1/L2-resident buffer, nothing else, reported as TB/s of gathered bytes.
//
//   A  baseline       12 lanes x LDG.128 on a 192-byte head slice of a cell-major value tensor (cell stride 1536 B),
//                     one broadcast LDS.128 of the sample record per gather group of 4 corners (what the kernel does)
//   B  shfl records   same gathers, the record comes from 4 SHFL instead of the LDS.128
//   C  x-pairs        head-major plane (cell stride 192 B): 24 lanes load the 384 contiguous bytes of two x-adjacent
//                     corners, lanes rotated so that every quarter warp stays inside one 128-byte line where it can
//                     (start at 0 or 64 mod 128 at random, as real samples would)
//   D  x-pairs, always 128-byte aligned (what a second, one-cell-shifted copy of the plane would buy)
//   E  shared window  12 lanes x LDS.128 on 192-byte slices of a 256-cell window staged in shared memory
//                     (the gather part of a TMA-staged window design; staging cost not included)
//
//   F  split planes   the layout a buffer of OUR OWN (the neighbour-frame sums) can take: per head a 128-byte plane
//                     (channels 0-31, one line per cell) and a 64-byte plane (channels 32-47) whose x-adjacent cells share a
//                     line (two copies, one shifted by a cell, so the pair is always line-aligned).  8 lanes = one quarter
//                     warp per query: 4 x LDG.128 on plane A + 2 x LDG.128 on plane B per sample, every wavefront a full line
//                     (6 per sample instead of ~8.4), record by one broadcast LDS.128 per quarter warp
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_shapes gather_shapes.cu && ./gather_shapes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kIters = 512;       // gather groups per lane; each group = 4 loads of 16 B
constexpr int kThreads = 192;

__device__ __forceinline__ unsigned lcg(unsigned s) { return s * 1664525u + 1013904223u; }
__device__ __forceinline__ void acc4(float4 &a, const float4 &v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

// MODE 0 = A (LDS.128 record), 1 = B (4 SHFL), 2 = no record at all (pure gather, the r01 ceiling)
template <int MODE>
__global__ void __launch_bounds__(kThreads) slices(const float4 *__restrict__ buf, float4 *out, int cells)
{
    __shared__ float4 rec[16 * 13];
    for (int i = threadIdx.x; i < 16 * 13; i += kThreads) rec[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x % 12, pair = threadIdx.x / 12, m = blockIdx.x & 7;
    unsigned s = (blockIdx.x * 16 + pair) * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 mine = rec[pair * 13 + lane];
#pragma unroll 2
    for (int i = 0; i < kIters; ++i) {
        s = lcg(s);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0) r = rec[pair * 13 + (i % 12)];
        if (MODE == 1) {
            const int src = (threadIdx.x & 31) & ~3;
            r.x = __shfl_sync(0xffffffffu, mine.x, src); r.y = __shfl_sync(0xffffffffu, mine.y, src);
            r.z = __shfl_sync(0xffffffffu, mine.z, src); r.w = __shfl_sync(0xffffffffu, mine.w, src);
        }
        const int cell = ((s >> 8) + __float_as_int(r.w)) % (cells - 101);
        const float4 *p = buf + (size_t)cell * 96 + m * 12 + lane;     // cell stride 1536 B, head slice 192 B
        float4 v0 = __ldg(p), v1 = __ldg(p + 96), v2 = __ldg(p + 100 * 96), v3 = __ldg(p + 101 * 96);
        acc4(acc, v0); acc4(acc, v1); acc4(acc, v2); acc4(acc, v3);
        acc.x += r.x + r.y + r.z;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// x-pairs on a head-major plane: 24 lanes per row pair, two row pairs (y0, y0+1) per sample
template <bool ALIGNED>
__global__ void __launch_bounds__(kThreads) xpairs(const float4 *__restrict__ plane, float4 *out, int cells)
{
    const int j = threadIdx.x % 24, grp = threadIdx.x / 24;
    unsigned s = (blockIdx.x * 8 + grp) * 2654435761u + 777u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int i = 0; i < kIters; ++i) {
        s = lcg(s);
        int cell = (s >> 8) % (cells - 102);
        if (ALIGNED) cell &= ~1;                 // 384-byte segments start on a line boundary
        const bool odd = cell & 1;              // segment starts at 64 mod 128
        // lane -> 16-byte piece of the 384-byte segment, grouped so a quarter warp touches one line where possible
        int piece = j;
        if (odd) piece = j < 16 ? j + 4 : (j < 20 ? j - 16 : j);
        const float4 *p = plane + (size_t)cell * 12 + piece;             // cell stride 192 B
        float4 v0 = __ldg(p), v1 = __ldg(p + 100 * 12);                   // rows y0 and y0 + 1 (W = 100)
        acc4(acc, v0); acc4(acc, v1);
        s = lcg(s);
        cell = (s >> 8) % (cells - 102);
        if (ALIGNED) cell &= ~1;
        piece = j;
        if (cell & 1) piece = j < 16 ? j + 4 : (j < 20 ? j - 16 : j);
        p = plane + (size_t)cell * 12 + piece;
        v0 = __ldg(p); v1 = __ldg(p + 100 * 12);
        acc4(acc, v0); acc4(acc, v1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}


// F: split planes, one quarter warp per query (see header)
__global__ void __launch_bounds__(kThreads) split_planes(const float4 *__restrict__ planeA, const float4 *__restrict__ planeB,
                                                         float4 *out, int cells)
{
    __shared__ float4 rec[24 * 13];
    for (int i = threadIdx.x; i < 24 * 13; i += kThreads) rec[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 7, pair = threadIdx.x >> 3, m = blockIdx.x & 7;
    unsigned s = (blockIdx.x * 24 + pair) * 2654435761u + 991u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), accb = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *pa = planeA + (size_t)m * cells * 8 + lane;          // head plane, 128-byte cells
    const float4 *pb = planeB + (size_t)m * cells * 4 + lane;          // head plane, 64-byte cells, lanes 0-3 | 4-7 = x0 | x0+1
#pragma unroll 2
    for (int i = 0; i < kIters; ++i) {
        s = lcg(s);
        const float4 r = rec[pair * 13 + (i % 12)];
        const int cell = ((s >> 8) + __float_as_int(r.w)) % (cells - 102);
        const float4 *a = pa + (size_t)cell * 8;
        float4 v0 = __ldg(a), v1 = __ldg(a + 8), v2 = __ldg(a + 100 * 8), v3 = __ldg(a + 101 * 8);
        const float4 *b = pb + (size_t)(cell & ~1) * 4;                // pair always line-aligned (the shifted copy serves odd x0)
        float4 u0 = __ldg(b), u1 = __ldg(b + 100 * 4);
        acc4(acc, v0); acc4(acc, v1); acc4(acc, v2); acc4(acc, v3);
        acc4(accb, u0); acc4(accb, u1);
        acc.x += r.x + r.y + r.z;
    }
    acc4(acc, accb);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

constexpr int kWindowCells = 256;
__global__ void __launch_bounds__(kThreads) shared_window(const float4 *__restrict__ buf, float4 *out)
{
    extern __shared__ float4 win[];             // kWindowCells x 12 float4 (192-byte cells, head-major)
    for (int i = threadIdx.x; i < kWindowCells * 12; i += kThreads) win[i] = buf[i];
    __syncthreads();
    const int lane = threadIdx.x % 12, pair = threadIdx.x / 12;
    unsigned s = (blockIdx.x * 16 + pair) * 2654435761u + 4242u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int i = 0; i < kIters; ++i) {
        s = lcg(s);
        const int cell = (s >> 8) % (kWindowCells - 18);
        const float4 *p = win + cell * 12 + lane;
        acc4(acc, p[0]); acc4(acc, p[12]); acc4(acc, p[16 * 12]); acc4(acc, p[17 * 12]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float timeit(F launch)
{
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    for (int w = 0; w < 3; ++w) launch();
    CHECK(cudaEventRecord(a));
    for (int r = 0; r < 10; ++r) launch();
    CHECK(cudaEventRecord(b));
    CHECK(cudaEventSynchronize(b));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    CHECK(cudaGetLastError());
    return ms / 10;
}

int main()
{
    const int grid = 148 * 8 * 4;
    const int cells = 9875;                      // one frame of the 600x800 pyramid: 15.2 MB cell-major, 1.9 MB per head plane
    float4 *buf, *out;
    CHECK(cudaMalloc(&buf, (size_t)cells * 96 * sizeof(float4)));
    CHECK(cudaMemset(buf, 0, (size_t)cells * 96 * sizeof(float4)));
    CHECK(cudaMalloc(&out, (size_t)grid * kThreads * sizeof(float4)));
    const double bytes = (double)grid * kThreads * kIters * 4 * 16;   // every variant gathers 4 x 16 B per lane and group
    const size_t win_bytes = (size_t)kWindowCells * 12 * sizeof(float4);
    CHECK(cudaFuncSetAttribute(shared_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_bytes));
    const float tA = timeit([&] { slices<0><<<grid, kThreads>>>(buf, out, cells); });
    const float tB = timeit([&] { slices<1><<<grid, kThreads>>>(buf, out, cells); });
    const float tP = timeit([&] { slices<2><<<grid, kThreads>>>(buf, out, cells); });
    const float tC = timeit([&] { xpairs<false><<<grid, kThreads>>>(buf, out, cells); });
    const float tD = timeit([&] { xpairs<true><<<grid, kThreads>>>(buf, out, cells); });
    const float tF = timeit([&] { split_planes<<<grid, kThreads>>>(buf, buf + (size_t)cells * 64, out, cells); });
    const double bytesF = (double)grid * kThreads * kIters * 6 * 16;   // 8 lanes x 6 loads = the same 768 B per sample
    const float tE = timeit([&] { shared_window<<<grid, kThreads, win_bytes>>>(buf, out); });
    printf("{\"what\": \"TB/s of gathered bytes, 16 B per lane per load, %d CTAs x %d threads x %d groups of 4 loads\", "
           "\"A_slices_lds_record\": %.2f, \"B_slices_shfl_record\": %.2f, \"slices_no_record\": %.2f, "
           "\"C_xpairs_head_major\": %.2f, \"D_xpairs_always_aligned\": %.2f, \"E_shared_window_slices\": %.2f, "
           "\"F_split_planes_quarter_warp_per_query\": %.2f, \"F_samples_per_us\": %.1f, \"A_samples_per_us\": %.1f, "
           "\"ms\": {\"A\": %.4f, \"B\": %.4f, \"none\": %.4f, \"C\": %.4f, \"D\": %.4f, \"E\": %.4f, \"F\": %.4f}}\n",
           grid, kThreads, kIters, bytes / tA / 1e9, bytes / tB / 1e9, bytes / tP / 1e9, bytes / tC / 1e9, bytes / tD / 1e9,
           bytes / tE / 1e9, bytesF / tF / 1e9, bytesF / 768 / tF / 1e3, bytes / 768 / tA / 1e3, tA, tB, tP, tC, tD, tE, tF);
    return 0;
}
