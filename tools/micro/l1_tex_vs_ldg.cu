// Micro-benchmark: is the texture path of L1TEX an ADDITIONAL source of gather bandwidth next to the
// LSU path?  Each lane group of 12 lanes gathers 192-byte "head slices" (16 B per lane) at
// pseudo-random cells of a small, L1-resident buffer -- the access shape of the MSDA forward -- through
// LDG.128, through tex1Dfetch<float4>, or through a mix.   nvcc -arch=sm_100a -O3 -o l1tex l1_tex_vs_ldg.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)


constexpr int kIters = 512;

template <int MODE>  // 0 = LDG only, 1 = TEX only, 2 = alternate LDG/TEX, 3 = 3 LDG : 1 TEX
__global__ void __launch_bounds__(192) gather(const float4 *__restrict__ buf, cudaTextureObject_t tex, float4 *out, int cells)
{
    const int lane = threadIdx.x % 12, pair = threadIdx.x / 12, m = blockIdx.x & 7;
    unsigned s = (blockIdx.x * 16 + pair) * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        s = s * 1664525u + 1013904223u;
        const int cell = (s >> 8) % cells;
        const int idx = cell * 96 + m * 12 + lane;   // float4 index: cell stride 1536 B, head slice 192 B
        float4 v;
        bool use_tex = MODE == 1 || (MODE == 2 && (i & 1)) || (MODE == 3 && (i & 3) == 3);
        if (use_tex) v = tex1Dfetch<float4>(tex, idx);
        else v = __ldg(buf + idx);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(const float4 *buf, cudaTextureObject_t tex, float4 *out, int cells, int grid)
{
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    for (int w = 0; w < 3; ++w) gather<MODE><<<grid, 192>>>(buf, tex, out, cells);
    CHECK(cudaEventRecord(a));
    for (int r = 0; r < 10; ++r) gather<MODE><<<grid, 192>>>(buf, tex, out, cells);
    CHECK(cudaEventRecord(b));
    CHECK(cudaEventSynchronize(b));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    return ms / 10;
}

int main()
{
    const int grid = 148 * 8 * 4;
    for (int cells : {64, 256, 4096}) {
        const size_t n4 = (size_t)cells * 96;
        float4 *buf, *out;
        CHECK(cudaMalloc(&buf, n4 * sizeof(float4)));
        CHECK(cudaMemset(buf, 0, n4 * sizeof(float4)));
        CHECK(cudaMalloc(&out, (size_t)grid * 192 * sizeof(float4)));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = buf;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = n4 * sizeof(float4);
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex;
        CHECK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        const double bytes = (double)grid * 192 * kIters * 16;
        const float t0 = run<0>(buf, tex, out, cells, grid), t1 = run<1>(buf, tex, out, cells, grid);
        const float t2 = run<2>(buf, tex, out, cells, grid), t3 = run<3>(buf, tex, out, cells, grid);
        printf("{\"cells\": %d, \"buffer_KB\": %zu, \"ldg_ms\": %.4f, \"tex_ms\": %.4f, \"alt_ms\": %.4f, \"ldg3_tex1_ms\": %.4f, "
               "\"ldg_TBps\": %.2f, \"tex_TBps\": %.2f, \"alt_TBps\": %.2f, \"ldg3_tex1_TBps\": %.2f}\n",
               cells, n4 * 16 / 1024, t0, t1, t2, t3, bytes / t0 / 1e9, bytes / t1 / 1e9, bytes / t2 / 1e9, bytes / t3 / 1e9);
        CHECK(cudaDestroyTextureObject(tex));
        CHECK(cudaFree(buf)); CHECK(cudaFree(out));
    }
    return 0;
}
