/* C-only caller of libmsda_b200.so: proves the boundary is a plain C ABI (no torch, no C++ types).
 * Built and run by tests/test_capi_c_gpu.py:
 *   gcc -std=c11 -I include tests/capi/capi_smoke.c -o capi_smoke -L snipper_b200/lib -lmsda_b200 \
 *       -L oracle/_build -lmsda_oracle -lcudart -lm
 * It runs msda_forward / msda_backward on a small seeded problem and compares with the CPU oracle
 * (oracle/msda_oracle.c, test infrastructure) through the oracle's own C entry points. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "msda_b200.h"

/* oracle/msda_oracle_body.inc */
void msda_oracle_forward_f32(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                             const float *attn, int64_t N, int64_t S, int64_t M, int64_t D, int64_t L,
                             int64_t Lq, int64_t P, float *out);
void msda_oracle_backward_f32(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                              const float *attn, const float *grad_out, int64_t N, int64_t S, int64_t M, int64_t D,
                              int64_t L, int64_t Lq, int64_t P, float *grad_value, float *grad_loc, float *grad_attn);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) { rng_state = rng_state * 1664525u + 1013904223u; return (float)(rng_state >> 8) / 16777216.0f; }

static double rel_err(const float *a, const float *b, size_t n)
{
    double num = 0.0, den = 1e-30;
    for (size_t i = 0; i < n; ++i) {
        const double d = fabs((double)a[i] - (double)b[i]);
        if (d > num) num = d;
        if (fabs((double)b[i]) > den) den = fabs((double)b[i]);
    }
    return num / den;
}

int main(void)
{
    enum { N = 2, M = 8, D = 48, L = 3, P = 4, Lq = 77 };
    const int64_t shapes[L * 2] = {9, 12, 5, 6, 3, 3};
    int64_t lsi[L];
    int64_t S = 0;
    for (int l = 0; l < L; ++l) { lsi[l] = S; S += shapes[2 * l] * shapes[2 * l + 1]; }
    const size_t nv = (size_t)N * S * M * D, ns = (size_t)N * Lq * M * L * P, no = (size_t)N * Lq * M * D;
    float *value = malloc(nv * 4), *loc = malloc(ns * 8), *attn = malloc(ns * 4), *go = malloc(no * 4);
    float *out = malloc(no * 4), *gv = malloc(nv * 4), *gl = malloc(ns * 8), *ga = malloc(ns * 4);
    float *r_out = malloc(no * 4), *r_gv = calloc(nv, 4), *r_gl = malloc(ns * 8), *r_ga = malloc(ns * 4);
    for (size_t i = 0; i < nv; ++i) value[i] = frand() * 2.f - 1.f;
    for (size_t i = 0; i < 2 * ns; ++i) loc[i] = frand() * 1.2f - 0.1f;   /* some samples outside [0,1] */
    for (size_t i = 0; i < ns; ++i) attn[i] = frand() / (L * P);
    for (size_t i = 0; i < no; ++i) go[i] = frand() * 2.f - 1.f;

    if (msda_abi_version() != MSDA_ABI_VERSION) { printf("ABI mismatch\n"); return 1; }
    void *d_value, *d_loc, *d_attn, *d_go, *d_out, *d_gv, *d_gl, *d_ga, *d_shapes, *d_lsi;
    CK(cudaMalloc(&d_value, nv * 4)); CK(cudaMalloc(&d_loc, ns * 8)); CK(cudaMalloc(&d_attn, ns * 4));
    CK(cudaMalloc(&d_go, no * 4)); CK(cudaMalloc(&d_out, no * 4)); CK(cudaMalloc(&d_gv, nv * 4));
    CK(cudaMalloc(&d_gl, ns * 8)); CK(cudaMalloc(&d_ga, ns * 4));
    CK(cudaMalloc(&d_shapes, sizeof shapes)); CK(cudaMalloc(&d_lsi, sizeof lsi));
    CK(cudaMemcpy(d_value, value, nv * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_loc, loc, ns * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_attn, attn, ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_go, go, no * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_shapes, shapes, sizeof shapes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_lsi, lsi, sizeof lsi, cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    int st = msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, N, (int)S, M, D, L, Lq, P, 0, 64,
                          MSDA_DTYPE_F32, stream);
    if (st != MSDA_OK) { printf("msda_forward: %s\n", msda_error_string(st)); return 1; }
    st = msda_backward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_go, d_gv, d_gl, d_ga, N, (int)S, M, D, L, Lq, P, 0,
                       64, MSDA_DTYPE_F32, 0u, NULL, 0, stream);
    if (st != MSDA_OK) { printf("msda_backward: %s\n", msda_error_string(st)); return 1; }
    /* the reference's error contract through the C ABI: batch 3 does not divide im2col_step 2 */
    if (msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, 3, (int)S, M, D, L, Lq, P, 0, 2, MSDA_DTYPE_F32,
                     stream) != MSDA_ERR_IM2COL_STEP) { printf("im2col_step not validated\n"); return 1; }
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemcpy(out, d_out, no * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gv, d_gv, nv * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gl, d_gl, ns * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ga, d_ga, ns * 4, cudaMemcpyDeviceToHost));

    msda_oracle_forward_f32(value, shapes, lsi, loc, attn, N, S, M, D, L, Lq, P, r_out);
    msda_oracle_backward_f32(value, shapes, lsi, loc, attn, go, N, S, M, D, L, Lq, P, r_gv, r_gl, r_ga);
    const double e0 = rel_err(out, r_out, no), e1 = rel_err(gv, r_gv, nv), e2 = rel_err(gl, r_gl, 2 * ns),
                 e3 = rel_err(ga, r_ga, ns);
    printf("capi_smoke: out %.2e grad_value %.2e grad_loc %.2e grad_attn %.2e\n", e0, e1, e2, e3);
    if (!(e0 < 1e-5 && e1 < 1e-4 && e2 < 1e-4 && e3 < 1e-4)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
