// Shared device/host helpers for libsvae_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/svae_b200.h"

namespace svae {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SVAE_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return svae::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
// every kernel launch of the library goes through this macro: it also feeds svae_launch_count()
void count_launch();
#define SVAE_LAUNCH_CHECK()                                                    \
    do {                                                                       \
        svae::count_launch();                                                  \
        SVAE_CUDA(cudaPeekAtLastError());                                      \
    } while (0)
#define SVAE_TRY(call)                                                         \
    do {                                                                       \
        int r__ = (call);                                                      \
        if (r__ != SVAE_OK) return r__;                                        \
    } while (0)
#define SVAE_REQUIRE(cond, code, ...)                                          \
    do {                                                                       \
        if (!(cond)) { svae::set_error(__VA_ARGS__); return (code); }          \
    } while (0)

__host__ __device__ static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
__host__ __device__ static inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

// ---- activations ----------------------------------------------------------------------------
// Forward value and derivative expressed through the OUTPUT h = act(a), which is what the
// workspace keeps (tanh' = 1-h^2, sigmoid' = h(1-h); for (leaky)relu sign(h) == sign(a)).
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool FAST>
__device__ __forceinline__ float act_apply(int act, float a) {
    switch (act) {
        case SVAE_ACT_TANH:      return FAST ? tanh_fast(a) : tanhf(a);
        case SVAE_ACT_LEAKYRELU: return a > 0.f ? a : 0.01f * a;
        case SVAE_ACT_RELU:      return a > 0.f ? a : 0.f;
        default:                 return 1.f / (1.f + expf(-a));
    }
}

__device__ __forceinline__ float act_deriv_from_out(int act, float h) {
    switch (act) {
        case SVAE_ACT_TANH:      return 1.f - h * h;
        case SVAE_ACT_LEAKYRELU: return h > 0.f ? 1.f : 0.01f;
        case SVAE_ACT_RELU:      return h > 0.f ? 1.f : 0.f;
        default:                 return h * (1.f - h);
    }
}

// ---- storage type of the (rows x Hp) activation matrices: float (parity) or bf16 (fast) -------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- fp32 SIMT GEMM (sgemm.cu) ---------------------------------------------------------------
// C[m,n] (op)= alpha * sum_k A(m,k) * B(k,n)  with arbitrary element strides, then the epilogue
//   v += bias[n];  v += add[m*ld_add + n];  v = act(v);  v *= act'(dsrc[m*ld_dsrc + n])
// (add: the skip connection of a ResidLinear layer, models.py:13-21 -- the layer input in the forward,
//  the incoming gradient in the backward)
// accumulate = 1 adds into C with atomics (required when split_k > 1).
struct SgemmArgs {
    const float* A; long sAm, sAk;
    const float* B; long sBk, sBn;
    float* C; long ldc;
    int M, N, K;
    const float* bias = nullptr;
    int act = -1;
    const float* dsrc = nullptr; long ld_dsrc = 0; int dact = -1;
    const float* add = nullptr; long ld_add = 0;
    int accumulate = 0;
    int split_k = 1;
    float alpha = 1.f;
};
int sgemm(const SgemmArgs& a, cudaStream_t st);

// ---- bf16 tcgen05 GEMMs (tc_gemm.cu) -----------------------------------------------------------
struct TcExtra {
    // mode 0: fused output-layer dot product  o_accum[m,c] += sum_n h[m,n] out_w[c,n]
    const float* out_w = nullptr; int out_w_ld = 0; int dot_c = 0; float* o_accum = nullptr;
    // modes 0/1: fp32 output (and fp32 aux) instead of bf16
    int out_f32 = 0;
    // mode 0 with out_f32: split K over the CTA pairs that would otherwise idle and ADD the raw partial sums to `out`
    // with fp32 atomics (no bias, no activation; zero `out` first and finish the layer with split3_act)
    int raw_split_k = 0;
    // mode 0: ResidLinear (models.py:13-21): out = act(A W^T + bias + resid), resid (M x N, bf16, ld elements)
    const void* resid = nullptr; int ld_resid = 0;
};
int tc_split_k_factor(int M, int N);      // CTA pairs (or CTAs) per output tile of an M x N forward GEMM
int tc_gemm(int mode, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias,
            int bias_n, const void* aux, int ldaux, int act, void* out, int ldo, cudaStream_t st,
            const TcExtra& ex = TcExtra());


// ---- fused backward kernels (tc_bwd.cu) ----------------------------------------------------------------
// First-layer context of the transposed dX GEMM that reduces delta_0 per image instead of storing it:
// h_0[row, n] = act(coord_w[n,0] x' + coord_w[n,1] y' + hz[b, n]) is recomputed from (grid, img);
// S[b0 + row / P, {1, c0, c1}, n] += delta_0[row, n] {1, grid[row % P]}   (S zeroed by the caller).
struct TcMoments {
    const float* grid = nullptr;      // (P, 2)
    const float* img = nullptr;       // (B, 4) cos, sin, dx0, dx1
    const float* coord_w = nullptr;   // (H, 2)
    const float* hz = nullptr;        // (B, Hp)
    float* S = nullptr;               // (B, 3, Hp)
    int P = 0, b0 = 0;
};
// delta (rows x Hp, bf16), W (Hp x Hp bf16, [j][n]): S += moments of (delta W) .* act'(h_0)
int tc_dx_moments(int rows, int H, int Hp, const void* delta, int ldd, const void* W, int ldw, int act,
                  const TcMoments& r, cudaStream_t st);


// Top hidden layer (tc_bwd.cu): dW (H x H fp32, ld ldo) += delta^T h_prev with
// delta[row, m] = (sum_c g_o[row, c] out_w[c, m]) * act'(h_top[row, m]) built in shared memory from h_top; also
// d_out_w (C, H) += g_o^T h_top, d_out_b (C) += colsum g_o, d_b (H) += colsum delta (d_b may be NULL), and delta is
// written to delta_out (rows x Hp bf16) unless that is NULL.  h_top, h_prev: rows x Hp bf16, ld Hp.
struct TcTop {
    const float* g_o = nullptr; int C = 0;
    const float* out_w = nullptr;
    float* d_out_w = nullptr; float* d_out_b = nullptr; float* d_b = nullptr;
};
int tc_dw_top(int rows, int H, int Hp, const void* h_top, const void* h_prev, int act, const TcTop& t, float* dW, int ldo,
              void* delta_out, cudaStream_t st);

}  // namespace svae
