// Launch wrappers of the SIMT kernels of the step (step_kernels.cu).  T is the storage type of
// the (rows x Hp) activation matrices: float (PARITY) or __nv_bfloat16 (FAST).
#pragma once
#include "common.cuh"

namespace svae {

// Reparameterisation + KL (train_mnist.py:33-39,62-63,84-86).  One thread per image.
//   zo (B,2I) = [mu | logstd]; writes lat (B,I), img (B,4) = cos,sin,dx0,dx1, zs (B,Z) scaled z,
//   stats[b*3+1] = kl_b.
// eps == NULL: eps is drawn in the kernel (Philox4x32-10 keyed by rng.seed, counter = (global image, latent block,
// *rng.step), Box-Muller) and written to eps_out (B,I) for the backward.
struct LatentRng { const int32_t* step = nullptr; uint64_t seed = 0; int64_t image_offset = 0; };
int latent_forward(const SvaeShape& s, const SvaeConfig& c, const float* zo, const float* eps,
                   const float* theta_offset, float* lat, float* img, float* zs, float* stats, cudaStream_t st,
                   const LatentRng& rng = LatentRng(), float* eps_out = nullptr);

// hz[b, n] = coord_b[n] for all b when Z == 0 (otherwise an sgemm with bias does it)
int fill_rows(float* dst, const float* row, int rows, int n, int ld, cudaStream_t st);

// First layer (models.py:104-124 + layers[0]): h0[r,n] = act(Wc[n,0]*x0' + Wc[n,1]*x1' + hz[b,n]).
// Coordinates come either from (grid, img) -> R(theta)*grid + dx, or from explicit x (B,P,2).
template <typename T>
int layer0_forward(const SvaeShape& s, int act, int b0, int nb, const float* coord_w, const float* hz,
                   const float* grid, const float* img, const float* x_explicit, int H, int Hp, T* h0,
                   cudaStream_t st);

// Output layer + sigmoid (+softplus ch0) (models.py:84-85,129-130): warp-shuffle dot per row.
// Writes logits o (rows,C) and, if non-NULL, y_hat (rows,C).
template <typename T>
int out_forward(const T* h, int rows, int H, int Hp, int C, const float* out_w, const float* out_b, int softplus,
                float* o, float* y_hat, cudaStream_t st);

// y_hat = sigmoid(o) (+softplus on channel 0) from stored logits (used when the dot product was fused)
int logits_to_yhat(const float* o, float* y_hat, long n, int C, int softplus, cudaStream_t st);

// Likelihood per image (train_mnist.py:80-81, train_particles.py:102-139, train_galaxy.py:118-119):
// reads logits o (nb,P,C), targets, optional CTF kernels and mask; writes stats[b*3+0] = logp_b and
// g_o (nb,P,C) = grad_scale * d(-logp_b)/do.
int likelihood(const SvaeShape& s, const SvaeConfig& c, int b0, int nb, const float* o, const float* y,
               const float* ctf, const uint8_t* mask, float* stats, float* g_o, cudaStream_t st);

// Output-layer backward: delta[r,n] = (sum_c g_o[r,c] Wo[c,n]) * act'(h[r,n]); accumulates
// dWo += g_o^T h, dbo += colsum(g_o), db_last += colsum(delta) (db_last may be NULL).
template <typename T>
int out_backward(const T* h, const float* g_o, int rows, int H, int Hp, int C, int act, const float* out_w,
                 T* delta, float* d_out_w, float* d_out_b, float* d_b_last, cudaStream_t st);

// colsum: dst[n] += sum_r src[r,n]   (hidden-layer bias gradients)
template <typename T>
int col_sum(const T* src, int rows, int H, int Hp, float* dst, cudaStream_t st);

// Per-image column reductions of delta0: S[b,0,n] = sum_p d, S[b,1,n] = sum_p c0 d,
// S[b,2,n] = sum_p c1 d with (c0,c1) the UNtransformed grid coordinate (or explicit x[b,p,:]).
template <typename T>
int image_col_reduce(const T* delta0, int b0, int nb, int P, int Hp, const float* grid, const float* x_explicit,
                     float* S, cudaStream_t st);

// Row gradients w.r.t. explicit coordinates: g_x[r,k] = sum_n delta0[r,n] Wc[n,k]
template <typename T>
int coord_row_grad(const T* delta0, int rows, int H, int Hp, const float* coord_w, float* g_x, cudaStream_t st);

// dWc, dbc from S and the per-image transform; explicit != 0: S already holds x-moments.
int coord_param_grad(const float* S, const float* img, int B, int H, int Hp, int explicit_x, float* d_coord_w,
                     float* d_coord_b, cudaStream_t st);

// Per-image latent gradient -> encoder head gradient g_zo (B,2I) (SURVEY 7.3 last line).
// coord_pre: optional (B,3) = (d theta, d t0, d t1) already computed by latent_coord_grad (option path); when set,
// S and coord_w are not read.
int latent_backward(const SvaeShape& s, const SvaeConfig& c, const float* S, int Hp, const float* img,
                    const float* coord_w, const float* dz, const float* zo, const float* eps, float* g_zo,
                    cudaStream_t st, const float* coord_pre = nullptr);

// ---- decoder options (option_kernels.cu): --expand-coords (F = 5 coordinate features) and --bilinear (per-image
// coordinate weights w[b*w_img_stride + n*F + i]; stride 0 = coord_linear.weight shared by all images).
// Tm (B, F+1, Hp) are the feature moments of delta0 (see first_layer.cuh).
template <typename T>
int layer0_opt_forward(int F, int P, int act, int b0, int nb, const float* w, long w_img_stride, const float* hz,
                       const float* grid, const float* img, const float* x_explicit, int H, int Hp, T* h0,
                       cudaStream_t st);
template <typename T>
int image_feat_reduce(int F, const T* delta0, int b0, int nb, int P, int Hp, const float* grid, const float* img,
                      const float* x_explicit, float* Tm, cudaStream_t st);
int coord_param_grad_opt(int F, const float* Tm, int B, int H, int Hp, float* d_coord_w, float* d_coord_b,
                         cudaStream_t st);
int latent_coord_grad(int F, int B, int H, int Hp, const float* w, long w_img_stride, const float* Tm,
                      const float* img, float* out, cudaStream_t st);
template <typename T>
int coord_row_grad_opt(int F, const T* delta0, int rows, int P, int H, int Hp, const float* w, long w_img_stride,
                       const float* x, float* g_x, cudaStream_t st);


// g_pre[i] = g[i] * act'(out[i]) with act' expressed through the layer OUTPUT (ResidLinear.forward's backward)
int act_backward(const float* out, const float* g, float* g_pre, long n, int act, cudaStream_t st);

int adam_tick(int* t_dev, float* bias_corr_dev, float b1, float b2, cudaStream_t st);
// bias_corr_dev: optional device pointer to {1 - b1^t, sqrt(1 - b2^t)}; when set it overrides t
int adam(float* p, float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, int t,
         int zero_grad, const float* bias_corr_dev, cudaStream_t st);
int gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n_rows, int64_t row_len, cudaStream_t st);
// dst (rows_p x cols_p, bf16, zero padded) = src (rows x cols fp32)
int to_bf16_padded(const float* src, int rows, int cols, __nv_bfloat16* dst, int rows_p, int cols_p, cudaStream_t st);

// Pillow-compatible bicubic rotation of B images (h, w, C) about their centres (see step_kernels.cu)
int rotate_bicubic(const float* src, float* dst, const double* mat, const int* mode, int B, int h, int w, int C,
                   int quantize_u8, cudaStream_t st);

// ResidLinear helpers of the tensor-core path: dst[i,i] = bf16(src[i,i] + 1) on a converted weight; fp32 W + I
int add_identity_bf16(const float* src, int n, __nv_bfloat16* dst, int ld, cudaStream_t st);
int copy_add_identity(const float* src, int n, float* dst, cudaStream_t st);

// per-particle real-space CTF kernels (ingest_kernels.cu): params (N, 8) fp64 device rows
// [defocus um, cs mm, voltage kV, apix, bfactor, ampcont %, dfdiff, dfang] -> out (N, n, m) fp32
int ctf_filter(const double* params, int N, int n, int m, double scale, float* out, cudaStream_t st);

// SM clock in MHz measured on the device (20 us spin of one thread)
int clock_probe(float* out_mhz, cudaStream_t st);

// fp32 -> (hi, hi|lo, lo|hi) bf16 terms for the 3-MMA error-compensated encoder GEMMs (see step_kernels.cu)
int split3(const float* src, int rows, int cols, long ld, __nv_bfloat16* dst, int rows_p, int cols_p, int kcat,
           int pattern, cudaStream_t st);
// one pass over src writing both layouts: dst_r row-stacked (kcat = 0) with the terms of pattern_r, dst_k
// K-concatenated (kcat = 1) with those of pattern_k
int split3_both(const float* src, int rows, int cols, long ld, __nv_bfloat16* dst_r, int pattern_r, __nv_bfloat16* dst_k,
                int pattern_k, int rows_p, int cols_p, cudaStream_t st);
// the same after finishing a layer whose GEMM left RAW sums in `raw` (rows x cols_p fp32, ld = cols_p):
// h = act(raw + bias) is written back IN PLACE (padded columns: act(0)) and split into dst (dst NULL: no split)
int split3_act(float* raw, const float* bias, int act, int rows, int cols, __nv_bfloat16* dst, int rows_p, int cols_p,
               int kcat, int pattern, cudaStream_t st);

}  // namespace svae
