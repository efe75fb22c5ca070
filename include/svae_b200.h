/*
 * svae_b200.h -- C ABI of libsvae_b200.so: the spatial-VAE training-step hot path on B200 (sm_100a).
 *
 * The reference (cfframe/spatial-VAE) is pure Python on PyTorch and defines no native ABI, so
 * these entry points are what a binding for this path replaces, one per reference call site
 * (paths relative to the reference checkout):
 *
 *   svae_encoder_forward / _backward   InferenceNetwork.forward             spatial_vae/models.py:46-54
 *   svae_decoder_forward / _backward   SpatialGenerator.forward             spatial_vae/models.py:90-132
 *   svae_step                          eval_minibatch + loss.backward()     train_mnist.py:24-90,147-148
 *                                                                           train_particles.py:22-148,178-179
 *                                                                           train_galaxy.py:27-128,207-208
 *   svae_adam_step                     optim.step(); optim.zero_grad()      train_mnist.py:149-150 (Adam :389-392)
 *   svae_gather_rows                   DataLoader(TensorDataset, shuffle)   train_mnist.py:334,395-396
 *   svae_rotate_bicubic                PIL Image.rotate(BICUBIC) loop       train_particles.py:28-43, train_galaxy.py:36-54
 *   svae_ctf_filter                    ctf_filter's per-particle ifft2 loop spatial_vae/ctf.py:33-56
 *   svae_resid_linear_forward/_backward ResidLinear.forward                 spatial_vae/models.py:13-21
 *   svae_gemm_bf16                     one nn.Linear of SpatialGenerator.layers (models.py:82,126) on
 *                                      tcgen05 tensor cores (building block, exposed for tests)
 *   svae_gemm_dw_top, svae_gemm_dx_moments  fused head / tail of loss.backward() through SpatialGenerator.layers
 *                                      (building blocks of svae_step, exposed for tests)
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers, sizes, a cudaStream_t passed as void*.
 *   - the caller owns every buffer (parameters, gradients, Adam state, inputs, outputs, workspace);
 *     the library never allocates device memory, never retains a caller pointer between calls and never
 *     synchronises: all work is enqueued on `stream` (CUDA-graph capturable) or forked from and joined back to it (5).
 *   - process-wide state, all of it: (1) four environment switches read ONCE per process, for A/B comparisons:
 *     SVAE_TC_CTA_GROUP=1 (single-CTA instead of CTA-pair tensor-core GEMMs), SVAE_RESID_TC=0 (ResidLinear networks
 *     on the fp32 kernels in FAST precision), SVAE_CTF_FAST=0 (generic instead of register-tiled 39x39 CTF
 *     correlation), SVAE_SIDE_STREAM=0 (see 5); they select between kernels / schedules that pass the same parity
 *     tests; (2) the per-kernel "max dynamic shared memory" attribute, set on a kernel's first launch; (3) the launch counter behind svae_launch_count(); (4) the last error text per thread;
 *     (5) ONE auxiliary non-blocking stream and two events per host thread and device, created on the first training
 *     call of svae_step: independent small kernels of the backward tail (parameter-gradient chains) are forked onto
 *     it and joined back onto `stream` before the call returns, so nothing is left running there and a captured CUDA
 *     graph simply gets parallel branches.
 *   - accumulated outputs (gradient buffers, `+=` in the text below) are summed with fp32 atomics across thread
 *     blocks: the summation ORDER, hence the last bits of a gradient, can differ between two runs on the same input
 *     (|difference| ~ 1e-7 relative; the reference's cuBLAS split-K GEMMs behave the same way).  The inference
 *     network's small-batch GEMMs split K over thread blocks in the same manner (hidden layers in FAST / PARITY_TC,
 *     the head in every precision), so the per-image statistics carry the same last-bit variation.
 *   - svae_step / svae_decoder_* walk the minibatch in chunks of whole images sized so that the activation part of the
 *     workspace stays near 6 GiB (cfg.chunk_images = 0) or in chunks of cfg.chunk_images; svae_workspace_bytes()
 *     reports the size for that choice, results do not depend on it, and a workspace smaller than reported is
 *     refused with SVAE_ENOSPACE (never silently re-chunked).
 *   - every entry returns 0 on success or a negative SVAE_E* code; svae_last_error() gives text.
 *   - there is NO CPU fallback: without a CUDA device the calls fail with SVAE_ECUDA.
 *   - tensors are row-major fp32 unless stated; weights use nn.Linear layout (out_features, in_features).
 */
#ifndef SVAE_B200_H
#define SVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVAE_OK       0
#define SVAE_EINVAL  (-1)   /* unsupported shape / flag combination (text in svae_last_error) */
#define SVAE_EALIGN  (-2)   /* a pointer or leading dimension violates an alignment rule      */
#define SVAE_ECUDA   (-3)   /* CUDA runtime/driver error                                      */
#define SVAE_ENOSPACE (-4)  /* workspace too small                                            */

#define SVAE_MAX_LAYERS 8

/* activation codes: nn.Tanh, nn.LeakyReLU(0.01), nn.ReLU, nn.Sigmoid (train_galaxy.py:426-434) */
enum { SVAE_ACT_TANH = 0, SVAE_ACT_LEAKYRELU = 1, SVAE_ACT_RELU = 2, SVAE_ACT_SIGMOID = 3 };
/* likelihoods: Bernoulli (train_mnist.py:80-81), Gaussian unit variance and Gaussian with the
 * learned ("fit-noise") variance (train_particles.py:136-139) */
enum { SVAE_LIK_BERNOULLI = 0, SVAE_LIK_GAUSS = 1, SVAE_LIK_GAUSS_FITNOISE = 2 };
/* precision of the decoder hidden-layer GEMMs:
 *   PARITY  fp32 FFMA kernels everywhere (bit-for-bit comparable with the fp32 reference up to
 *           summation order);
 *   FAST    bf16 operands on tcgen05 tensor cores, fp32 TMEM accumulation; first layer, output
 *           layer, likelihood, KL, reductions, encoder, dW accumulation and Adam stay fp32.
 *   PARITY_TC  fp32 activations; every hidden GEMM of both networks runs on tcgen05 as ONE bf16 GEMM over three
 *           hi/lo split terms of its operands (A_hi B_hi + A_hi B_lo + A_lo B_hi, error ~2^-16, fp32 TMEM
 *           accumulation, accurate tanhf in the epilogue); everything else as PARITY.  Meets both parity gates
 *           (per-image ELBO 1e-3, parameters 1e-4 after 10 Adam steps) on tensor cores. */
enum { SVAE_PRECISION_PARITY = 0, SVAE_PRECISION_FAST = 1, SVAE_PRECISION_PARITY_TC = 2 };

typedef struct {
    int32_t B;        /* images in this call (this rank's slice of the minibatch)              */
    int32_t P;        /* coordinate rows per image (n_rows * n_cols)                           */
    int32_t n_rows;   /* image height (only used by the CTF correlation)                       */
    int32_t n_cols;   /* image width                                                           */
    int32_t C;        /* decoder n_out: 1, 2 (fit-noise) or 3 (RGB)                            */
    int32_t Cin;      /* channels per pixel seen by the encoder; encoder input = P * Cin       */
    int32_t Z;        /* unstructured latent dimension                                         */
    int32_t I;        /* inference dimension = Z + rotate + 2*translate                        */
    int32_t H;        /* decoder hidden width                                                  */
    int32_t L;        /* decoder num_layers: L-1 hidden HxH Linears (models.py:78-83)          */
    int32_t Hq;       /* encoder hidden width                                                  */
    int32_t Lq;       /* encoder num_layers: Lq hidden Linears, then the 2*I head (models.py:31-41) */
    int32_t k_ctf;    /* CTF kernel side length, 0 = no CTF (train_particles.py:112-119)       */
} SvaeShape;

typedef struct {
    int32_t rotate;          /* latent column 0 is theta (train_mnist.py:42-59)                */
    int32_t translate;       /* next two columns are dx (train_mnist.py:65-74)                 */
    int32_t likelihood;      /* SVAE_LIK_*                                                     */
    int32_t theta_kl_mean;   /* 1: mnist theta-KL with mean penalty (train_mnist.py:63); 0: particles/galaxy */
    int32_t activation;      /* SVAE_ACT_*                                                     */
    int32_t precision;       /* SVAE_PRECISION_*                                               */
    int32_t softplus;        /* softplus on output channel 0 after the sigmoid (models.py:129-130) */
    int32_t chunk_images;    /* images per decoder pass (bounds the activation workspace); 0 = library default */
    float   theta_prior;     /* std of the rotation prior                                      */
    float   dx_scale;        /* std of the translation prior                                   */
    float   z_scale;         /* multiplies z only (train_particles.py:99); mnist: 1            */
    float   grad_scale;      /* 1 / (global minibatch size): gradients are of -mean_b(elbo_b)  */
    int32_t resid;           /* hidden layers of BOTH networks are ResidLinear, act(W h + b + h) (models.py:13-21,35-36,79-80) */
    int32_t expand_coords;   /* first layer sees (x0, x1, x0^2, x1^2, x0*x1): coord_w is (H, 5) (models.py:65-67,99-102) */
    int32_t bilinear;        /* h0 += Bilinear(features, z) (models.py:74-75,114-121); needs bilinear_w and Z > 0 */
} SvaeConfig;

/* SpatialGenerator parameters (models.py:69-87). Gradient structs use the same layout. */
typedef struct {
    float* coord_w;                       /* (H, 2)  coord_linear.weight; (H, 5) with expand_coords */
    float* coord_b;                       /* (H)     coord_linear.bias                         */
    float* latent_w;                      /* (H, Z)  latent_linear.weight, NULL when Z == 0    */
    float* hidden_w[SVAE_MAX_LAYERS];     /* (H, H)  layers.{1,3,..}.weight, L-1 entries       */
    float* hidden_b[SVAE_MAX_LAYERS];     /* (H)                                               */
    float* out_w;                         /* (C, H)  last Linear                               */
    float* out_b;                         /* (C)                                               */
    float* bilinear_w;                    /* (H, F, Z) bilinear.weight, F = 2 or 5; NULL unless cfg.bilinear */
} SvaeDecoderParams;

/* InferenceNetwork parameters (models.py:31-41): Lq hidden Linears then the head (2I, Hq). */
typedef struct {
    float* w[SVAE_MAX_LAYERS + 1];
    float* b[SVAE_MAX_LAYERS + 1];
} SvaeEncoderParams;

/* Inputs of one eval_minibatch. */
typedef struct {
    const float* grid;          /* (P, 2) pixel coordinates (train_mnist.py:316-320)             */
    const float* y;             /* (B, P*C_target) targets; C_target = Cin                       */
    const float* y_enc;         /* (B, P*Cin) what the encoder sees; NULL = y (augmentation, train_particles.py:28-50) */
    const float* theta_offset;  /* (B) added to theta before rotating (train_particles.py:71-74); may be NULL */
    const float* eps;           /* (B, I) the N(0,1) draw of train_mnist.py:38; NULL = draw it in the kernel (rng_* below) */
    const float* ctf;           /* (B, k_ctf, k_ctf) real-space kernels or NULL                  */
    const uint8_t* mask;        /* (P) 0/1 pixel mask or NULL (train_particles.py:126-132)       */
    /* In-kernel draw of eps (used when eps == NULL): Philox4x32-10 keyed by rng_seed, counter = (global image index,
     * latent block, *rng_step), Box-Muller.  The value of image g of the minibatch depends only on (seed, step, g), so
     * a data-parallel run draws the same eps however the minibatch is split across ranks (SURVEY 7.2 "RNG parity").
     * rng_step is a DEVICE pointer so that a captured CUDA graph draws fresh numbers on every replay. */
    const int32_t* rng_step;    /* device pointer to the step counter; required when eps == NULL  */
    uint64_t rng_seed;
    int64_t  rng_image_offset;  /* global minibatch index of this call's first image (this rank's slice) */
    /* Optional cudaEvent_t (NULL = none), recorded by a training call at the point of its work where every DECODER
     * parameter gradient is final, i.e. before the encoder backward (train_mnist.py:147 runs them in this order): a
     * data-parallel caller waits for it on a second stream and starts reducing the decoder gradients while the
     * encoder backward still runs.  Capturable: the event becomes a node of the caller's CUDA graph. */
    void* decoder_grads_event;
} SvaeStepInputs;

/* Outputs of one eval_minibatch (all optional except stats). */
typedef struct {
    float* stats;        /* (B, 3): per image [logp_i, kl_i, elbo_i]; the reference's scalars are their batch means */
    float* y_hat;        /* (B, P, C) decoder output or NULL                                   */
    float* latent;       /* (B, I) sampled latent or NULL                                      */
    float* stats_sum;    /* (4) or NULL: [sum logp_i, sum kl_i, sum elbo_i, 0] over this call's images, WRITTEN (not
                            accumulated) by one thread block in a fixed order: what train_mnist.py:152-165 accumulates
                            on the host; a data-parallel caller points it at the tail of its gradient buffer so the
                            loss sums ride in the gradient allreduce */
} SvaeStepOutputs;

int  svae_version(void);
/* copies the last error text of the calling thread into buf (NUL terminated); returns its length */
int  svae_last_error(char* buf, int n);
/* kernels launched by this library since it was loaded (all threads); evidence for bench.py */
unsigned long long svae_launch_count(void);
/* number of SMs of the current device, <0 on error (used by the host to size row chunks) */
int  svae_device_sm_count(void);

/* Bytes of caller-provided scratch svae_step / svae_decoder_* need for this shape+config. */
int  svae_workspace_bytes(const SvaeShape* shape, const SvaeConfig* cfg, size_t* bytes);

/* InferenceNetwork.forward: x (B, P*Cin) -> out (B, 2I) = [z_mu | z_logstd].
 * acts: scratch (Lq, B, Hq) receiving the hidden activations (needed by the backward).
 * activation: SVAE_ACT_*, OR-ed with SVAE_ENC_RESID when hidden layers 1..Lq-1 are ResidLinear (models.py:35-36). */
#define SVAE_ENC_RESID 0x100
int  svae_encoder_forward(const SvaeShape* shape, int activation, const SvaeEncoderParams* params,
                          const float* x, float* out, float* acts, void* stream);
/* Backward of the above: g_out (B, 2I) is overwritten as scratch; grads are ACCUMULATED (+=)
 * into `grads` (zero them first); g_x (B, P*Cin) optional. scratch: (2, B, Hq) floats. */
int  svae_encoder_backward(const SvaeShape* shape, int activation, const SvaeEncoderParams* params,
                           const float* x, const float* acts, float* g_out, SvaeEncoderParams* grads,
                           float* g_x, float* scratch, void* stream);

/* SpatialGenerator.forward on explicit coordinates: x (B, P, 2), z (B, Z) -> y_hat (B, P, C). */
int  svae_decoder_forward(const SvaeShape* shape, const SvaeConfig* cfg, const SvaeDecoderParams* params,
                          const float* x, const float* z, float* y_hat, void* workspace, size_t workspace_bytes,
                          void* stream);
/* Backward of the above given g_y (B, P, C) = dLoss/dy_hat: accumulates parameter gradients
 * into `grads`, writes g_x (B, P, 2) and g_z (B, Z) when non-NULL. Re-runs the forward. */
int  svae_decoder_backward(const SvaeShape* shape, const SvaeConfig* cfg, const SvaeDecoderParams* params,
                           const float* x, const float* z, const float* g_y, SvaeDecoderParams* grads,
                           float* g_x, float* g_z, void* workspace, size_t workspace_bytes, void* stream);

/* One eval_minibatch: encoder -> reparameterise -> rotate/translate -> decoder -> ELBO.
 * When dec_grads/enc_grads are non-NULL also the backward of -grad_scale * sum_b elbo_b,
 * ACCUMULATED into the gradient buffers (zero them, or keep them for gradient accumulation). */
int  svae_step(const SvaeShape* shape, const SvaeConfig* cfg,
               const SvaeDecoderParams* dec, const SvaeEncoderParams* enc,
               const SvaeStepInputs* in, const SvaeStepOutputs* out,
               SvaeDecoderParams* dec_grads, SvaeEncoderParams* enc_grads,
               void* workspace, size_t workspace_bytes, void* stream);

/* Fused Adam over a flat fp32 buffer (torch.optim.Adam semantics, no weight decay, no amsgrad);
 * t is the 1-based step count; zero_grad != 0 clears the gradient afterwards (optim.zero_grad()). */
int  svae_adam_step(float* param, float* grad, float* m, float* v, size_t n,
                    float lr, float beta1, float beta2, float eps, int t, int zero_grad, void* stream);

/* Adam for CUDA-graph replays.  svae_adam_tick increments the device-resident step counter *t_dev and writes the
 * step's bias corrections {1 - beta1^t, sqrt(1 - beta2^t)} to bias_corr_dev (2 floats, device);
 * svae_adam_step_graph is svae_adam_step reading those scalars from device memory, so a captured
 * tick + step pair can be replayed any number of times without host-side state. */
int  svae_adam_tick(int32_t* t_dev, float* bias_corr_dev, float beta1, float beta2, void* stream);
int  svae_adam_step_graph(float* param, float* grad, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                          float eps, const float* bias_corr_dev, int zero_grad, void* stream);

/* dst[i, :] = src[index[i], :] for i < n_rows: one launch replaces the per-sample DataLoader fetch. */
int  svae_gather_rows(const float* src, const int64_t* index, float* dst, int64_t n_rows, int64_t row_len,
                      void* stream);

/* --augment-rotation on the device (reference train_particles.py:28-43, train_galaxy.py:36-54, which loop over
 * the minibatch on the host with PIL): rotates B images (n_rows, n_cols, channels), fp32, about their centres,
 * bit-compatible with Pillow's Image.rotate(angle, resample=BICUBIC).  inv_affine: (B, 6) float64 destination->
 * source matrices as PIL/Image.py builds them; mode: (B) 0 = general, 1 = copy (angle 0), 2 = 180, 3 = 90, 4 = 270
 * degrees (Pillow's transpose fast paths).  quantize_u8 != 0 reproduces the galaxy driver's round trip through
 * uint8: sample = (uint8)(x*255), result = clip/truncate to uint8, then / 255. */
/* HOST helper (host pointers, no device work): the (B, 6) destination->source matrices and (B) modes that
 * svae_rotate_bicubic takes, for counter-clockwise angles in degrees, built exactly as PIL/Image.py rotate() does
 * (angle % 360, math.radians, cos/sin rounded to 15 decimals, rotation about (w/2, h/2)). */
int  svae_rotation_matrices(const double* angles_deg, int B, int n_rows, int n_cols, double* inv_affine,
                            int32_t* mode);
int  svae_rotate_bicubic(const float* src, float* dst, const double* inv_affine, const int32_t* mode, int B,
                         int n_rows, int n_cols, int channels, int quantize_u8, void* stream);

/* Real-space CTF kernels on the device, all particles in one launch (reference spatial_vae/ctf.py:33-56, a Python loop
 * with one numpy ifft2 per particle): params (n_particles, 8) fp64 DEVICE rows in the column order of the CTF table
 * [defocus um, cs mm, voltage kV, apix A/pixel, bfactor, ampcont %, dfdiff, dfang] (ctf.py:27-30);
 * out (n_particles, n, m) fp32 = -fftshift(ifft2(CTF)).real, computed in fp64. */
int  svae_ctf_filter(const double* params, int n_particles, int n, int m, double scale, float* out, void* stream);

/* Measures the SM clock on the device: enqueues a one-thread kernel that spins ~20 us and writes
 * cycles/time in MHz to *out_mhz (device pointer).  Measurement aid for bench.py. */
int  svae_sm_clock_probe(float* out_mhz, void* stream);

/* bf16 tensor-core GEMM building blocks (tcgen05 / TMEM / TMA), fp32 accumulation.
 *   mode 0  FWD : out[M,N]  = act(A[M,K] * W[N,K]^T + bias[N])                   (bf16 out)
 *   mode 1  DX  : out[M,N]  = (A[M,K] * W[K,N]) .* act'(aux[M,N])                (bf16 out)
 *   mode 2  DW  : outf[M,N] += A[Kr,M]^T * Bm[Kr,N]   (fp32 out, accumulated with atomics)
 * All bf16 matrices are row-major with leading dimension ld* (elements, multiple of 8). */
int  svae_gemm_bf16(int mode, int M, int N, int K,
                    const void* A, int lda, const void* W, int ldw,
                    const float* bias, const void* aux, int ldaux, int activation,
                    void* out, int ldo, void* stream);

/* ResidLinear.forward (reference spatial_vae/models.py:13-21): out (rows, n) = act(x W^T + b + x), W (n, n), fp32 FFMA
 * with the skip connection in the GEMM epilogue.  The backward takes the forward's out and g_out = dLoss/dout, uses
 * g_pre (rows, n) as scratch, ACCUMULATES g_w (n, n) and g_b (n) and writes g_x (rows, n) = g_pre W + g_pre. */
int  svae_resid_linear_forward(const float* x, const float* w, const float* b, float* out, int rows, int n,
                               int activation, void* stream);
int  svae_resid_linear_backward(const float* x, const float* w, const float* out, const float* g_out, float* g_pre,
                                float* g_x, float* g_w, float* g_b, int rows, int n, int activation, void* stream);

/* Fused tail of the decoder backward (loss.backward() through SpatialGenerator.forward's first layer,
 * models.py:104-124, train_mnist.py:50-59,70-74): delta_0 = (delta (rows, Hp) * W (Hp, Hp; [j][n])) .* act'(h_0) is
 * never stored; h_0[row, n] = act(coord_w[n,0] x' + coord_w[n,1] y' + hz[b, n]) is recomputed from the pixel grid
 * (P, 2) and the per-image transforms img (B, 4) = cos, sin, dx0, dx1, and the per-image column moments
 * S (B, 3, Hp) += sum_p delta_0 {1, grid_x, grid_y} are accumulated with fp32 atomics (zero S first; summation
 * order, hence the last bits, varies from run to run).  Building block of svae_step, exposed for tests. */
int  svae_gemm_dx_moments(int rows, int H, int Hp, const void* delta, int ldd, const void* W, int ldw, int activation,
                          const float* grid, const float* img, const float* coord_w, const float* hz, float* S, int P,
                          void* stream);

/* Fused head of the decoder backward (loss.backward() through the output Linear and the top hidden Linear,
 * models.py:82-85): delta (rows, Hp) = (g_o (rows, C) * out_w (C, H)) .* act'(h_top) is built on the fly from h_top
 * (rows, Hp bf16) inside the weight-gradient GEMM  dW (H, H) += delta^T h_prev,  which also accumulates
 * d_out_w (C, H) += g_o^T h_top, d_out_b (C) += colsum(g_o), d_b (H) += colsum(delta) (d_b may be NULL) and writes delta
 * to delta_out (rows, Hp bf16; may be NULL).  All accumulations are fp32 atomics (+=, order not deterministic).
 * C = 1..3; g_o is copied in 16-byte units: it must be 16-byte aligned and readable up to the next 16-byte boundary
 * past its last element.  Building block of svae_step, exposed for tests. */
int  svae_gemm_dw_top(int rows, int H, int Hp, const void* h_top, const void* h_prev, int activation, const float* g_o,
                      int C, const float* out_w, float* d_out_w, float* d_out_b, float* d_b, float* dW, void* delta_out,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVAE_B200_H */
