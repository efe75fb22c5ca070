// SIMT kernels of the spatial-VAE step: everything that is not a dense HxH contraction.
// Reference line numbers are relative to the reference checkout (see include/svae_b200.h).
#include <stdlib.h>

#include <type_traits>

#include "kernels.cuh"

namespace svae {

// ------------------------------------------------------------------------------------------------
// small load/store helpers for pairs of adjacent columns of an activation row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load2(const float* p, float& a, float& b) {
    float2 v = *reinterpret_cast<const float2*>(p);
    a = v.x; b = v.y;
}
__device__ __forceinline__ void load2(const __nv_bfloat16* p, float& a, float& b) {
    __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
    a = __low2float(v); b = __high2float(v);
}
__device__ __forceinline__ void store2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// eight adjacent columns of a row: 16-byte (bf16) or 2 x 16-byte (fp32) accesses
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (w == 0) t = warp_sum(t);
    return t;  // valid in warp 0
}

// ------------------------------------------------------------------------------------------------
// reparameterisation + KL   (train_mnist.py:33-39,62-63,84-86; particles :85-86,99)
// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> four uniform 32-bit words
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// element i of the N(0,1) row of global image g at step t: Box-Muller on the Philox block (g, i / 4, t)
__device__ __forceinline__ float philox_normal(uint64_t seed, long g, int i, int step) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)(i >> 2), (uint32_t)step),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t a = (i & 2) ? r.z : r.x, b = (i & 2) ? r.w : r.y;
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.f / 16777216.f);          // (0, 1)
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.f / 16777216.f);
    const float rad = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    return rad * ((i & 1) ? sn : cs);
}

__global__ void __launch_bounds__(256) latent_forward_k(SvaeShape s, SvaeConfig c, const float* __restrict__ zo,
                                                        const float* __restrict__ eps, const float* __restrict__ toff,
                                                        float* __restrict__ lat, float* __restrict__ img,
                                                        float* __restrict__ zs, float* __restrict__ stats,
                                                        LatentRng rng, float* __restrict__ eps_out) {
    // one warp per image, lanes over the I latent dimensions
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= s.B) return;
    const int I = s.I;
    const float* mu = zo + (long)b * 2 * I;
    const float* ls = mu + I;
    const int rot = c.rotate ? 1 : 0;
    const int zcol = rot + (c.translate ? 2 : 0);
    const int step = (eps == nullptr) ? *rng.step : 0;
    float kl = 0.f;
    for (int i = lane; i < I; i += 32) {
        const float m = mu[i], l = ls[i];
        const float sd = expf(l);
        float e;
        if (eps != nullptr) {
            e = eps[(long)b * I + i];
        } else {
            e = philox_normal(rng.seed, rng.image_offset + b, i, step);
            eps_out[(long)b * I + i] = e;
        }
        const float v = sd * e + m;
        if (lat) lat[(long)b * I + i] = v;
        if (rot && i == 0) {
            const float th = v + (toff ? toff[b] : 0.f);
            img[b * 4 + 0] = cosf(th);
            img[b * 4 + 1] = sinf(th);
            const float sp = c.theta_prior;
            const float num = c.theta_kl_mean ? (sd * sd + m * m) : (sd * sd);
            kl += -l + logf(sp) + num / 2.f / (sp * sp) - 0.5f;
        } else {
            kl += -l + 0.5f * sd * sd + 0.5f * m * m - 0.5f;
            if (i < zcol) img[b * 4 + 2 + (i - rot)] = v * c.dx_scale;
            else zs[(long)b * s.Z + (i - zcol)] = v * c.z_scale;
        }
    }
    if (lane == 0) {
        if (!rot) { img[b * 4 + 0] = 1.f; img[b * 4 + 1] = 0.f; }
        if (!c.translate) { img[b * 4 + 2] = 0.f; img[b * 4 + 3] = 0.f; }
    }
    kl = warp_sum(kl);
    if (lane == 0) stats[b * 3 + 1] = kl;
}

int latent_forward(const SvaeShape& s, const SvaeConfig& c, const float* zo, const float* eps,
                   const float* theta_offset, float* lat, float* img, float* zs, float* stats, cudaStream_t st,
                   const LatentRng& rng, float* eps_out) {
    SVAE_REQUIRE(eps != nullptr || (rng.step != nullptr && eps_out != nullptr), SVAE_EINVAL,
                 "either eps or the in-kernel generator (rng_step) is required");
    latent_forward_k<<<ceil_div(s.B, 8), 256, 0, st>>>(s, c, zo, eps, theta_offset, lat, img, zs, stats, rng, eps_out);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

__global__ void fill_rows_k(float* dst, const float* row, int rows, int n, int ld) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * n) return;
    int r = (int)(i / n), c = (int)(i % n);
    dst[(long)r * ld + c] = row[c];
}
int fill_rows(float* dst, const float* row, int rows, int n, int ld, cudaStream_t st) {
    fill_rows_k<<<ceil_div((long)rows * n, 256), 256, 0, st>>>(dst, row, rows, n, ld);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// first layer: rotate/translate the grid in registers, K=2 coordinate layer + per-image z projection
// (train_mnist.py:50-59,70-74; models.py:104-124 and layers[0])
// ------------------------------------------------------------------------------------------------
constexpr int L0_MAX_ROWS = 256;
// rows per block: an even split of the image into blocks of at most 224 rows.  Large blocks amortise the per-thread
// weight / hz loads: measured at C2 (P = 784) 64 rows 207 us, 128 rows 160 us, 196 rows 135 us, 256 rows 140 us.
static int l0_rows(int P) { const int blocks = ceil_div(P, 224); return ceil_div(P, blocks); }

template <typename T, bool FAST>
__global__ void __launch_bounds__(256) layer0_k(int P, int act, int b0, const float* __restrict__ coord_w,
                                                const float* __restrict__ hz, const float* __restrict__ grid,
                                                const float* __restrict__ img, const float* __restrict__ xe,
                                                int H, int Hp, T* __restrict__ h0, int rows_per_block) {
    __shared__ float sx[L0_MAX_ROWS][2];
    const int bl = blockIdx.y;          // image within the chunk
    const int b = b0 + bl;              // image within the call
    const int p0 = blockIdx.x * rows_per_block;
    const int nrows = min(rows_per_block, P - p0);
    for (int t = threadIdx.x; t < nrows; t += blockDim.x) {
        const int p = p0 + t;
        float x0, x1;
        if (xe) {
            x0 = xe[((long)b * P + p) * 2 + 0];
            x1 = xe[((long)b * P + p) * 2 + 1];
        } else {
            const float g0 = grid[p * 2 + 0], g1 = grid[p * 2 + 1];
            const float cs = img[b * 4 + 0], sn = img[b * 4 + 1];
            x0 = g0 * cs - g1 * sn + img[b * 4 + 2];
            x1 = g0 * sn + g1 * cs + img[b * 4 + 3];
        }
        sx[t][0] = x0;
        sx[t][1] = x1;
    }
    __syncthreads();
    T* out = h0 + ((long)bl * P + p0) * Hp;
    if ((Hp & 7) == 0) {
        // thread = (8-column group, row lane): each row is written with one 16-byte store per thread
        const int groups = Hp >> 3;
        for (int g = threadIdx.x % min(groups, 256); g < groups; g += 256) {
            const int lanes = max(1, 256 / groups);
            const int rl = threadIdx.x / groups;
            if (rl >= lanes) break;
            // column pairs in packed (f32x2) registers: half the FMA issue slots, the MUFU pipe (tanh) sets the pace
            float2 w0[4], w1[4], hb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = g * 8 + 2 * e;
                w0[e].x = n < H ? coord_w[n * 2 + 0] : 0.f;         w0[e].y = n + 1 < H ? coord_w[n * 2 + 2] : 0.f;
                w1[e].x = n < H ? coord_w[n * 2 + 1] : 0.f;         w1[e].y = n + 1 < H ? coord_w[n * 2 + 3] : 0.f;
                hb[e].x = n < H ? hz[(long)b * Hp + n] : 0.f;       hb[e].y = n + 1 < H ? hz[(long)b * Hp + n + 1] : 0.f;
            }
#pragma unroll 2
            for (int r = rl; r < nrows; r += lanes) {
                const float2 x0 = make_float2(sx[r][0], sx[r][0]), x1 = make_float2(sx[r][1], sx[r][1]);
                float v[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = __ffma2_rn(w0[e], x0, __ffma2_rn(w1[e], x1, hb[e]));
                    v[2 * e] = act_apply<FAST>(act, a.x);
                    v[2 * e + 1] = act_apply<FAST>(act, a.y);
                }
                store8(out + (long)r * Hp + g * 8, v);
            }
        }
        return;
    }
    for (int n = threadIdx.x; n < Hp; n += blockDim.x) {
        float w0 = 0.f, w1 = 0.f, hb = 0.f;
        if (n < H) {
            w0 = coord_w[n * 2 + 0];
            w1 = coord_w[n * 2 + 1];
            hb = hz[(long)b * Hp + n];
        }
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
            const float a = fmaf(w0, sx[r][0], fmaf(w1, sx[r][1], hb));
            out[(long)r * Hp + n] = from_f32<T>(act_apply<FAST>(act, a));
        }
    }
}

template <typename T>
int layer0_forward(const SvaeShape& s, int act, int b0, int nb, const float* coord_w, const float* hz,
                   const float* grid, const float* img, const float* x_explicit, int H, int Hp, T* h0,
                   cudaStream_t st) {
    dim3 g(ceil_div(s.P, l0_rows(s.P)), nb);
    constexpr bool FAST = !std::is_same<T, float>::value;
    layer0_k<T, FAST><<<g, 256, 0, st>>>(s.P, act, b0, coord_w, hz, grid, img, x_explicit, H, Hp, h0, l0_rows(s.P));
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int layer0_forward<float>(const SvaeShape&, int, int, int, const float*, const float*, const float*,
                                   const float*, const float*, int, int, float*, cudaStream_t);
template int layer0_forward<__nv_bfloat16>(const SvaeShape&, int, int, int, const float*, const float*,
                                           const float*, const float*, const float*, int, int, __nv_bfloat16*,
                                           cudaStream_t);

// ------------------------------------------------------------------------------------------------
// output layer: warp-shuffle dot product per row, sigmoid (+softplus)   (models.py:84-85,129-130)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// value of output element with channel ch and its derivative w.r.t. the logit
__device__ __forceinline__ void post_output(float o, int ch, int softplus, float& v, float& dv) {
    const float sg = sigmoidf_(o);
    if (softplus && ch == 0) {
        v = log1pf(expf(sg));
        dv = sigmoidf_(sg) * sg * (1.f - sg);
    } else {
        v = sg;
        dv = sg * (1.f - sg);
    }
}

template <typename T, int C>
__global__ void __launch_bounds__(256) out_forward_k(const T* __restrict__ h, int rows, int H, int Hp,
                                                     const float* __restrict__ out_w,
                                                     const float* __restrict__ out_b, int softplus,
                                                     float* __restrict__ o, float* __restrict__ y_hat) {
    extern __shared__ float sw[];  // C x Hp, zero padded
    for (int i = threadIdx.x; i < C * Hp; i += blockDim.x) {
        const int c = i / Hp, n = i % Hp;
        sw[i] = (n < H) ? out_w[c * H + n] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long r = (long)blockIdx.x * 8 + warp; r < rows; r += (long)gridDim.x * 8) {
        const T* hr = h + r * Hp;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        for (int n = lane * 2; n < Hp; n += 64) {
            float a, b;
            load2(hr + n, a, b);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(a, sw[c * Hp + n], fmaf(b, sw[c * Hp + n + 1], acc[c]));
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = warp_sum(acc[c]);
        if (lane < C) {
            float val = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) if (lane == c) val = acc[c];
            val += out_b[lane];
            o[r * C + lane] = val;
            if (y_hat) {
                float v, dv;
                post_output(val, lane, softplus, v, dv);
                y_hat[r * C + lane] = v;
            }
        }
    }
}

template <typename T>
int out_forward(const T* h, int rows, int H, int Hp, int C, const float* out_w, const float* out_b, int softplus,
                float* o, float* y_hat, cudaStream_t st) {
    const int blocks = min(ceil_div(rows, 8), 148 * 8);
    const size_t smem = (size_t)C * Hp * sizeof(float);
    switch (C) {
        case 1: out_forward_k<T, 1><<<blocks, 256, smem, st>>>(h, rows, H, Hp, out_w, out_b, softplus, o, y_hat); break;
        case 2: out_forward_k<T, 2><<<blocks, 256, smem, st>>>(h, rows, H, Hp, out_w, out_b, softplus, o, y_hat); break;
        case 3: out_forward_k<T, 3><<<blocks, 256, smem, st>>>(h, rows, H, Hp, out_w, out_b, softplus, o, y_hat); break;
        case 4: out_forward_k<T, 4><<<blocks, 256, smem, st>>>(h, rows, H, Hp, out_w, out_b, softplus, o, y_hat); break;
        default: set_error("n_out=%d not supported (1..4)", C); return SVAE_EINVAL;
    }
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int out_forward<float>(const float*, int, int, int, int, const float*, const float*, int, float*, float*, cudaStream_t);
template int out_forward<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, const float*, const float*, int, float*, float*, cudaStream_t);

__global__ void logits_to_yhat_k(const float* __restrict__ o, float* __restrict__ y, long n, int C, int softplus) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v, dv;
    post_output(o[i], (int)(i % C), softplus, v, dv);
    y[i] = v;
}
int logits_to_yhat(const float* o, float* y_hat, long n, int C, int softplus, cudaStream_t st) {
    logits_to_yhat_k<<<ceil_div(n, 256), 256, 0, st>>>(o, y_hat, n, C, softplus);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// likelihood, one block per image
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) likelihood_k(SvaeShape s, SvaeConfig c, int b0, const float* __restrict__ o,
                                                    const float* __restrict__ y, const float* __restrict__ ctf,
                                                    const uint8_t* __restrict__ mask, float* __restrict__ stats,
                                                    float* __restrict__ g_o) {
    extern __shared__ float sm[];
    __shared__ float red[8];
    const int bl = blockIdx.x, b = b0 + bl;
    const int P = s.P, C = s.C;
    const long E = (long)P * C;
    const float* ob = o + (long)bl * E;
    float* gb = g_o ? g_o + (long)bl * E : nullptr;
    const float gs = c.grad_scale;
    float ll = 0.f;

    if (c.likelihood == SVAE_LIK_BERNOULLI) {
        // -BCE with both logs clamped at -100, backward divides by max(p(1-p),1e-12)
        // (train_mnist.py:80-81; ATen binary_cross_entropy)
        const float* yb = y + (long)b * E;
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            float v, dv;
            post_output(ob[e], e % C, c.softplus, v, dv);
            const float t = yb[e];
            const float lp = fmaxf(logf(v), -100.f), lq = fmaxf(log1pf(-v), -100.f);
            ll += t * lp + (1.f - t) * lq;
            if (gb) gb[e] = gs * (v - t) / fmaxf(v * (1.f - v), 1e-12f) * dv;
        }
    } else if (c.likelihood == SVAE_LIK_GAUSS_FITNOISE) {
        // y_params.view(B,-1): first P flat entries are the mean, last P the log-variance
        // (train_particles.py:102-110,136-137): the interleave quirk of SURVEY Appendix A #2.
        const float* yb = y + (long)b * P;
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            float mu, dmu, lv, dlv;
            post_output(ob[j], j % C, c.softplus, mu, dmu);
            post_output(ob[P + j], (P + j) % C, c.softplus, lv, dlv);
            const float m = mask ? (mask[j] ? 1.f : 0.f) : 1.f;
            const float r = mu - yb[j];
            const float iv = expf(-lv);
            ll -= 0.5f * m * (r * r * iv + lv);
            if (gb) {
                gb[j] = gs * m * r * iv * dmu;
                gb[P + j] = gs * 0.5f * m * (1.f - r * r * iv) * dlv;
            }
        }
    } else if (ctf == nullptr) {
        // unit-variance Gaussian (train_particles.py:138-139)
        const float* yb = y + (long)b * P;
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            float mu, dmu;
            post_output(ob[j], 0, c.softplus, mu, dmu);
            const float m = mask ? (mask[j] ? 1.f : 0.f) : 1.f;
            const float r = m * (mu - yb[j]);
            ll -= 0.5f * r * r;
            if (gb) gb[j] = gs * r * dmu;
        }
    } else {
        // CTF: per-image cross-correlation with its own k x k kernel, zero padding k/2
        // (train_particles.py:112-119), then the unit-variance Gaussian on the filtered mean.
        const int nr = s.n_rows, nc = s.n_cols, k = s.k_ctf, pad = k / 2;
        float* smu = sm;             // P
        float* sres = sm + P;        // P
        float* sk = sm + 2 * P;      // k*k
        const float* yb = y + (long)b * P;
        const float* kb = ctf + (long)b * k * k;
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            float mu, dmu;
            post_output(ob[j], 0, c.softplus, mu, dmu);
            smu[j] = mu;
        }
        for (int j = threadIdx.x; j < k * k; j += blockDim.x) sk[j] = kb[j];
        __syncthreads();
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            const int i0 = j / nc, j0 = j % nc;
            float acc = 0.f;
            for (int a = 0; a < k; ++a) {
                const int ii = i0 + a - pad;
                if (ii < 0 || ii >= nr) continue;
                const int blo = max(0, pad - j0), bhi = min(k, nc + pad - j0);
                const float* mrow = smu + ii * nc + (j0 - pad);
                const float* krow = sk + a * k;
                for (int bb = blo; bb < bhi; ++bb) acc = fmaf(mrow[bb], krow[bb], acc);
            }
            const float m = mask ? (mask[j] ? 1.f : 0.f) : 1.f;
            const float r = m * (acc - yb[j]);
            sres[j] = r;
            ll -= 0.5f * r * r;
        }
        __syncthreads();
        if (gb) {
            for (int j = threadIdx.x; j < P; j += blockDim.x) {
                const int u = j / nc, v = j % nc;
                float acc = 0.f;
                for (int a = 0; a < k; ++a) {
                    const int ii = u - a + pad;
                    if (ii < 0 || ii >= nr) continue;
                    // jj = v - bb + pad in [0, nc)
                    const int blo = max(0, v + pad - nc + 1), bhi = min(k, v + pad + 1);
                    const float* rrow = sres + ii * nc + (v + pad);
                    const float* krow = sk + a * k;
                    for (int bb = blo; bb < bhi; ++bb) acc = fmaf(rrow[-bb], krow[bb], acc);
                }
                float mu, dmu;
                post_output(ob[j], 0, c.softplus, mu, dmu);
                gb[j] = gs * acc * dmu;
            }
        }
    }
    const float tot = block_sum_256(ll, red);
    if (threadIdx.x == 0) stats[b * 3 + 0] = tot;
}

// ------------------------------------------------------------------------------------------------
// CTF fast path (k = 39, e.g. 40x40 particles): the decoded image and the residual live zero-padded in shared
// memory, each thread owns a strip of 8 adjacent output pixels and slides a 46-wide register window along the
// 39 taps of one kernel row: 22 128-bit shared loads per 312 FMAs instead of 2 loads per FMA.
// FLIP = false: cross-correlation (forward, train_particles.py:117); FLIP = true: its transpose (backward).
// ------------------------------------------------------------------------------------------------
template <int K, bool FLIP>
__device__ __forceinline__ void ctf_strip(const float* __restrict__ img_pad, int W, const float* __restrict__ sk,
                                          int i, int j0, float (&acc)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int a = 0; a < K; ++a) {
        const float* mrow = img_pad + (i + a) * W + j0;                 // padded coordinates: row i+a, col j0..
        const float* krow = sk + (FLIP ? (K - 1 - a) : a) * K;
        float w[K + 7];
#pragma unroll
        for (int t = 0; t < (K + 7) / 4; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(mrow + 4 * t);
            w[4 * t] = v.x; w[4 * t + 1] = v.y; w[4 * t + 2] = v.z; w[4 * t + 3] = v.w;
        }
#pragma unroll
        for (int t = ((K + 7) / 4) * 4; t < K + 7; ++t) w[t] = mrow[t];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const float kv = krow[FLIP ? (K - 1 - b) : b];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(w[b + q], kv, acc[q]);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256) likelihood_ctf_k(SvaeShape s, SvaeConfig c, int b0, const float* __restrict__ o,
                                                        const float* __restrict__ y, const float* __restrict__ ctf,
                                                        const uint8_t* __restrict__ mask, float* __restrict__ stats,
                                                        float* __restrict__ g_o) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[8];
    constexpr int PAD = K / 2;
    const int nr = s.n_rows, nc = s.n_cols, P = s.P;
    const int spr = (nc + 7) / 8;                        // strips per image row
    const int W = ((spr * 8 + 2 * PAD + 7) + 3) & ~3;    // padded row pitch (floats), 16-byte aligned rows
    const int Hh = nr + 2 * PAD;
    float* smu = sm;                                     // Hh x W, decoded mean, zero border
    float* sres = sm + Hh * W;                           // Hh x W, masked residual, zero border
    float* sk = sm + 2 * Hh * W;                         // K x K (rounded up to 4 floats)
    const int bl = blockIdx.x, b = b0 + bl;
    const float* ob = o + (long)bl * P;
    const float* yb = y + (long)b * P;
    const float* kb = ctf + (long)b * K * K;
    for (int i = threadIdx.x; i < 2 * Hh * W; i += blockDim.x) sm[i] = 0.f;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) sk[i] = kb[i];
    __syncthreads();
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        float mu, dmu;
        post_output(ob[j], 0, c.softplus, mu, dmu);
        smu[(j / nc + PAD) * W + (j % nc) + PAD] = mu;
    }
    __syncthreads();
    float ll = 0.f;
    for (int st = threadIdx.x; st < nr * spr; st += blockDim.x) {
        const int i = st / spr, j0 = (st % spr) * 8;
        float acc[8];
        ctf_strip<K, false>(smu, W, sk, i, j0, acc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (j < nc) {
                const int p = i * nc + j;
                const float m = mask ? (mask[p] ? 1.f : 0.f) : 1.f;
                const float r = m * (acc[q] - yb[p]);
                sres[(i + PAD) * W + j + PAD] = r;
                ll -= 0.5f * r * r;
            }
        }
    }
    __syncthreads();
    if (g_o != nullptr) {
        float* gb = g_o + (long)bl * P;
        const float gs = c.grad_scale;
        for (int st = threadIdx.x; st < nr * spr; st += blockDim.x) {
            const int i = st / spr, j0 = (st % spr) * 8;
            float acc[8];
            ctf_strip<K, true>(sres, W, sk, i, j0, acc);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = j0 + q;
                if (j < nc) {
                    const int p = i * nc + j;
                    float mu, dmu;
                    post_output(ob[p], 0, c.softplus, mu, dmu);
                    gb[p] = gs * acc[q] * dmu;
                }
            }
        }
    }
    const float tot = block_sum_256(ll, red);
    if (threadIdx.x == 0) stats[b * 3 + 0] = tot;
}

int likelihood(const SvaeShape& s, const SvaeConfig& c, int b0, int nb, const float* o, const float* y,
               const float* ctf, const uint8_t* mask, float* stats, float* g_o, cudaStream_t st) {
    size_t smem = 0;
    // register-tiled path for 39x39 kernels (the 40x40 particle configs): parity-green on a B200 and 18 % faster on
    // the whole C5 step (16.6 -> 13.6 ms) than the generic shared-memory correlation; SVAE_CTF_FAST=0 selects the
    // generic kernel (A/B comparisons; read once per process)
    static const bool ctf_fast = !(getenv("SVAE_CTF_FAST") != nullptr && getenv("SVAE_CTF_FAST")[0] == '0');
    if (ctf_fast && c.likelihood == SVAE_LIK_GAUSS && ctf != nullptr && s.k_ctf == 39) {
        const int spr = (s.n_cols + 7) / 8;
        const int W = ((spr * 8 + 2 * 19 + 7) + 3) & ~3;
        const size_t bytes = ((size_t)2 * (s.n_rows + 2 * 19) * W + 39 * 39 + 3) * sizeof(float);
        if (bytes <= 200 * 1024) {
            static bool configured = false;
            if (!configured) {
                SVAE_CUDA(cudaFuncSetAttribute(likelihood_ctf_k<39>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                configured = true;
            }
            likelihood_ctf_k<39><<<nb, 256, bytes, st>>>(s, c, b0, o, y, ctf, mask, stats, g_o);
            SVAE_LAUNCH_CHECK();
            return SVAE_OK;
        }
    }
    if (c.likelihood == SVAE_LIK_GAUSS && ctf != nullptr) {
        smem = ((size_t)2 * s.P + (size_t)s.k_ctf * s.k_ctf) * sizeof(float);
        if (smem > 48 * 1024) {
            SVAE_CUDA(cudaFuncSetAttribute(likelihood_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    likelihood_k<<<nb, 256, smem, st>>>(s, c, b0, o, y, ctf, mask, stats, g_o);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// output-layer backward: thread <-> column pair, loop over rows; no cross-thread reductions
// ------------------------------------------------------------------------------------------------
constexpr int OB_ROWS = 128;
constexpr int OB_MAXPAIRS = 4;   // fallback path: Hp <= 2048

// Fast path (Hp % 8 == 0, Hp <= 2048): thread = (8-column group g, row lane rl); a block covers OB_ROWS
// rows, each thread walks rows rl, rl+lanes, ... with 16-byte loads/stores and keeps its column sums in
// registers; row lanes are combined through shared memory, then one atomicAdd per column per block.
template <typename T, int C>
__global__ void __launch_bounds__(256) out_backward_v8_k(const T* __restrict__ h, const float* __restrict__ g_o,
                                                         int rows, int H, int Hp, int act,
                                                         const float* __restrict__ out_w, T* __restrict__ delta,
                                                         float* __restrict__ d_out_w, float* __restrict__ d_out_b,
                                                         float* __restrict__ d_b_last) {
    extern __shared__ float red[];            // lanes x (C+1) x Hp partial column sums
    __shared__ float sg[OB_ROWS][C];
    const long r0 = (long)blockIdx.x * OB_ROWS;
    const int nrows = (int)min((long)OB_ROWS, rows - r0);
    for (int i = threadIdx.x; i < nrows * C; i += blockDim.x) sg[i / C][i % C] = g_o[r0 * C + i];
    __syncthreads();
    const int groups = Hp >> 3;
    const int lanes = max(1, 256 / groups);
    const int passes = ceil_div(groups, 256);
    for (int ps = 0; ps < passes; ++ps) {
        const int g = ps * 256 + threadIdx.x % min(groups, 256);
        const int rl = threadIdx.x / groups;
        const bool active = (g < groups) && (rl < lanes);
        float w[C][8], aw[C][8], ab[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ab[e] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                aw[c][e] = 0.f;
                w[c][e] = (active && g * 8 + e < H) ? out_w[c * H + g * 8 + e] : 0.f;
            }
        }
        if (active) {
#pragma unroll 4
            for (int r = rl; r < nrows; r += lanes) {
                float hv[8], dv[8];
                load8(h + (r0 + r) * Hp + g * 8, hv);
                float gc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) gc[c] = sg[r][c];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float t = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        t = fmaf(gc[c], w[c][e], t);
                        aw[c][e] = fmaf(gc[c], hv[e], aw[c][e]);
                    }
                    dv[e] = t * act_deriv_from_out(act, hv[e]);
                    ab[e] += dv[e];
                }
                store8(delta + (r0 + r) * Hp + g * 8, dv);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
#pragma unroll
                for (int c = 0; c < C; ++c) red[(rl * (C + 1) + c) * Hp + g * 8 + e] = aw[c][e];
                red[(rl * (C + 1) + C) * Hp + g * 8 + e] = ab[e];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (C + 1) * Hp; i += blockDim.x) {
            const int c = i / Hp, n = i % Hp;
            if (n >= H || n / 8 / 256 != ps) continue;
            float t = 0.f;
            for (int l = 0; l < lanes; ++l) t += red[(l * (C + 1) + c) * Hp + n];
            if (c < C) atomicAdd(d_out_w + c * H + n, t);
            else if (d_b_last) atomicAdd(d_b_last + n, t);
        }
        __syncthreads();
    }
    if (threadIdx.x < C) {
        float t = 0.f;
        for (int r = 0; r < nrows; ++r) t += sg[r][threadIdx.x];
        atomicAdd(d_out_b + threadIdx.x, t);
    }
}

template <typename T, int C>
__global__ void __launch_bounds__(256) out_backward_k(const T* __restrict__ h, const float* __restrict__ g_o,
                                                      int rows, int H, int Hp, int act,
                                                      const float* __restrict__ out_w, T* __restrict__ delta,
                                                      float* __restrict__ d_out_w, float* __restrict__ d_out_b,
                                                      float* __restrict__ d_b_last) {
    __shared__ float sg[OB_ROWS][C];
    const long r0 = (long)blockIdx.x * OB_ROWS;
    const int nrows = (int)min((long)OB_ROWS, rows - r0);
    for (int i = threadIdx.x; i < nrows * C; i += blockDim.x) sg[i / C][i % C] = g_o[r0 * C + i];
    __syncthreads();
    const int npairs = Hp / 2;
    float w[OB_MAXPAIRS][2][C], aw[OB_MAXPAIRS][2][C], ab[OB_MAXPAIRS][2];
#pragma unroll
    for (int q = 0; q < OB_MAXPAIRS; ++q) {
        const int n = (threadIdx.x + q * 256) * 2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            ab[q][e] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                aw[q][e][c] = 0.f;
                w[q][e][c] = (n + e < H) ? out_w[c * H + n + e] : 0.f;
            }
        }
    }
    for (int r = 0; r < nrows; ++r) {
        float g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = sg[r][c];
#pragma unroll
        for (int q = 0; q < OB_MAXPAIRS; ++q) {
            const int pr = threadIdx.x + q * 256;
            if (pr < npairs) {
                float hv[2];
                load2(h + (r0 + r) * Hp + pr * 2, hv[0], hv[1]);
                float dv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float t = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        t = fmaf(g[c], w[q][e][c], t);
                        aw[q][e][c] = fmaf(g[c], hv[e], aw[q][e][c]);
                    }
                    dv[e] = t * act_deriv_from_out(act, hv[e]);
                    ab[q][e] += dv[e];
                }
                store2(delta + (r0 + r) * Hp + pr * 2, dv[0], dv[1]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < OB_MAXPAIRS; ++q) {
        const int n = (threadIdx.x + q * 256) * 2;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (n + e < H) {
#pragma unroll
                for (int c = 0; c < C; ++c) atomicAdd(d_out_w + c * H + n + e, aw[q][e][c]);
                if (d_b_last) atomicAdd(d_b_last + n + e, ab[q][e]);
            }
        }
    }
    if (threadIdx.x < C) {
        float t = 0.f;
        for (int r = 0; r < nrows; ++r) t += sg[r][threadIdx.x];
        atomicAdd(d_out_b + threadIdx.x, t);
    }
}

template <typename T>
int out_backward(const T* h, const float* g_o, int rows, int H, int Hp, int C, int act, const float* out_w,
                 T* delta, float* d_out_w, float* d_out_b, float* d_b_last, cudaStream_t st) {
    SVAE_REQUIRE(Hp <= 2 * 256 * OB_MAXPAIRS, SVAE_EINVAL, "hidden width %d too large", Hp);
    const int blocks = ceil_div(rows, OB_ROWS);
    if ((Hp & 7) == 0) {
        const int groups = Hp >> 3;
        const int lanes = groups >= 256 ? 1 : 256 / groups;
        const size_t smem = (size_t)lanes * (C + 1) * Hp * sizeof(float);
        if (smem <= 40 * 1024) {
            switch (C) {
                case 1: out_backward_v8_k<T, 1><<<blocks, 256, smem, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
                case 2: out_backward_v8_k<T, 2><<<blocks, 256, smem, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
                case 3: out_backward_v8_k<T, 3><<<blocks, 256, smem, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
                case 4: out_backward_v8_k<T, 4><<<blocks, 256, smem, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
                default: set_error("n_out=%d not supported (1..4)", C); return SVAE_EINVAL;
            }
            SVAE_LAUNCH_CHECK();
            return SVAE_OK;
        }
    }
    switch (C) {
        case 1: out_backward_k<T, 1><<<blocks, 256, 0, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
        case 2: out_backward_k<T, 2><<<blocks, 256, 0, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
        case 3: out_backward_k<T, 3><<<blocks, 256, 0, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
        case 4: out_backward_k<T, 4><<<blocks, 256, 0, st>>>(h, g_o, rows, H, Hp, act, out_w, delta, d_out_w, d_out_b, d_b_last); break;
        default: set_error("n_out=%d not supported (1..4)", C); return SVAE_EINVAL;
    }
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int out_backward<float>(const float*, const float*, int, int, int, int, int, const float*, float*, float*, float*, float*, cudaStream_t);
template int out_backward<__nv_bfloat16>(const __nv_bfloat16*, const float*, int, int, int, int, int, const float*, __nv_bfloat16*, float*, float*, float*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients) and per-image coordinate moments of delta0
// ------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 256;

template <typename T>
__global__ void __launch_bounds__(256) col_sum_k(const T* __restrict__ src, int rows, int H, int Hp,
                                                 float* __restrict__ dst, int rows_per_block) {
    const long r0 = (long)blockIdx.x * rows_per_block;
    const int nrows = (int)min((long)rows_per_block, rows - r0);
    if (Hp & 1) {   // odd row stride: no aligned pairs
        for (int n = threadIdx.x; n < H; n += blockDim.x) {
            float a = 0.f;
            for (int r = 0; r < nrows; ++r) a += to_f32(src[(r0 + r) * Hp + n]);
            atomicAdd(dst + n, a);
        }
        return;
    }
    for (int pr = threadIdx.x; pr < Hp / 2; pr += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int r = 0; r < nrows; ++r) {
            float x, y;
            load2(src + (r0 + r) * Hp + pr * 2, x, y);
            a0 += x; a1 += y;
        }
        if (pr * 2 < H) atomicAdd(dst + pr * 2, a0);
        if (pr * 2 + 1 < H) atomicAdd(dst + pr * 2 + 1, a1);
    }
}
// Wide variant (Hp % 8 == 0, Hp >= 256): thread = (8-column group, row lane), 16-byte loads, four rows in flight per
// thread, 256 column groups per pass.  The 4-byte-per-thread kernel above ran at 2.2 TB/s on the (524288 x 1024) bf16 delta matrices of C4
// (490 us each, two per step).
template <typename T>
__global__ void __launch_bounds__(256) col_sum_v8_k(const T* __restrict__ src, int rows, int H, int Hp,
                                                    float* __restrict__ dst, int rows_per_block) {
    __shared__ float red[256 * 8];                     // [row lane][column group][8] partial sums of one pass
    const long r0 = (long)blockIdx.x * rows_per_block;
    const int nrows = (int)min((long)rows_per_block, rows - r0);
    const int groups = Hp >> 3;
    const int gpp = min(groups, 256);                  // column groups handled per pass
    const int lanes = 256 / gpp;                       // row lanes (1, 2, 4 or 8 for Hp >= 256)
    for (int g0 = 0; g0 < groups; g0 += gpp) {
        const int g = g0 + threadIdx.x % gpp, rl = threadIdx.x / gpp;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        if (g < groups && rl < lanes) {
            const T* base = src + r0 * Hp + g * 8;
            int r = rl;
            for (; r + 3 * lanes < nrows; r += 4 * lanes) {
                float v0[8], v1[8], v2[8], v3[8];
                load8(base + (long)r * Hp, v0);
                load8(base + (long)(r + lanes) * Hp, v1);
                load8(base + (long)(r + 2 * lanes) * Hp, v2);
                load8(base + (long)(r + 3 * lanes) * Hp, v3);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]);
            }
            for (; r < nrows; r += lanes) {
                float v0[8];
                load8(base + (long)r * Hp, v0);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += v0[e];
            }
        }
        if (lanes > 1) {
            __syncthreads();
            if (g < groups && rl < lanes && rl > 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) red[(rl * gpp + threadIdx.x % gpp) * 8 + e] = acc[e];
            }
            __syncthreads();
            if (g < groups && rl == 0) {
                for (int l = 1; l < lanes; ++l)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += red[(l * gpp + threadIdx.x % gpp) * 8 + e];
            }
        }
        if (g < groups && rl == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (g * 8 + e < H) atomicAdd(dst + g * 8 + e, acc[e]);
        }
    }
}

template <typename T>
int col_sum(const T* src, int rows, int H, int Hp, float* dst, cudaStream_t st) {
    if ((Hp & 7) == 0 && Hp >= 256 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int rpb = rows >= 148 * 8 * 512 ? 512 : max(32, ceil_div(rows, 148 * 4));
        col_sum_v8_k<T><<<ceil_div(rows, rpb), 256, 0, st>>>(src, rows, H, Hp, dst, rpb);
        SVAE_LAUNCH_CHECK();
        return SVAE_OK;
    }
    const int rpb = rows >= 148 * 4 * CS_ROWS ? CS_ROWS : max(8, ceil_div(rows, 148 * 2));
    col_sum_k<T><<<ceil_div(rows, rpb), 256, 0, st>>>(src, rows, H, Hp, dst, rpb);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int col_sum<float>(const float*, int, int, int, float*, cudaStream_t);
template int col_sum<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, float*, cudaStream_t);

template <typename T>
__global__ void __launch_bounds__(256) image_col_reduce_k(const T* __restrict__ d0, int b0, int P, int Hp,
                                                          const float* __restrict__ grid,
                                                          const float* __restrict__ xe, float* __restrict__ S) {
    const int bl = blockIdx.x, b = b0 + bl;
    const float* coords = xe ? xe + (long)b * P * 2 : grid;
    const T* base = d0 + (long)bl * P * Hp;
    float* Sb = S + (long)b * 3 * Hp;
    for (int pr = threadIdx.x + blockIdx.y * blockDim.x; pr < Hp / 2; pr += blockDim.x * gridDim.y) {
        float s0 = 0.f, s1 = 0.f, m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f;
        for (int p = 0; p < P; ++p) {
            float x, y;
            load2(base + (long)p * Hp + pr * 2, x, y);
            const float c0 = __ldg(coords + p * 2), c1 = __ldg(coords + p * 2 + 1);
            s0 += x; s1 += y;
            m00 = fmaf(c0, x, m00); m01 = fmaf(c0, y, m01);
            m10 = fmaf(c1, x, m10); m11 = fmaf(c1, y, m11);
        }
        Sb[pr * 2] = s0; Sb[pr * 2 + 1] = s1;
        Sb[Hp + pr * 2] = m00; Sb[Hp + pr * 2 + 1] = m01;
        Sb[2 * Hp + pr * 2] = m10; Sb[2 * Hp + pr * 2 + 1] = m11;
    }
}

// Wide variant (Hp % 8 == 0, Hp <= 2048): thread = (8-column group, row lane), 16-byte loads, several rows in
// flight per thread; row lanes are combined through shared memory.  One block per image.
template <typename T>
__global__ void __launch_bounds__(256) image_col_reduce_v8_k(const T* __restrict__ d0, int b0, int P, int Hp,
                                                             const float* __restrict__ grid,
                                                             const float* __restrict__ xe, float* __restrict__ S) {
    extern __shared__ float red[];            // lanes x 3 x Hp
    const int bl = blockIdx.x, b = b0 + bl;
    const float* coords = xe ? xe + (long)b * P * 2 : grid;
    const T* base = d0 + (long)bl * P * Hp;
    const int groups = Hp >> 3;
    const int lanes = max(1, 256 / groups);
    const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
    float s[8], m0[8], m1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; m0[e] = 0.f; m1[e] = 0.f; }
    if (rl < lanes) {
#pragma unroll 4
        for (int p = rl; p < P; p += lanes) {
            float v[8];
            load8(base + (long)p * Hp + g * 8, v);
            const float2 c = __ldg(reinterpret_cast<const float2*>(coords) + p);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s[e] += v[e];
                m0[e] = fmaf(c.x, v[e], m0[e]);
                m1[e] = fmaf(c.y, v[e], m1[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            red[(rl * 3 + 0) * Hp + g * 8 + e] = s[e];
            red[(rl * 3 + 1) * Hp + g * 8 + e] = m0[e];
            red[(rl * 3 + 2) * Hp + g * 8 + e] = m1[e];
        }
    }
    __syncthreads();
    float* Sb = S + (long)b * 3 * Hp;
    for (int i = threadIdx.x; i < 3 * Hp; i += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[l * 3 * Hp + i];
        Sb[i] = t;
    }
}

template <typename T>
int image_col_reduce(const T* delta0, int b0, int nb, int P, int Hp, const float* grid, const float* x_explicit,
                     float* S, cudaStream_t st) {
    if ((Hp & 7) == 0 && Hp <= 2048) {
        const int groups = Hp >> 3;
        const int lanes = groups >= 256 ? 1 : 256 / groups;
        const size_t smem = (size_t)lanes * 3 * Hp * sizeof(float);
        if (smem <= 48 * 1024) {
            image_col_reduce_v8_k<T><<<nb, 256, smem, st>>>(delta0, b0, P, Hp, grid, x_explicit, S);
            SVAE_LAUNCH_CHECK();
            return SVAE_OK;
        }
    }
    dim3 g(nb, ceil_div(Hp / 2, 256));
    image_col_reduce_k<T><<<g, 256, 0, st>>>(delta0, b0, P, Hp, grid, x_explicit, S);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int image_col_reduce<float>(const float*, int, int, int, int, const float*, const float*, float*, cudaStream_t);
template int image_col_reduce<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, const float*, const float*, float*, cudaStream_t);

template <typename T>
__global__ void __launch_bounds__(256) coord_row_grad_k(const T* __restrict__ d0, int rows, int H, int Hp,
                                                        const float* __restrict__ coord_w, float* __restrict__ gx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long r = (long)blockIdx.x * 8 + warp; r < rows; r += (long)gridDim.x * 8) {
        float a0 = 0.f, a1 = 0.f;
        for (int n = lane * 2; n < Hp; n += 64) {
            float x, y;
            load2(d0 + r * Hp + n, x, y);
            if (n < H) { a0 = fmaf(x, __ldg(coord_w + n * 2), a0); a1 = fmaf(x, __ldg(coord_w + n * 2 + 1), a1); }
            if (n + 1 < H) { a0 = fmaf(y, __ldg(coord_w + n * 2 + 2), a0); a1 = fmaf(y, __ldg(coord_w + n * 2 + 3), a1); }
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1);
        if (lane == 0) { gx[r * 2] = a0; gx[r * 2 + 1] = a1; }
    }
}
template <typename T>
int coord_row_grad(const T* delta0, int rows, int H, int Hp, const float* coord_w, float* g_x, cudaStream_t st) {
    coord_row_grad_k<T><<<min(ceil_div(rows, 8), 148 * 8), 256, 0, st>>>(delta0, rows, H, Hp, coord_w, g_x);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}
template int coord_row_grad<float>(const float*, int, int, int, const float*, float*, cudaStream_t);
template int coord_row_grad<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, const float*, float*, cudaStream_t);

// dWc / dbc from the per-image sums
__global__ void coord_param_grad_k(const float* __restrict__ S, const float* __restrict__ img, int B, int H, int Hp,
                                   int explicit_x, float* __restrict__ dw, float* __restrict__ db) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= H) return;
    const int per = ceil_div(B, gridDim.y);
    const int bs = blockIdx.y * per, be = min(B, bs + per);
    float a0 = 0.f, a1 = 0.f, ab = 0.f;
    for (int b = bs; b < be; ++b) {
        const float* Sb = S + (long)b * 3 * Hp;
        const float s = Sb[n], m0 = Sb[Hp + n], m1 = Sb[2 * Hp + n];
        ab += s;
        if (explicit_x) { a0 += m0; a1 += m1; }
        else {
            const float cs = img[b * 4], sn = img[b * 4 + 1], dx0 = img[b * 4 + 2], dx1 = img[b * 4 + 3];
            a0 += cs * m0 - sn * m1 + dx0 * s;
            a1 += sn * m0 + cs * m1 + dx1 * s;
        }
    }
    atomicAdd(dw + n * 2, a0);
    atomicAdd(dw + n * 2 + 1, a1);
    atomicAdd(db + n, ab);
}
int coord_param_grad(const float* S, const float* img, int B, int H, int Hp, int explicit_x, float* d_coord_w,
                     float* d_coord_b, cudaStream_t st) {
    dim3 g(ceil_div(H, 128), min(ceil_div(B, 32), 64));
    coord_param_grad_k<<<g, 128, 0, st>>>(S, img, B, H, Hp, explicit_x, d_coord_w, d_coord_b);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// per-image d(theta), d(dx) and the chain into the encoder head; one warp per image
__global__ void __launch_bounds__(256) latent_backward_k(SvaeShape s, SvaeConfig c, const float* __restrict__ S,
                                                         int Hp, const float* __restrict__ img,
                                                         const float* __restrict__ coord_w,
                                                         const float* __restrict__ dz, const float* __restrict__ zo,
                                                         const float* __restrict__ eps, float* __restrict__ g_zo,
                                                         const float* __restrict__ coord_pre) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= s.B) return;
    float dth = 0.f, d0 = 0.f, d1 = 0.f;
    if (coord_pre != nullptr) {     // option path: already reduced by latent_coord_grad_k
        dth = coord_pre[b * 3]; d0 = coord_pre[b * 3 + 1]; d1 = coord_pre[b * 3 + 2];
    } else {
        const float* Sb = S + (long)b * 3 * Hp;
        const float cs = img[b * 4], sn = img[b * 4 + 1];
        for (int n = lane; n < s.H; n += 32) {
            const float w0 = coord_w[n * 2], w1 = coord_w[n * 2 + 1];
            const float sv = Sb[n], m0 = Sb[Hp + n], m1 = Sb[2 * Hp + n];
            dth = fmaf(w0, -sn * m0 - cs * m1, fmaf(w1, cs * m0 - sn * m1, dth));
            d0 = fmaf(w0, sv, d0);
            d1 = fmaf(w1, sv, d1);
        }
        dth = warp_sum(dth); d0 = warp_sum(d0); d1 = warp_sum(d1);
    }
    const int I = s.I, rot = c.rotate ? 1 : 0, zcol = rot + (c.translate ? 2 : 0);
    const float gs = c.grad_scale;
    for (int i = lane; i < I; i += 32) {
        const float mu = zo[(long)b * 2 * I + i], ls = zo[(long)b * 2 * I + I + i];
        const float sd = expf(ls);
        float gl, kmu, kls;
        if (rot && i == 0) {
            gl = dth;
            const float sp2 = c.theta_prior * c.theta_prior;
            kmu = c.theta_kl_mean ? mu / sp2 : 0.f;
            kls = -1.f + sd * sd / sp2;
        } else {
            if (i < zcol) gl = ((i - rot) == 0 ? d0 : d1) * c.dx_scale;
            else gl = dz ? dz[(long)b * s.Z + (i - zcol)] : 0.f;
            kmu = mu;
            kls = -1.f + sd * sd;
        }
        g_zo[(long)b * 2 * I + i] = gl + gs * kmu;
        g_zo[(long)b * 2 * I + I + i] = gl * sd * eps[(long)b * I + i] + gs * kls;
    }
}
int latent_backward(const SvaeShape& s, const SvaeConfig& c, const float* S, int Hp, const float* img,
                    const float* coord_w, const float* dz, const float* zo, const float* eps, float* g_zo,
                    cudaStream_t st, const float* coord_pre) {
    latent_backward_k<<<ceil_div(s.B, 8), 256, 0, st>>>(s, c, S, Hp, img, coord_w, dz, zo, eps, g_zo, coord_pre);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, train_mnist.py:389-392,149-150), gather, fp32 -> padded bf16
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float lr, float b1, float b2, float eps,
                                         float bc1, float bc2_sqrt, int zero_grad) {
    const float gi = g;
    const float mi = b1 * m + (1.f - b1) * gi;
    const float vi = b2 * v + (1.f - b2) * gi * gi;
    m = mi; v = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p -= (lr / bc1) * (mi / denom);
    if (zero_grad) g = 0.f;
}
// 28 bytes of traffic per parameter: four elements per thread with 16-byte accesses (the flat buffers are 16-byte
// aligned; the scalar version ran at 3.9 TB/s on the 89.7 M parameters of C4), scalar tail
__global__ void __launch_bounds__(256) adam_k(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                              float bc1, float bc2_sqrt, int zero_grad, const float* __restrict__ bc_dev,
                                              int vec) {
    if (bc_dev != nullptr) { bc1 = bc_dev[0]; bc2_sqrt = bc_dev[1]; }   // CUDA-graph replays: step-dependent scalars live in memory
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t i = tid; i < n4; i += nthreads) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, lr, b1, b2, eps, bc1, bc2_sqrt, zero_grad);
        adam_one(pp.y, gg.y, mm.y, vv.y, lr, b1, b2, eps, bc1, bc2_sqrt, zero_grad);
        adam_one(pp.z, gg.z, mm.z, vv.z, lr, b1, b2, eps, bc1, bc2_sqrt, zero_grad);
        adam_one(pp.w, gg.w, mm.w, vv.w, lr, b1, b2, eps, bc1, bc2_sqrt, zero_grad);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = gg;
    }
    for (size_t i = 4 * n4 + tid; i < n; i += nthreads)
        adam_one(p[i], g[i], m[i], v[i], lr, b1, b2, eps, bc1, bc2_sqrt, zero_grad);
}
// device-resident step counter: t += 1, then the bias corrections of step t (double precision, like the host path)
__global__ void adam_tick_k(int* t_dev, float* bc, float b1, float b2) {
    const int t = ++(*t_dev);
    bc[0] = (float)(1.0 - pow((double)b1, (double)t));
    bc[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}
int adam_tick(int* t_dev, float* bias_corr_dev, float b1, float b2, cudaStream_t st) {
    adam_tick_k<<<1, 1, 0, st>>>(t_dev, bias_corr_dev, b1, b2);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

int adam(float* p, float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, int t,
         int zero_grad, const float* bias_corr_dev, cudaStream_t st) {
    if (n == 0) return SVAE_OK;
    const double bc1 = 1.0 - pow((double)b1, t > 0 ? t : 1), bc2 = 1.0 - pow((double)b2, t > 0 ? t : 1);
    const int blocks = (int)min((size_t)148 * 16, (n / 4 + 255) / 256 + 1);
    const int vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    adam_k<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), zero_grad, bias_corr_dev, vec);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

__global__ void gather_rows_k(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst,
                              int64_t n_rows, int64_t row_len) {
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const float* s = src + idx[r] * row_len;
        float* d = dst + r * row_len;
        for (int64_t i = threadIdx.x; i < row_len; i += blockDim.x) d[i] = __ldg(s + i);
    }
}
int gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n_rows, int64_t row_len, cudaStream_t st) {
    if (n_rows == 0) return SVAE_OK;
    gather_rows_k<<<(int)min((int64_t)148 * 8, n_rows), 256, 0, st>>>(src, idx, dst, n_rows, row_len);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

__global__ void to_bf16_padded_k(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* __restrict__ dst,
                                 int rows_p, int cols_p) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows_p * cols_p) return;
    const int r = (int)(i / cols_p), c = (int)(i % cols_p);
    dst[i] = __float2bfloat16_rn((r < rows && c < cols) ? src[(long)r * cols + c] : 0.f);
}
int to_bf16_padded(const float* src, int rows, int cols, __nv_bfloat16* dst, int rows_p, int cols_p, cudaStream_t st) {
    to_bf16_padded_k<<<ceil_div((long)rows_p * cols_p, 256), 256, 0, st>>>(src, rows, cols, dst, rows_p, cols_p);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// Bicubic rotation of a minibatch about the image centre, bit-compatible with Pillow's
// Image.rotate(angle, resample=BICUBIC) as the reference uses it for --augment-rotation
// (train_particles.py:39-43 on float32 images, train_galaxy.py:47-54 through uint8).  Pillow resamples in
// double precision with the a = -1 cubic: 4x4 window with clamped columns, rows outside the image repeat the
// previous row's value, destination pixels that map outside the source are 0; for float images the
// coefficient sums of the horizontal pass are float32 (C float arithmetic), integer samples are exact.
// Explicit _rn intrinsics keep nvcc from contracting a*b+c into FMAs, which would change the rounding.
// mode: 0 general (inverse affine in mat[b*6..]), 1 copy, 2 rot180, 3 rot90 (ccw), 4 rot270.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double horner3(double p1, double p2, double p3, double p4, double d) {
    return __dadd_rn(p1, __dmul_rn(d, __dadd_rn(p2, __dmul_rn(d, __dadd_rn(p3, __dmul_rn(d, p4))))));
}
__device__ __forceinline__ double cubic_f32(float v1, float v2, float v3, float v4, double d) {
    const float p2 = __fadd_rn(-v1, v3);
    const float p3 = __fsub_rn(__fadd_rn(__fmul_rn(2.f, __fsub_rn(v1, v2)), v3), v4);
    const float p4 = __fadd_rn(__fsub_rn(__fadd_rn(-v1, v2), v3), v4);
    return horner3((double)v2, (double)p2, (double)p3, (double)p4, d);
}
__device__ __forceinline__ double cubic_f64(double v1, double v2, double v3, double v4, double d) {
    const double p2 = __dadd_rn(-v1, v3);
    const double p3 = __dsub_rn(__dadd_rn(__dmul_rn(2.0, __dsub_rn(v1, v2)), v3), v4);
    const double p4 = __dadd_rn(__dsub_rn(__dadd_rn(-v1, v2), v3), v4);
    return horner3(v2, p2, p3, p4, d);
}

template <bool U8>
__global__ void rotate_bicubic_k(const float* __restrict__ src, float* __restrict__ dst, const double* __restrict__ mat,
                                 const int* __restrict__ mode, int B, int h, int w, int C) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)B * h * w) return;
    const int b = (int)(idx / (h * w));
    const int pix = (int)(idx % (h * w));
    const int i = pix / w, j = pix % w;
    const float* im = src + (long)b * h * w * C;
    float* o = dst + ((long)b * h * w + pix) * C;
    const int md = mode[b];
    auto sample = [&](int r, int c, int ch) -> float {      // the value Pillow sees at (row r, col c)
        const float v = im[((long)r * w + c) * C + ch];
        return U8 ? (float)(unsigned char)__fmul_rn(v, 255.f) : v;
    };
    auto emit = [&](int ch, double v, bool inside) {
        if (U8) {
            const double q = !inside ? 0.0 : (v <= 0.0 ? 0.0 : (v >= 255.0 ? 255.0 : floor(v)));
            o[ch] = (float)__ddiv_rn(q, 255.0);
        } else {
            o[ch] = inside ? (float)v : 0.f;
        }
    };
    if (md != 0) {
        int r = i, c = j;
        if (md == 2) { r = h - 1 - i; c = w - 1 - j; }
        else if (md == 3) { r = j; c = w - 1 - i; }
        else if (md == 4) { r = h - 1 - j; c = i; }
        for (int ch = 0; ch < C; ++ch) emit(ch, (double)sample(r, c, ch), true);
        return;
    }
    const double* m = mat + (long)b * 6;
    const double xs = j + 0.5, ys = i + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xs), __dmul_rn(m[1], ys)), m[2]);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xs), __dmul_rn(m[4], ys)), m[5]);
    const bool inside = (xin >= 0.0) && (xin < (double)w) && (yin >= 0.0) && (yin < (double)h);
    if (!inside) {
        for (int ch = 0; ch < C; ++ch) emit(ch, 0.0, false);
        return;
    }
    xin -= 0.5; yin -= 0.5;
    int x = (int)floor(xin), y = (int)floor(yin);
    const double dx = xin - x, dy = yin - y;
    --x; --y;
    int cc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cc[k] = min(max(x + k, 0), w - 1);
    for (int ch = 0; ch < C; ++ch) {
        double rowv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = y + k;
            if (k == 0 || (yy >= 0 && yy < h)) {
                const int r = min(max(yy, 0), h - 1);
                const float a0 = sample(r, cc[0], ch), a1 = sample(r, cc[1], ch), a2 = sample(r, cc[2], ch),
                            a3 = sample(r, cc[3], ch);
                rowv[k] = U8 ? cubic_f64(a0, a1, a2, a3, dx) : cubic_f32(a0, a1, a2, a3, dx);
            } else {
                rowv[k] = rowv[k - 1];
            }
        }
        emit(ch, cubic_f64(rowv[0], rowv[1], rowv[2], rowv[3], dy), true);
    }
}
int rotate_bicubic(const float* src, float* dst, const double* mat, const int* mode, int B, int h, int w, int C,
                   int quantize_u8, cudaStream_t st) {
    const long n = (long)B * h * w;
    if (n == 0) return SVAE_OK;
    if (quantize_u8) rotate_bicubic_k<true><<<ceil_div(n, 256), 256, 0, st>>>(src, dst, mat, mode, B, h, w, C);
    else rotate_bicubic_k<false><<<ceil_div(n, 256), 256, 0, st>>>(src, dst, mat, mode, B, h, w, C);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// SM clock probe: one thread spins ~20 us and reports cycles / wall time.  bench.py enqueues it between steps
// to sample the clock under load without NVML (whose queries were measured to stall NCCL steps).
// ------------------------------------------------------------------------------------------------
__global__ void clock_probe_k(float* out_mhz) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 20000ull);
    const long long c1 = clock64();
    *out_mhz = (float)((double)(c1 - c0) * 1e3 / (double)(t1 - t0));
}
int clock_probe(float* out_mhz, cudaStream_t st) {
    clock_probe_k<<<1, 1, 0, st>>>(out_mhz);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> three bf16 terms for error-compensated tensor-core GEMMs (encoder):
//   x = hi + lo (+ 2^-17 |x|),  A*B ~ Ahi*Bhi + Ahi*Blo + Alo*Bhi
// pattern 0 ("A side") emits (hi, hi, lo), pattern 1 ("B side") emits (hi, lo, hi); the three terms are laid
// side by side along K: kcat = 1 -> dst (rows_p, 3*cols_p) with term t at columns [t*cols_p, ...),
// kcat = 0 -> dst (3*rows_p, cols_p) with term t at rows [t*rows_p, ...).  Padding is written as zeros.
// ------------------------------------------------------------------------------------------------
__global__ void split3_k(const float* __restrict__ src, int rows, int cols, long ld, __nv_bfloat16* __restrict__ dst,
                         int rows_p, int cols_p, int kcat, int pattern) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows_p * cols_p) return;
    const int r = (int)(i / cols_p), c = (int)(i % cols_p);
    const float x = (r < rows && c < cols) ? src[(long)r * ld + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const __nv_bfloat16 t1 = pattern == 0 ? hi : lo, t2 = pattern == 0 ? lo : hi;
    if (kcat) {
        __nv_bfloat16* d = dst + (long)r * 3 * cols_p + c;
        d[0] = hi; d[cols_p] = t1; d[2 * cols_p] = t2;
    } else {
        __nv_bfloat16* d = dst + (long)r * cols_p + c;
        const long seg = (long)rows_p * cols_p;
        d[0] = hi; d[seg] = t1; d[2 * seg] = t2;
    }
}
// four columns per thread: one 16-byte load (when the source row allows it) and three 8-byte stores.  cols_p is a
// multiple of 64 in every caller, so the destination is always 8-byte aligned.
// FINISH: src holds the RAW sums of a K-split GEMM (ld = cols_p, 16-byte aligned); h = act(src + bias) goes back in
// place (padded columns: act(0), what the fused GEMM epilogue writes there) before it is split; dst may be NULL.
template <bool FINISH>
__global__ void __launch_bounds__(256) split3_v4_k(const float* __restrict__ src, int rows, int cols, long ld,
                                                   __nv_bfloat16* __restrict__ dst, int rows_p, int cols_p, int kcat,
                                                   int pattern, int src_vec, const float* __restrict__ bias, int act,
                                                   float* act_out, __nv_bfloat16* __restrict__ dst_alt,
                                                   int pattern_alt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = cols_p >> 2;
    if (i >= (long)rows_p * c4) return;
    const int r = (int)(i / c4), c = (int)(i % c4) * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
        const float* sp = src + (long)r * ld + c;
        if (FINISH) {
            const float4 v = *reinterpret_cast<const float4*>(sp);
            const float raw[4] = {v.x, v.y, v.z, v.w};
            float h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool live = c + e < cols;
                h[e] = act_apply<false>(act, live ? raw[e] + (bias != nullptr ? __ldg(bias + c + e) : 0.f) : 0.f);
                x[e] = live ? h[e] : 0.f;
            }
            *reinterpret_cast<float4*>(act_out + (long)r * ld + c) = make_float4(h[0], h[1], h[2], h[3]);
        } else if (src_vec && c + 3 < cols) {
            const float4 v = *reinterpret_cast<const float4*>(sp);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (c + e < cols) ? sp[e] : 0.f;
        }
    }
    if (FINISH && dst == nullptr) return;
    uint32_t hi2[2], lo2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * e] - __low2float(h), x[2 * e + 1] - __high2float(h));
        hi2[e] = *reinterpret_cast<const uint32_t*>(&h);
        lo2[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint2 hi = make_uint2(hi2[0], hi2[1]), lo = make_uint2(lo2[0], lo2[1]);
    // dst in the requested layout; dst_alt (optional) receives the terms of pattern_alt in the other one
    __nv_bfloat16* dk = kcat ? dst : dst_alt;
    __nv_bfloat16* dr = kcat ? dst_alt : dst;
    const int pk = kcat ? pattern : pattern_alt, pr = kcat ? pattern_alt : pattern;
    if (dk != nullptr) {
        __nv_bfloat16* d = dk + (long)r * 3 * cols_p + c;
        *reinterpret_cast<uint2*>(d) = hi;
        *reinterpret_cast<uint2*>(d + cols_p) = pk == 0 ? hi : lo;
        *reinterpret_cast<uint2*>(d + 2 * cols_p) = pk == 0 ? lo : hi;
    }
    if (dr != nullptr) {
        __nv_bfloat16* d = dr + (long)r * cols_p + c;
        const long seg = (long)rows_p * cols_p;
        *reinterpret_cast<uint2*>(d) = hi;
        *reinterpret_cast<uint2*>(d + seg) = pr == 0 ? hi : lo;
        *reinterpret_cast<uint2*>(d + 2 * seg) = pr == 0 ? lo : hi;
    }
}
__global__ void act_backward_k(const float* __restrict__ out, const float* __restrict__ g, float* __restrict__ g_pre,
                               long n, int act) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g_pre[i] = g[i] * act_deriv_from_out(act, out[i]);
}
int act_backward(const float* out, const float* g, float* g_pre, long n, int act, cudaStream_t st) {
    act_backward_k<<<ceil_div(n, 256), 256, 0, st>>>(out, g, g_pre, n, act);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

int split3(const float* src, int rows, int cols, long ld, __nv_bfloat16* dst, int rows_p, int cols_p, int kcat,
           int pattern, cudaStream_t st) {
    if ((cols_p & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        const int src_vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ld % 4 == 0);
        split3_v4_k<false><<<ceil_div((long)rows_p * (cols_p >> 2), 256), 256, 0, st>>>(src, rows, cols, ld, dst, rows_p, cols_p, kcat,
                                                                                        pattern, src_vec, nullptr, -1, nullptr, nullptr, pattern);
        SVAE_LAUNCH_CHECK();
        return SVAE_OK;
    }
    split3_k<<<ceil_div((long)rows_p * cols_p, 256), 256, 0, st>>>(src, rows, cols, ld, dst, rows_p, cols_p, kcat, pattern);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

int split3_act(float* raw, const float* bias, int act, int rows, int cols, __nv_bfloat16* dst, int rows_p, int cols_p,
               int kcat, int pattern, cudaStream_t st) {
    SVAE_REQUIRE((cols_p & 3) == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 &&
                 (dst == nullptr || (reinterpret_cast<uintptr_t>(dst) & 7) == 0) && rows_p >= rows, SVAE_EALIGN,
                 "split3_act: the raw sums must be 16-byte aligned rows of a multiple of 4 columns");
    // without a split destination only the live rows are finished
    const int rp = dst != nullptr ? rows_p : rows;
    split3_v4_k<true><<<ceil_div((long)rp * (cols_p >> 2), 256), 256, 0, st>>>(raw, rows, cols, cols_p, dst, rp, cols_p, kcat,
                                                                              pattern, 1, bias, act, raw, nullptr, pattern);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

// one pass over src, both layouts: dst_r (3*rows_p, cols_p) row-stacked with the terms of pattern_r and dst_k
// (rows_p, 3*cols_p) K-concatenated with those of pattern_k
int split3_both(const float* src, int rows, int cols, long ld, __nv_bfloat16* dst_r, int pattern_r, __nv_bfloat16* dst_k,
                int pattern_k, int rows_p, int cols_p, cudaStream_t st) {
    SVAE_REQUIRE((cols_p & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_r) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(dst_k) & 7) == 0, SVAE_EALIGN, "split3_both: unaligned destination");
    const int src_vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ld % 4 == 0);
    split3_v4_k<false><<<ceil_div((long)rows_p * (cols_p >> 2), 256), 256, 0, st>>>(src, rows, cols, ld, dst_r, rows_p, cols_p, 0,
                                                                                    pattern_r, src_vec, nullptr, -1, nullptr, dst_k,
                                                                                    pattern_k);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

}  // namespace svae
