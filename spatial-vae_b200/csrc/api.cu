// C-ABI entry points of libsvae_b200.so and the host-side orchestration of one step.
// See include/svae_b200.h for the contract and the reference call sites each entry replaces.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "kernels.cuh"

namespace svae {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return SVAE_ECUDA;
}

// ---- workspace plan ------------------------------------------------------------------------------
struct Plan {
    int Hp = 0;            // padded hidden width (row stride of the activation matrices)
    int chunk = 0;         // images per decoder pass
    size_t esize = 4;      // bytes per activation element
    size_t total = 0;
    // offsets (bytes)
    size_t enc_acts, zo, enc_scratch, lat, eps, img, zs, hz, S, dz, g_zo, o, g_o, acts, delta, wbf16;
    size_t act_stride = 0, delta_stride = 0, w_stride = 0;
    // encoder on tensor cores (FAST): leading dimension of the fp32 activations and the bf16 split buffers
    int enc_ld = 0;
    size_t xs_k, gs_r, as_r, ws_k[SVAE_MAX_LAYERS], ws_r[SVAE_MAX_LAYERS];
    // decoder options (models.py:65-67,74-75): F coordinate features, K1 per-image moment rows in S,
    // opt = first layer runs through option_kernels.cu; weff (B, H*F) per-image coordinate weights (bilinear),
    // dlat (B,3) = (d theta, d t0, d t1); enc_wres (Hq,Hq) fp32 W + I for the encoder (resid on tensor cores)
    int F = 2, K1 = 3;
    bool opt = false;
    size_t weff = 0, dlat = 0, enc_wres = 0;
    bool resid_tc = false;          // ResidLinear layers on the tensor-core GEMMs: wbf16 also holds the W + I copies
    // PARITY_TC: bf16 split buffers of the decoder's 3-term GEMMs (K-concatenated rows x 3Hp; row-stacked 3rows x Hp)
    bool tc3 = false;
    size_t d_xs = 0, d_gs = 0, d_as = 0, d_wk[SVAE_MAX_LAYERS], d_wr[SVAE_MAX_LAYERS];
    long w_img_stride = 0;
};

static size_t take(size_t& cur, size_t bytes) {
    size_t off = cur;
    cur += (bytes + 1023) / 1024 * 1024;
    return off;
}

// ---- side stream: independent SMALL kernels of the backward tail run next to each other -------------------------------
// The tail of a step (first-layer parameter gradients, the encoder's dW / dX GEMMs and bias sums) is a chain of ~25
// kernels with grids of 16-128 CTAs on a 148-SM device; half of them are independent of each other.  They are forked
// onto one auxiliary stream per host thread and device (created on first use, kept for the life of the process;
// non-blocking) with a fork / join event pair, so eager launches overlap and a captured CUDA graph gets parallel
// branches.  The join happens before svae_step returns: nothing is left running on the auxiliary stream.
// SVAE_SIDE_STREAM=0 keeps everything on the caller's stream (A/B comparisons; read once per process).
struct Side {
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool forked = false;
};
static bool side_stream_enabled() {
    static const bool on = !(getenv("SVAE_SIDE_STREAM") != nullptr && getenv("SVAE_SIDE_STREAM")[0] == '0');
    return on;
}
static int side_get(Side** out) {
    static thread_local Side table[16];
    *out = nullptr;
    if (!side_stream_enabled()) return SVAE_OK;
    int dev = 0;
    SVAE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return SVAE_OK;
    Side& s = table[dev];
    if (s.aux == nullptr) {
        SVAE_CUDA(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
        SVAE_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        SVAE_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    s.forked = false;
    *out = &s;
    return SVAE_OK;
}
// everything enqueued on `st` so far is visible to work enqueued on the auxiliary stream from here on
static int side_fork(Side* s, cudaStream_t st) {
    if (s == nullptr) return SVAE_OK;
    SVAE_CUDA(cudaEventRecord(s->fork, st));
    SVAE_CUDA(cudaStreamWaitEvent(s->aux, s->fork, 0));
    s->forked = true;
    return SVAE_OK;
}
// `st` waits for everything enqueued on the auxiliary stream so far
static int side_join(Side* s, cudaStream_t st) {
    if (s == nullptr || !s->forked) return SVAE_OK;
    SVAE_CUDA(cudaEventRecord(s->join, s->aux));
    SVAE_CUDA(cudaStreamWaitEvent(st, s->join, 0));
    return SVAE_OK;
}
static cudaStream_t side_or(Side* s, cudaStream_t st) { return s != nullptr ? s->aux : st; }

static int validate(const SvaeShape& s, const SvaeConfig& c) {
    SVAE_REQUIRE(s.B >= 0 && s.P > 0 && s.H > 0 && s.L >= 1 && s.L <= SVAE_MAX_LAYERS, SVAE_EINVAL,
                 "bad shape: B=%d P=%d H=%d L=%d", s.B, s.P, s.H, s.L);
    SVAE_REQUIRE(s.C >= 1 && s.C <= 4, SVAE_EINVAL, "n_out=%d not supported (1..4)", s.C);
    SVAE_REQUIRE(s.Z >= 0 && s.I == s.Z + (c.rotate ? 1 : 0) + (c.translate ? 2 : 0), SVAE_EINVAL,
                 "inference dim %d != z_dim %d + rotate + 2*translate", s.I, s.Z);
    SVAE_REQUIRE(c.activation >= 0 && c.activation <= 3, SVAE_EINVAL, "unknown activation %d", c.activation);
    SVAE_REQUIRE(c.precision == SVAE_PRECISION_PARITY || c.precision == SVAE_PRECISION_FAST ||
                 c.precision == SVAE_PRECISION_PARITY_TC, SVAE_EINVAL, "unknown precision %d", c.precision);
    if (c.likelihood == SVAE_LIK_GAUSS_FITNOISE) {
        SVAE_REQUIRE(s.C == 2, SVAE_EINVAL, "fit-noise needs n_out == 2 (train_particles.py:447-449)");
        // the reference's variance convolution lacks groups= and crashes (train_particles.py:121-124,137)
        SVAE_REQUIRE(s.k_ctf == 0, SVAE_EINVAL, "CTF with fit-noise is rejected by the reference");
    } else if (c.likelihood == SVAE_LIK_GAUSS) {
        SVAE_REQUIRE(s.C == 1, SVAE_EINVAL, "Gaussian likelihood needs n_out == 1");
        if (s.k_ctf > 0)
            SVAE_REQUIRE((s.k_ctf & 1) && s.n_rows * s.n_cols == s.P, SVAE_EINVAL,
                         "CTF kernels must be odd-sized and n_rows*n_cols == P");
    } else {
        SVAE_REQUIRE(c.likelihood == SVAE_LIK_BERNOULLI, SVAE_EINVAL, "unknown likelihood %d", c.likelihood);
        SVAE_REQUIRE(s.k_ctf == 0, SVAE_EINVAL, "CTF only applies to the Gaussian likelihood");
    }
    SVAE_REQUIRE(!c.bilinear || s.Z > 0, SVAE_EINVAL, "bilinear needs a latent (models.py:73-75)");
    return SVAE_OK;
}

// ResidLinear networks (models.py:13-21) in FAST precision run on the tensor cores like every other network: the
// forward GEMM adds the layer input tile exactly in its epilogue (tc_gemm RES), the dX GEMM sees a bf16 copy of W + I
// (its diagonal rounding only touches the gradient), the encoder's 3-term GEMMs an fp32 copy of W + I.  (Folding the
// skip connection into the FORWARD operand was measured at 3e-3..9e-3 relative per-image ELBO error and is not done.)
// Validated on a B200 in round 1 (tests/test_gpu_zz_options.py).  SVAE_RESID_TC=0 falls back to the fp32 FFMA kernels,
// where the skip connection rides in the GEMM epilogue (A/B comparisons; read once per process).
static bool resid_on_tensor_cores() {
    static const bool on = !(getenv("SVAE_RESID_TC") != nullptr && getenv("SVAE_RESID_TC")[0] == '0');
    return on;
}
static bool use_fast(const SvaeConfig& c) {
    return c.precision == SVAE_PRECISION_FAST && (!c.resid || resid_on_tensor_cores());
}
// PARITY_TC: fp32 activations, every hidden GEMM of BOTH networks as one bf16 tcgen05 GEMM over three hi/lo split
// terms (error ~2^-16, fp32 accumulation); everything else as PARITY.  ResidLinear networks and the first-layer
// options run PARITY instead (their fp32 epilogue addends have no tensor-core variant).
static bool use_tc3(const SvaeConfig& c) {
    return c.precision == SVAE_PRECISION_PARITY_TC && !c.resid && !c.expand_coords && !c.bilinear;
}

static int make_plan(const SvaeShape& s, const SvaeConfig& c, Plan& p) {
    const bool fast = use_fast(c);
    const bool tc3 = use_tc3(c);
    p.tc3 = tc3;
    p.Hp = (fast || tc3) ? (int)round_up(s.H, 64) : (int)round_up(s.H, 2);
    p.esize = fast ? 2 : 4;
    p.F = c.expand_coords ? 5 : 2;
    p.opt = c.expand_coords || c.bilinear;
    p.K1 = p.opt ? p.F + 1 : 3;
    p.w_img_stride = c.bilinear ? (long)s.H * p.F : 0;
    int chunk = c.chunk_images;
    if (chunk <= 0) {
        // bound the activation workspace (L act + 2 delta matrices, + the split buffers of PARITY_TC) to ~6 GiB per pass
        const double per_image = (double)s.P * p.Hp * (p.esize * (s.L + 2) + (tc3 ? 18 : 0));
        chunk = (int)fmax(1.0, floor(6.0 * 1024 * 1024 * 1024 / per_image));
    }
    if (chunk > s.B) chunk = s.B;
    if (chunk < 1) chunk = 1;
    p.chunk = chunk;
    const size_t B = (size_t)(s.B > 0 ? s.B : 1), I = (size_t)s.I, rows = (size_t)chunk * s.P;
    size_t cur = 0;
    const size_t Hqp = (size_t)round_up(s.Hq, 64);
    p.enc_ld = (fast || tc3) ? (int)Hqp : s.Hq;
    p.enc_acts = take(cur, (size_t)s.Lq * B * p.enc_ld * 4);
    p.zo = take(cur, B * 2 * I * 4);
    const size_t wide = (size_t)((size_t)p.enc_ld > 2 * I ? (size_t)p.enc_ld : 2 * I);
    p.enc_scratch = take(cur, 2 * B * wide * 4);
    if (fast || tc3) {
        const size_t kp0 = (size_t)round_up((long)s.P * s.Cin, 64);
        const size_t kmax = kp0 > Hqp ? kp0 : Hqp;
        p.xs_k = take(cur, B * 3 * kmax * 2);
        p.gs_r = take(cur, 3 * B * Hqp * 2);
        p.as_r = take(cur, 3 * B * kmax * 2);
        for (int l = 0; l < s.Lq; ++l) {
            const size_t kp = l == 0 ? kp0 : Hqp;
            p.ws_k[l] = take(cur, Hqp * 3 * kp * 2);
            p.ws_r[l] = take(cur, l == 0 ? 0 : 3 * Hqp * kp * 2);
        }
    }
    p.lat = take(cur, B * I * 4);
    p.eps = take(cur, B * I * 4);          // the in-kernel eps draw, kept for the backward
    p.img = take(cur, B * 4 * 4);
    p.zs = take(cur, B * (size_t)(s.Z > 0 ? s.Z : 1) * 4);
    p.hz = take(cur, B * p.Hp * 4);
    p.S = take(cur, B * p.K1 * p.Hp * 4);
    p.weff = take(cur, c.bilinear ? B * (size_t)s.H * p.F * 4 : 0);
    p.dlat = take(cur, B * 3 * 4);
    p.resid_tc = fast && c.resid;
    p.enc_wres = take(cur, (p.resid_tc && s.Lq > 1) ? (size_t)s.Hq * s.Hq * 4 : 0);
    p.dz = take(cur, B * (size_t)(s.Z > 0 ? s.Z : 1) * 4);
    p.g_zo = take(cur, B * 2 * I * 4);
    p.o = take(cur, rows * s.C * 4);
    p.g_o = take(cur, rows * s.C * 4);
    p.act_stride = (rows * p.Hp * p.esize + 1023) / 1024 * 1024;
    p.acts = take(cur, p.act_stride * s.L);
    p.delta_stride = p.act_stride;
    p.delta = take(cur, p.delta_stride * 2);
    if (tc3) {
        p.d_xs = take(cur, rows * 3 * p.Hp * 2);
        p.d_gs = take(cur, 3 * rows * p.Hp * 2);
        p.d_as = take(cur, 3 * rows * p.Hp * 2);
    }
    p.w_stride = (size_t)p.Hp * p.Hp * 2;
    // bf16 hidden weights; with resid_tc a second set with the identity added follows (operands of the dX GEMMs)
    p.wbf16 = take(cur, fast ? p.w_stride * (s.L > 1 ? s.L - 1 : 1) * (p.resid_tc ? 2 : 1) : 0);
    for (int l = 0; l < s.L - 1; ++l) {
        p.d_wk[l] = take(cur, tc3 ? (size_t)p.Hp * 3 * p.Hp * 2 : 0);
        p.d_wr[l] = take(cur, tc3 ? (size_t)3 * p.Hp * p.Hp * 2 : 0);
    }
    p.total = cur;
    return SVAE_OK;
}

// ---- encoder -------------------------------------------------------------------------------------
// the 2I-wide head (models.py:41): small, stays on the fp32 FFMA kernel in both precisions
static int encoder_head_forward(const SvaeShape& s, const SvaeEncoderParams& q, const float* h, long ld, float* out,
                                cudaStream_t st) {
    SgemmArgs a{};
    a.A = h; a.sAm = ld; a.sAk = 1;
    a.B = q.w[s.Lq]; a.sBk = 1; a.sBn = s.Hq;
    a.C = out; a.ldc = 2 * s.I;
    a.M = s.B; a.N = 2 * s.I; a.K = s.Hq;
    a.bias = q.b[s.Lq];
    // narrow output (2I columns): few tiles walking a long K are latency bound -> spread K over 4x the CTAs
    if (s.Hq >= 256 && (long)s.B * 2 * s.I <= (1L << 20)) {
        SVAE_CUDA(cudaMemsetAsync(out, 0, (size_t)s.B * 2 * s.I * sizeof(float), st));
        a.split_k = 4;
    }
    return sgemm(a, st);
}

// resid: hidden layers 1..Lq-1 are ResidLinear, act(W h + b + h) (models.py:13-21,35-36)
static int encoder_forward_impl(const SvaeShape& s, int act, const SvaeEncoderParams& q, const float* x, float* out,
                                float* acts, cudaStream_t st, bool resid = false) {
    const int n_in = s.P * s.Cin;
    const float* cur = x;
    int k = n_in;
    for (int l = 0; l < s.Lq; ++l) {
        SgemmArgs a{};
        a.A = cur; a.sAm = k; a.sAk = 1;
        a.B = q.w[l]; a.sBk = 1; a.sBn = k;
        float* dst = acts + (size_t)l * s.B * s.Hq;
        a.C = dst; a.ldc = s.Hq;
        a.M = s.B; a.N = s.Hq; a.K = k;
        a.bias = q.b[l]; a.act = act;
        if (resid && l > 0) { a.add = cur; a.ld_add = k; }
        SVAE_TRY(sgemm(a, st));
        cur = dst; k = s.Hq;
    }
    return encoder_head_forward(s, q, cur, s.Hq, out, st);
}

// ---- encoder hidden layers on tensor cores (FAST): every fp32 GEMM = one bf16 tcgen05 GEMM over K' = 3K
// built from hi/lo splits of both operands (split3), fp32 accumulate, fp32 output.
struct EncTc {
    const SvaeShape* s; const Plan* p; char* ws; cudaStream_t st;
    Side* side = nullptr;      // independent dW / bias work of the backward goes to the auxiliary stream
    // weight of hidden layer l as the GEMMs see it: W_l, or an fp32 copy of W_l + I for a ResidLinear layer (l >= 1)
    int weight(const SvaeEncoderParams& q, int l, const float** w) const {
        *w = q.w[l];
        if (p->resid_tc && l > 0) {
            float* tmp = reinterpret_cast<float*>(ws + p->enc_wres);
            SVAE_TRY(copy_add_identity(q.w[l], s->Hq, tmp, st));
            *w = tmp;
        }
        return SVAE_OK;
    }
    int Hqp() const { return p->enc_ld; }
    int kin(int l) const { return l == 0 ? s->P * s->Cin : s->Hq; }
    int kinp(int l) const { return l == 0 ? (int)round_up((long)s->P * s->Cin, 64) : p->enc_ld; }
    float* acts(int l) const { return reinterpret_cast<float*>(ws + p->enc_acts) + (size_t)l * s->B * p->enc_ld; }
    __nv_bfloat16* b16(size_t off) const { return reinterpret_cast<__nv_bfloat16*>(ws + off); }
};

static int encoder_forward_tc(const EncTc& e, int act, const SvaeEncoderParams& q, const float* x, float* out) {
    const SvaeShape& s = *e.s;
    const float* cur = x;
    long ld = e.kin(0);
    // A minibatch gives these GEMMs only a few 256 x 256 output tiles, each walking a long K (3 x 784 at C2) on one
    // CTA pair while the rest of the GPU idles: when at least two thirds of the pairs would idle, K is split over
    // them, the GEMM adds RAW partial sums into the zeroed activation buffer and the layer is finished (bias,
    // activation, in place) by the kernel that splits it for the next layer anyway.  Same-box A/B: C1 0.371 -> 0.351
    // ms, C2 -14 us, C3 -30 us, C4 -0.1 ms; with only 2 splits (C5, 4096 images) the extra passes cost 0.4 %.
    const bool raw = tc_split_k_factor(s.B, e.Hqp()) >= 3;
    if (raw) SVAE_CUDA(cudaMemsetAsync(e.acts(0), 0, (size_t)s.Lq * s.B * e.Hqp() * sizeof(float), e.st));
    for (int l = 0; l < s.Lq; ++l) {
        const int k = e.kin(l), kp = e.kinp(l);
        if (raw && l > 0)
            SVAE_TRY(split3_act(e.acts(l - 1), q.b[l - 1], act, s.B, s.Hq, e.b16(e.p->xs_k), s.B, kp, 1, 0, e.st));
        else
            SVAE_TRY(split3(cur, s.B, k, ld, e.b16(e.p->xs_k), s.B, kp, 1, 0, e.st));
        const float* wl = nullptr;
        SVAE_TRY(e.weight(q, l, &wl));
        SVAE_TRY(split3(wl, s.Hq, k, k, e.b16(e.p->ws_k[l]), e.Hqp(), kp, 1, 1, e.st));
        TcExtra f32out;
        f32out.out_f32 = 1;
        f32out.raw_split_k = raw ? 1 : 0;
        SVAE_TRY(tc_gemm(0, s.B, e.Hqp(), 3 * kp, e.b16(e.p->xs_k), 3 * kp, e.b16(e.p->ws_k[l]), 3 * kp, q.b[l], s.Hq,
                         nullptr, 0, act, e.acts(l), e.Hqp(), e.st, f32out));
        cur = e.acts(l); ld = e.Hqp();
    }
    if (raw) SVAE_TRY(split3_act(e.acts(s.Lq - 1), q.b[s.Lq - 1], act, s.B, s.Hq, nullptr, s.B, e.Hqp(), 1, 0, e.st));
    return encoder_head_forward(s, q, cur, ld, out, e.st);
}

// g_out (B,2I) is the gradient w.r.t. the head output; scratch holds two (B, max(Hq,2I)) buffers.
static int encoder_backward_tc(const EncTc& e, int act, const SvaeEncoderParams& q, const float* x,
                               const float* g_out, SvaeEncoderParams& gq, float* scratch) {
    const SvaeShape& s = *e.s;
    const int Hqp = e.Hqp();
    const size_t wide = (size_t)(Hqp > 2 * s.I ? Hqp : 2 * s.I);
    float* buf[2] = {scratch, scratch + (size_t)s.B * wide};
    // Two chains per layer: dW / db (reads g and the layer input, accumulates into the gradient buffers) on the
    // auxiliary stream, the dX GEMM that produces the next g on the caller's stream.
    cudaStream_t sw = side_or(e.side, e.st);
    // head (2I x Hq): fp32 FFMA
    const float* a_top = e.acts(s.Lq - 1);
    SVAE_TRY(side_fork(e.side, e.st));
    SgemmArgs w{};
    w.A = g_out; w.sAm = 1; w.sAk = 2 * s.I;
    w.B = a_top; w.sBk = Hqp; w.sBn = 1;
    w.C = gq.w[s.Lq]; w.ldc = s.Hq;
    w.M = 2 * s.I; w.N = s.Hq; w.K = s.B;
    w.accumulate = 1; w.split_k = s.B >= 512 ? 8 : (s.B >= 128 ? 2 : 1);
    SVAE_TRY(sgemm(w, sw));
    SVAE_TRY(col_sum<float>(g_out, s.B, 2 * s.I, 2 * s.I, gq.b[s.Lq], sw));
    SgemmArgs d{};
    d.A = g_out; d.sAm = 2 * s.I; d.sAk = 1;
    d.B = q.w[s.Lq]; d.sBk = s.Hq; d.sBn = 1;
    d.C = buf[0]; d.ldc = Hqp;
    d.M = s.B; d.N = s.Hq; d.K = 2 * s.I;
    d.dsrc = a_top; d.ld_dsrc = Hqp; d.dact = act;
    SVAE_TRY(sgemm(d, e.st));
    const float* g = buf[0];
    int cur = 0;
    for (int l = s.Lq - 1; l >= 0; --l) {
        const float* a_in = (l == 0) ? x : e.acts(l - 1);
        const int k_in = e.kin(l), kinp = e.kinp(l);
        const long ld_in = (l == 0) ? k_in : Hqp;
        // the auxiliary stream sees g (just produced on the caller's stream) ...
        SVAE_TRY(side_fork(e.side, e.st));
        // dW_l (Hq, k_in) += g^T a_in : terms stacked along the reduction dimension (rows)
        SVAE_TRY(split3(g, s.B, s.Hq, Hqp, e.b16(e.p->gs_r), s.B, Hqp, 0, 0, sw));
        SVAE_TRY(split3(a_in, s.B, k_in, ld_in, e.b16(e.p->as_r), s.B, kinp, 0, 1, sw));
        SVAE_TRY(tc_gemm(2, s.Hq, k_in, 3 * s.B, e.b16(e.p->gs_r), Hqp, e.b16(e.p->as_r), kinp, nullptr, 0, nullptr, 0, -1,
                         gq.w[l], k_in, sw));
        SVAE_TRY(col_sum<float>(g, s.B, s.Hq, Hqp, gq.b[l], sw));
        if (l == 0) break;
        // g_in (B, Hq) = (g W_l) .* act'(a_in)
        SVAE_TRY(split3(g, s.B, s.Hq, Hqp, e.b16(e.p->xs_k), s.B, Hqp, 1, 0, e.st));
        const float* wl = nullptr;
        SVAE_TRY(e.weight(q, l, &wl));
        SVAE_TRY(split3(wl, s.Hq, k_in, k_in, e.b16(e.p->ws_r[l]), Hqp, kinp, 0, 1, e.st));
        cur ^= 1;
        TcExtra f32out;
        f32out.out_f32 = 1;
        // (the K split of the forward does not pay here: with the separate act' pass it measured 1 us slower at C1)
        SVAE_TRY(tc_gemm(1, s.B, kinp, 3 * Hqp, e.b16(e.p->xs_k), 3 * Hqp, e.b16(e.p->ws_r[l]), kinp, nullptr, 0, a_in,
                         Hqp, act, buf[cur], Hqp, e.st, f32out));
        // ... and must be done reading it (and its split buffers) before the NEXT layer's dX overwrites the other
        // ping-pong buffer's predecessor: join here keeps the two chains one layer apart at most
        SVAE_TRY(side_join(e.side, e.st));
        g = buf[cur];
    }
    return SVAE_OK;
}

static int encoder_backward_impl(const SvaeShape& s, int act, const SvaeEncoderParams& q, const float* x,
                                 const float* acts, const float* g_out, SvaeEncoderParams& gq, float* g_x,
                                 float* scratch, cudaStream_t st, bool resid = false) {
    const int n_in = s.P * s.Cin;
    const size_t wide = (size_t)(s.Hq > 2 * s.I ? s.Hq : 2 * s.I);
    float* buf[2] = {scratch, scratch + (size_t)s.B * wide};
    const float* g = g_out;
    int gn = 2 * s.I;                      // width of g
    int split = s.B >= 2048 ? 8 : (s.B >= 512 ? 4 : 1);
    for (int l = s.Lq; l >= 0; --l) {
        const float* a_in = (l == 0) ? x : acts + (size_t)(l - 1) * s.B * s.Hq;
        const int k_in = (l == 0) ? n_in : s.Hq;
        // dW[l] (gn, k_in) += g^T a_in
        SgemmArgs w{};
        w.A = g; w.sAm = 1; w.sAk = gn;
        w.B = a_in; w.sBk = k_in; w.sBn = 1;
        w.C = gq.w[l]; w.ldc = k_in;
        w.M = gn; w.N = k_in; w.K = s.B;
        w.accumulate = 1; w.split_k = split;
        SVAE_TRY(sgemm(w, st));
        SVAE_TRY(col_sum<float>(g, s.B, gn, gn, gq.b[l], st));
        if (l == 0 && g_x == nullptr) break;
        // g_in (B, k_in) = (g W[l]) .* act'(a_in)
        SgemmArgs d{};
        d.A = g; d.sAm = gn; d.sAk = 1;
        d.B = q.w[l]; d.sBk = k_in; d.sBn = 1;
        float* dst = (l == 0) ? g_x : buf[l & 1];
        d.C = dst; d.ldc = k_in;
        d.M = s.B; d.N = k_in; d.K = gn;
        if (l > 0) { d.dsrc = a_in; d.ld_dsrc = k_in; d.dact = act; }
        if (resid && l > 0 && l < s.Lq) { d.add = g; d.ld_add = gn; }      // skip connection of ResidLinear l
        SVAE_TRY(sgemm(d, st));
        g = dst; gn = k_in;
    }
    return SVAE_OK;
}

// ---- decoder passes over one chunk of images -----------------------------------------------------
template <typename T>
struct DecoderCtx {
    const SvaeShape* s;
    const SvaeConfig* c;
    const Plan* p;
    char* ws;
    cudaStream_t st;
    T* act(int l) const { return reinterpret_cast<T*>(ws + p->acts + p->act_stride * l); }
    T* delta(int i) const { return reinterpret_cast<T*>(ws + p->delta + p->delta_stride * i); }
    float* logits() const { return reinterpret_cast<float*>(ws + p->o); }
    float* g_logits() const { return reinterpret_cast<float*>(ws + p->g_o); }
    float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
    __nv_bfloat16* wbf(int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + p->wbf16 + p->w_stride * l); }
    __nv_bfloat16* b16(size_t off) const { return reinterpret_cast<__nv_bfloat16*>(ws + off); }
    // operand of the dX GEMM of hidden layer l+1: W, or the W + I copy for ResidLinear layers on tensor cores
    __nv_bfloat16* wbf_dx(int l) const { return wbf(p->resid_tc ? (s->L - 1) + l : l); }
    bool bwd_follows = false;        // set by the callers that run decoder_chunk_backward after the forward
    bool fwd_splits_for_bwd() const { return bwd_follows && p->tc3 && s->L == 2; }
    // first-layer coordinate weights of image 0: W_eff (B, H*F) with bilinear, else coord_linear.weight
    const float* l0_w(const SvaeDecoderParams& dp) const { return c->bilinear ? f(p->weff) : dp.coord_w; }
};

// fuse_out: also accumulate the output-layer logits o (rows, C) (pre-filled with out_b) in the epilogue
template <typename T>
static int hidden_forward(const DecoderCtx<T>& d, const SvaeDecoderParams& dp, int l, int rows, bool fuse_out);
template <>
int hidden_forward<float>(const DecoderCtx<float>& d, const SvaeDecoderParams& dp, int l, int rows, bool) {
    const int H = d.s->H, Hp = d.p->Hp;
    if (d.p->tc3) {
        // act[l] = act(act[l-1] W^T + b): operands split into (hi, hi, lo) x (hi, lo, hi) bf16 terms along K
        __nv_bfloat16* xs = d.b16(d.p->d_xs);
        __nv_bfloat16* wk = reinterpret_cast<__nv_bfloat16*>(d.ws + d.p->d_wk[l - 1]);
        // a backward over the same chunk follows and there is one hidden GEMM: its dW operand (the same matrix,
        // row-stacked, B-side terms) is written in the same pass
        if (d.fwd_splits_for_bwd())
            SVAE_TRY(split3_both(d.act(l - 1), rows, H, Hp, d.b16(d.p->d_as), 1, xs, 0, rows, Hp, d.st));
        else
            SVAE_TRY(split3(d.act(l - 1), rows, H, Hp, xs, rows, Hp, 1, 0, d.st));
        TcExtra f32out;
        f32out.out_f32 = 1;
        return tc_gemm(0, rows, Hp, 3 * Hp, xs, 3 * Hp, wk, 3 * Hp, dp.hidden_b[l - 1], H, nullptr, 0, d.c->activation,
                       d.act(l), Hp, d.st, f32out);
    }
    SgemmArgs a{};
    a.A = d.act(l - 1); a.sAm = Hp; a.sAk = 1;
    a.B = dp.hidden_w[l - 1]; a.sBk = 1; a.sBn = H;
    a.C = d.act(l); a.ldc = Hp;
    a.M = rows; a.N = H; a.K = H;
    a.bias = dp.hidden_b[l - 1]; a.act = d.c->activation;
    if (d.c->resid) { a.add = d.act(l - 1); a.ld_add = Hp; }     // ResidLinear: act(W h + b + h)
    return sgemm(a, d.st);
}
template <>
int hidden_forward<__nv_bfloat16>(const DecoderCtx<__nv_bfloat16>& d, const SvaeDecoderParams& dp, int l, int rows,
                                  bool fuse_out) {
    const int Hp = d.p->Hp;
    if (fuse_out) SVAE_TRY(fill_rows(d.logits(), dp.out_b, rows, d.s->C, d.s->C, d.st));
    TcExtra ex;
    if (fuse_out) { ex.out_w = dp.out_w; ex.out_w_ld = d.s->H; ex.dot_c = d.s->C; ex.o_accum = d.logits(); }
    if (d.p->resid_tc) { ex.resid = d.act(l - 1); ex.ld_resid = Hp; }      // act(W h + b + h), h added in the epilogue
    return tc_gemm(0, rows, Hp, Hp, d.act(l - 1), Hp, d.wbf(l - 1), Hp, dp.hidden_b[l - 1], d.s->H, nullptr, 0,
                   d.c->activation, d.act(l), Hp, d.st, ex);
}

// red: (FAST only) delta_prev is not stored: the transposed dX GEMM reduces it per image into S (tc_bwd.cu)
struct RedSpec { TcMoments m; bool on = false; bool top = false; };
template <typename T>
static int hidden_backward(const DecoderCtx<T>& d, const SvaeDecoderParams& dp, SvaeDecoderParams& g, int l, int rows,
                           const T* delta, T* delta_prev, const RedSpec& red);
template <>
int hidden_backward<float>(const DecoderCtx<float>& d, const SvaeDecoderParams& dp, SvaeDecoderParams& g, int l,
                           int rows, const float* delta, float* delta_prev, const RedSpec&) {
    const int H = d.s->H, Hp = d.p->Hp;
    if (d.p->tc3) {
        // dW_l += delta^T act[l-1]: the three terms stacked along the reduction dimension (rows)
        // (delta is split ONCE: row-stacked for this GEMM, K-concatenated for the dX GEMM below)
        SVAE_TRY(split3_both(delta, rows, H, Hp, d.b16(d.p->d_gs), 0, d.b16(d.p->d_xs), 0, rows, Hp, d.st));
        // (with two hidden layers the forward already left act[l-1] here, see hidden_forward)
        if (!d.fwd_splits_for_bwd())
            SVAE_TRY(split3(d.act(l - 1), rows, H, Hp, d.b16(d.p->d_as), rows, Hp, 0, 1, d.st));
        SVAE_TRY(tc_gemm(2, H, H, 3 * rows, d.b16(d.p->d_gs), Hp, d.b16(d.p->d_as), Hp, nullptr, 0, nullptr, 0, -1,
                         g.hidden_w[l - 1], H, d.st));
        // delta_prev = (delta W_l) .* act'(act[l-1])
        TcExtra f32out;
        f32out.out_f32 = 1;
        return tc_gemm(1, rows, Hp, 3 * Hp, d.b16(d.p->d_xs), 3 * Hp, reinterpret_cast<__nv_bfloat16*>(d.ws + d.p->d_wr[l - 1]),
                       Hp, nullptr, 0, d.act(l - 1), Hp, d.c->activation, delta_prev, Hp, d.st, f32out);
    }
    // dW_l (H,H) += delta^T act[l-1]
    SgemmArgs w{};
    w.A = delta; w.sAm = 1; w.sAk = Hp;
    w.B = d.act(l - 1); w.sBk = Hp; w.sBn = 1;
    w.C = g.hidden_w[l - 1]; w.ldc = H;
    w.M = H; w.N = H; w.K = rows;
    w.accumulate = 1;
    const int tiles = ceil_div(H, 128) * ceil_div(H, 128);
    w.split_k = max(1, min(ceil_div(rows, 256), ceil_div(2 * 148, tiles)));
    SVAE_TRY(sgemm(w, d.st));
    // delta_prev = (delta W_l) .* act'(act[l-1])
    SgemmArgs x{};
    x.A = delta; x.sAm = Hp; x.sAk = 1;
    x.B = dp.hidden_w[l - 1]; x.sBk = H; x.sBn = 1;
    x.C = delta_prev; x.ldc = Hp;
    x.M = rows; x.N = H; x.K = H;
    x.dsrc = d.act(l - 1); x.ld_dsrc = Hp; x.dact = d.c->activation;
    if (d.c->resid) { x.add = delta; x.ld_add = Hp; }            // gradient through the skip connection
    return sgemm(x, d.st);
}
template <>
int hidden_backward<__nv_bfloat16>(const DecoderCtx<__nv_bfloat16>& d, const SvaeDecoderParams& dp,
                                   SvaeDecoderParams& g, int l, int rows, const __nv_bfloat16* delta,
                                   __nv_bfloat16* delta_prev, const RedSpec& red) {
    const int H = d.s->H, Hp = d.p->Hp;
    (void)dp;
    // dW_l (H,H; ld H) += delta^T act[l-1]   (fp32 accumulate, atomics across row splits)
    if (red.top) {
        // top layer: delta is built from act[l] and g_o inside the GEMM and written to `delta` for the dX pass below
        TcTop t;
        t.g_o = d.g_logits(); t.C = d.s->C; t.out_w = dp.out_w;
        t.d_out_w = g.out_w; t.d_out_b = g.out_b; t.d_b = g.hidden_b[l - 1];
        SVAE_TRY(tc_dw_top(rows, H, Hp, d.act(l), d.act(l - 1), d.c->activation, t, g.hidden_w[l - 1], H,
                           const_cast<__nv_bfloat16*>(delta), d.st));
    } else {
        SVAE_TRY(tc_gemm(2, H, H, rows, delta, Hp, d.act(l - 1), Hp, nullptr, 0, nullptr, 0, -1, g.hidden_w[l - 1], H, d.st));
    }
    if (red.on) return tc_dx_moments(rows, H, Hp, delta, Hp, d.wbf_dx(l - 1), Hp, d.c->activation, red.m, d.st);
    return tc_gemm(1, rows, Hp, Hp, delta, Hp, d.wbf_dx(l - 1), Hp, nullptr, 0, d.act(l - 1), Hp, d.c->activation,
                   delta_prev, Hp, d.st);
}

// columns [H, Hp) of `mats` row-major (rows, Hp) fp32 matrices that lie `stride` floats apart <- 0; thread = one row
__global__ void __launch_bounds__(256) zero_pad_columns_k(float* __restrict__ base, int rows, int mats, size_t stride,
                                                          int H, int Hp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * mats) return;
    float* p = base + (size_t)(i / rows) * stride + (size_t)(i % rows) * Hp;
    for (int c = H; c < Hp; ++c) p[c] = 0.f;
}

// forward of the decoder over images [b0, b0+nb): fills act[0..L-1] and logits o; optional y_hat
template <typename T>
static int decoder_chunk_forward(const DecoderCtx<T>& d, const SvaeDecoderParams& dp, int b0, int nb,
                                 const float* grid, const float* x_explicit, float* y_hat) {
    const SvaeShape& s = *d.s;
    const int Hp = d.p->Hp, rows = nb * s.P;
    if (rows == 0) return SVAE_OK;
    if (std::is_same<T, float>::value && Hp != s.H) {
        // fp32 activations with padded columns: kernels that write H columns only must find zeros in the other
        // Hp - H (they are multiplied by zero-padded weights, but 0 * garbage may be NaN).  Only those columns of the
        // chunk's rows are cleared (clearing whole matrices cost 1.6 of 9.4 ms per step in parity_tc at C2; pitched
        // 2-D memsets are no faster and split into many launches).
        // The L activation and 2 delta matrices lie one act_stride apart: one launch clears them all.
        const long n = (long)rows * (s.L + 2);
        zero_pad_columns_k<<<ceil_div(n, 256), 256, 0, d.st>>>(reinterpret_cast<float*>(d.ws + d.p->acts), rows, s.L + 2,
                                                               d.p->act_stride / sizeof(float), s.H, Hp);
        SVAE_LAUNCH_CHECK();
    }
    if (d.p->opt)
        SVAE_TRY(layer0_opt_forward<T>(d.p->F, s.P, d.c->activation, b0, nb, d.l0_w(dp), d.p->w_img_stride,
                                       d.f(d.p->hz), grid, d.f(d.p->img), x_explicit, s.H, Hp, d.act(0), d.st));
    else
        SVAE_TRY(layer0_forward<T>(s, d.c->activation, b0, nb, dp.coord_w, d.f(d.p->hz), grid, d.f(d.p->img),
                                   x_explicit, s.H, Hp, d.act(0), d.st));
    // FAST precision: the output-layer dot product rides in the epilogue of the last hidden GEMM
    const bool fuse_out = !std::is_same<T, float>::value && s.L >= 2 && s.C <= 3 && !d.p->resid_tc;
    for (int l = 1; l < s.L; ++l) SVAE_TRY(hidden_forward<T>(d, dp, l, rows, fuse_out && l == s.L - 1));
    float* yh = y_hat ? y_hat + (size_t)b0 * s.P * s.C : nullptr;
    if (fuse_out) return yh ? logits_to_yhat(d.logits(), yh, (long)rows * s.C, s.C, d.c->softplus, d.st) : SVAE_OK;
    return out_forward<T>(d.act(s.L - 1), rows, s.H, Hp, s.C, dp.out_w, dp.out_b, d.c->softplus, d.logits(), yh, d.st);
}

// backward over the same chunk given g_o (rows, C) in the workspace: parameter grads, S[b0..]
template <typename T>
static int decoder_chunk_backward(const DecoderCtx<T>& d, const SvaeDecoderParams& dp, SvaeDecoderParams& g, int b0,
                                  int nb, const float* grid, const float* x_explicit, float* g_x) {
    const SvaeShape& s = *d.s;
    const int Hp = d.p->Hp, rows = nb * s.P;
    if (rows == 0) return SVAE_OK;
    int cur = 0;
    // FAST: delta_{L-1} = (g_o W_o) .* act'(h_{L-1}) is produced inside the top layer's dW GEMM (tc_bwd.cu), which
    // also accumulates dW_o, db_o, db_{L-1} and writes delta_{L-1} once for the dX GEMM: no separate pass
    const bool fuse_top = !std::is_same<T, float>::value && s.L >= 2 && s.C <= 3;
    if (!fuse_top)
        SVAE_TRY(out_backward<T>(d.act(s.L - 1), d.g_logits(), rows, s.H, Hp, s.C, d.c->activation, dp.out_w,
                                 d.delta(cur), g.out_w, g.out_b, s.L >= 2 ? g.hidden_b[s.L - 2] : nullptr, d.st));
    // The dX GEMM of the first hidden layer does not store delta_0: it recomputes h_0 and reduces delta_0 per image
    // into S in its epilogue (tc_bwd.cu).  Explicit coordinates, coordinate gradients and the first-layer options
    // keep the stored delta_0 and the separate reductions.
    const bool fuse_red = !std::is_same<T, float>::value && s.L >= 2 && x_explicit == nullptr && g_x == nullptr &&
                          grid != nullptr && !d.p->opt;
    bool reduced = false;
    for (int l = s.L - 1; l >= 1; --l) {
        RedSpec red;
        if (fuse_red && l == 1) {
            red.on = reduced = true;
            red.m.grid = grid; red.m.img = d.f(d.p->img); red.m.coord_w = dp.coord_w; red.m.hz = d.f(d.p->hz);
            red.m.S = d.f(d.p->S); red.m.P = s.P; red.m.b0 = b0;
        }
        red.top = fuse_top && l == s.L - 1;
        SVAE_TRY(hidden_backward<T>(d, dp, g, l, rows, d.delta(cur), d.delta(cur ^ 1), red));
        cur ^= 1;
        if (l - 1 >= 1) SVAE_TRY(col_sum<T>(d.delta(cur), rows, s.H, Hp, g.hidden_b[l - 2], d.st));
    }
    if (d.p->opt) {
        SVAE_TRY(image_feat_reduce<T>(d.p->F, d.delta(cur), b0, nb, s.P, Hp, grid, d.f(d.p->img), x_explicit,
                                      d.f(d.p->S), d.st));
        if (g_x)
            SVAE_TRY(coord_row_grad_opt<T>(d.p->F, d.delta(cur), rows, s.P, s.H, Hp,
                                           d.l0_w(dp) + (size_t)b0 * d.p->w_img_stride, d.p->w_img_stride,
                                           x_explicit + (size_t)b0 * s.P * 2, g_x + (size_t)b0 * s.P * 2, d.st));
        return SVAE_OK;
    }
    if (!reduced) SVAE_TRY(image_col_reduce<T>(d.delta(cur), b0, nb, s.P, Hp, grid, x_explicit, d.f(d.p->S), d.st));
    if (g_x) SVAE_TRY(coord_row_grad<T>(d.delta(cur), rows, s.H, Hp, dp.coord_w, g_x + (size_t)b0 * s.P * 2, d.st));
    return SVAE_OK;
}

static int prepare_bf16_weights(const SvaeShape& s, const Plan& p, const SvaeDecoderParams& dp, char* ws,
                                cudaStream_t st) {
    for (int l = 0; l < s.L - 1; ++l) {
        __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(ws + p.wbf16 + p.w_stride * l);
        SVAE_TRY(to_bf16_padded(dp.hidden_w[l], s.H, s.H, w, p.Hp, p.Hp, st));
        if (p.resid_tc) {      // second copy with the identity added: the dX operand of a ResidLinear layer
            __nv_bfloat16* wi = reinterpret_cast<__nv_bfloat16*>(ws + p.wbf16 + p.w_stride * ((s.L - 1) + l));
            SVAE_TRY(to_bf16_padded(dp.hidden_w[l], s.H, s.H, wi, p.Hp, p.Hp, st));
            SVAE_TRY(add_identity_bf16(dp.hidden_w[l], s.H, wi, p.Hp, st));
        }
    }
    return SVAE_OK;
}

// PARITY_TC: hi/lo split terms of the hidden weights, once per call: K-concatenated for the forward GEMM
// (Hp x 3Hp, B side), row-stacked for the dX GEMM (3Hp x Hp, B side)
static int prepare_split_weights(const SvaeShape& s, const Plan& p, const SvaeDecoderParams& dp, char* ws, cudaStream_t st) {
    for (int l = 0; l < s.L - 1; ++l) {
        SVAE_TRY(split3(dp.hidden_w[l], s.H, s.H, s.H, reinterpret_cast<__nv_bfloat16*>(ws + p.d_wk[l]), p.Hp, p.Hp, 1, 1, st));
        SVAE_TRY(split3(dp.hidden_w[l], s.H, s.H, s.H, reinterpret_cast<__nv_bfloat16*>(ws + p.d_wr[l]), p.Hp, p.Hp, 0, 1, st));
    }
    return SVAE_OK;
}

// hz (B,Hp) = zs Wz^T + coord_b
static int latent_projection(const SvaeShape& s, const Plan& p, const SvaeDecoderParams& dp, const float* zs,
                             float* hz, cudaStream_t st) {
    if (s.B == 0) return SVAE_OK;
    if (s.Z > 0 && dp.latent_w != nullptr) {
        SgemmArgs a{};
        a.A = zs; a.sAm = s.Z; a.sAk = 1;
        a.B = dp.latent_w; a.sBk = 1; a.sBn = s.Z;
        a.C = hz; a.ldc = p.Hp;
        a.M = s.B; a.N = s.H; a.K = s.Z;
        a.bias = dp.coord_b;
        return sgemm(a, st);
    }
    return fill_rows(hz, dp.coord_b, s.B, s.H, p.Hp, st);
}

// bilinear (models.py:114-121): per-image coordinate weights W_eff (B, H*F) = zs Wb^T + Wc, Wb viewed as (H*F, Z)
static int bilinear_weights(const SvaeShape& s, const Plan& p, const SvaeDecoderParams& dp, const float* zs,
                            float* weff, cudaStream_t st) {
    if (s.B == 0) return SVAE_OK;
    SgemmArgs a{};
    a.A = zs; a.sAm = s.Z; a.sAk = 1;
    a.B = dp.bilinear_w; a.sBk = 1; a.sBn = s.Z;
    a.C = weff; a.ldc = (long)s.H * p.F;
    a.M = s.B; a.N = s.H * p.F; a.K = s.Z;
    a.bias = dp.coord_w;
    return sgemm(a, st);
}

// option path of first_layer_param_grads: T (B, F+1, Hp) are the feature moments of delta0 (first_layer.cuh)
static int first_layer_param_grads_opt(const SvaeShape& s, const SvaeConfig& c, const Plan& p,
                                       const SvaeDecoderParams& dp, SvaeDecoderParams& g, const float* T,
                                       const float* zs, float z_scale, float* dz, cudaStream_t st) {
    if (s.B == 0) return SVAE_OK;
    const int F = p.F;
    const long ldT = (long)(F + 1) * p.Hp;
    SVAE_TRY(coord_param_grad_opt(F, T, s.B, s.H, p.Hp, g.coord_w, g.coord_b, st));
    if (s.Z == 0 || dp.latent_w == nullptr) return SVAE_OK;
    const int split_b = s.B >= 512 ? 8 : (s.B >= 128 ? 2 : 1);
    // dWz (H,Z) += T_0^T zs
    SgemmArgs w{};
    w.A = T; w.sAm = 1; w.sAk = ldT;
    w.B = zs; w.sBk = s.Z; w.sBn = 1;
    w.C = g.latent_w; w.ldc = s.Z;
    w.M = s.H; w.N = s.Z; w.K = s.B;
    w.accumulate = 1; w.split_k = split_b;
    SVAE_TRY(sgemm(w, st));
    if (dz) {
        // dz (B,Z) = z_scale * T_0 Wz
        SVAE_CUDA(cudaMemsetAsync(dz, 0, (size_t)s.B * s.Z * sizeof(float), st));
        SgemmArgs x{};
        x.A = T; x.sAm = ldT; x.sAk = 1;
        x.B = dp.latent_w; x.sBk = s.Z; x.sBn = 1;
        x.C = dz; x.ldc = s.Z;
        x.M = s.B; x.N = s.Z; x.K = s.H;
        x.alpha = z_scale; x.accumulate = 1;
        SVAE_TRY(sgemm(x, st));
    }
    if (!c.bilinear) return SVAE_OK;
    for (int i = 0; i < F; ++i) {
        const float* Ti = T + (size_t)(1 + i) * p.Hp;              // dW_eff[b][n,i] = T[b][1+i][n]
        // dWb[n,i,j] += sum_b T[b][1+i][n] zs[b,j]
        SgemmArgs wb{};
        wb.A = Ti; wb.sAm = 1; wb.sAk = ldT;
        wb.B = zs; wb.sBk = s.Z; wb.sBn = 1;
        wb.C = g.bilinear_w + (size_t)i * s.Z; wb.ldc = (long)F * s.Z;
        wb.M = s.H; wb.N = s.Z; wb.K = s.B;
        wb.accumulate = 1; wb.split_k = split_b;
        SVAE_TRY(sgemm(wb, st));
        if (dz) {
            // dz[b,j] += z_scale * sum_n T[b][1+i][n] Wb[n,i,j]
            SgemmArgs xb{};
            xb.A = Ti; xb.sAm = ldT; xb.sAk = 1;
            xb.B = dp.bilinear_w + (size_t)i * s.Z; xb.sBk = (long)F * s.Z; xb.sBn = 1;
            xb.C = dz; xb.ldc = s.Z;
            xb.M = s.B; xb.N = s.Z; xb.K = s.H;
            xb.alpha = z_scale; xb.accumulate = 1;
            SVAE_TRY(sgemm(xb, st));
        }
    }
    return SVAE_OK;
}

// gradients that flow through S: coord layer, latent_linear, and dz (B,Z)
static int first_layer_param_grads(const SvaeShape& s, const SvaeConfig& c, const Plan& p,
                                   const SvaeDecoderParams& dp, SvaeDecoderParams& g, const float* S,
                                   const float* img, const float* zs, float z_scale, int explicit_x, float* dz,
                                   cudaStream_t st_main, Side* side = nullptr) {
    (void)c;
    if (s.B == 0) return SVAE_OK;
    // parameter gradients (accumulated, consumed by Adam only) on the auxiliary stream; dz, which the latent
    // backward waits for, on the caller's
    SVAE_TRY(side_fork(side, st_main));
    cudaStream_t st = side_or(side, st_main);
    SVAE_TRY(coord_param_grad(S, img, s.B, s.H, p.Hp, explicit_x, g.coord_w, g.coord_b, st));
    if (s.Z > 0 && dp.latent_w != nullptr) {
        // dWz (H,Z) += S_s^T zs
        SgemmArgs w{};
        w.A = S; w.sAm = 1; w.sAk = 3L * p.Hp;
        w.B = zs; w.sBk = s.Z; w.sBn = 1;
        w.C = g.latent_w; w.ldc = s.Z;
        w.M = s.H; w.N = s.Z; w.K = s.B;
        w.accumulate = 1;
        w.split_k = s.B >= 512 ? 8 : (s.B >= 128 ? 2 : 1);     // few output tiles: spread the long K over more CTAs
        SVAE_TRY(sgemm(w, st));
        if (dz) {
            // dz (B,Z) = z_scale * S_s Wz
            SgemmArgs x{};
            x.A = S; x.sAm = 3L * p.Hp; x.sAk = 1;
            x.B = dp.latent_w; x.sBk = s.Z; x.sBn = 1;
            x.C = dz; x.ldc = s.Z;
            x.M = s.B; x.N = s.Z; x.K = s.H;
            x.alpha = z_scale;
            if (s.H >= 256) {      // narrow output, long K: split-K (see encoder_head_forward)
                SVAE_CUDA(cudaMemsetAsync(dz, 0, (size_t)s.B * s.Z * sizeof(float), st_main));
                x.split_k = 4;
            }
            SVAE_TRY(sgemm(x, st_main));
        }
    }
    return SVAE_OK;
}

__global__ void finalize_stats_k(float* stats, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) stats[b * 3 + 2] = stats[b * 3 + 0] - stats[b * 3 + 1];
}
// the same with the three column sums over the call's images written to sums[0..2] (sums[3] = 0): one block, so the
// result is written, not accumulated, and its summation order is fixed
__global__ void __launch_bounds__(1024) finalize_stats_sum_k(float* stats, int B, float* __restrict__ sums) {
    __shared__ float red[3][32];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float lp = stats[b * 3 + 0], kl = stats[b * 3 + 1];
        const float el = lp - kl;
        stats[b * 3 + 2] = el;
        a0 += lp; a1 += kl; a2 += el;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        a0 = lane < nw ? red[0][lane] : 0.f;
        a1 = lane < nw ? red[1][lane] : 0.f;
        a2 = lane < nw ? red[2][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { sums[0] = a0; sums[1] = a1; sums[2] = a2; sums[3] = 0.f; }
    }
}

template <typename T>
static int step_impl(const SvaeShape& s, const SvaeConfig& c, const Plan& p, const SvaeDecoderParams& dp,
                     const SvaeEncoderParams& qp, const SvaeStepInputs& in, const SvaeStepOutputs& out,
                     SvaeDecoderParams* gd, SvaeEncoderParams* gq, char* ws, cudaStream_t st) {
    DecoderCtx<T> d{&s, &c, &p, ws, st};
    const bool train = (gd != nullptr);
    d.bwd_follows = train;
    const float* x_enc = in.y_enc ? in.y_enc : in.y;
    float* zo = d.f(p.zo);
    constexpr bool kFast = !std::is_same<T, float>::value;
    const bool enc_on_tc = kFast || p.tc3;
    EncTc enc_tc{&s, &p, ws, st};
    if (enc_on_tc) SVAE_TRY(encoder_forward_tc(enc_tc, c.activation, qp, x_enc, zo));
    else SVAE_TRY(encoder_forward_impl(s, c.activation, qp, x_enc, zo, d.f(p.enc_acts), st, c.resid != 0));
    float* lat = out.latent ? out.latent : d.f(p.lat);
    LatentRng rng;
    rng.step = in.rng_step; rng.seed = in.rng_seed; rng.image_offset = in.rng_image_offset;
    const float* eps = in.eps ? in.eps : d.f(p.eps);
    SVAE_TRY(latent_forward(s, c, zo, in.eps, in.theta_offset, lat, d.f(p.img), d.f(p.zs), out.stats, st, rng, d.f(p.eps)));
    SVAE_TRY(latent_projection(s, p, dp, d.f(p.zs), d.f(p.hz), st));
    if (c.bilinear) SVAE_TRY(bilinear_weights(s, p, dp, d.f(p.zs), d.f(p.weff), st));
    if (!std::is_same<T, float>::value) SVAE_TRY(prepare_bf16_weights(s, p, dp, ws, st));
    if (p.tc3) SVAE_TRY(prepare_split_weights(s, p, dp, ws, st));
    if (train && kFast) SVAE_CUDA(cudaMemsetAsync(ws + p.S, 0, (size_t)s.B * p.K1 * p.Hp * sizeof(float), st));
    for (int b0 = 0; b0 < s.B; b0 += p.chunk) {
        const int nb = (s.B - b0 < p.chunk) ? (s.B - b0) : p.chunk;
        const DecoderCtx<T>& dc = d;
        SVAE_TRY(decoder_chunk_forward<T>(dc, dp, b0, nb, in.grid, nullptr, out.y_hat));
        SVAE_TRY(likelihood(s, c, b0, nb, dc.logits(), in.y, in.ctf, in.mask, out.stats, train ? dc.g_logits() : nullptr,
                            dc.st));
        if (train) SVAE_TRY(decoder_chunk_backward<T>(dc, dp, *gd, b0, nb, in.grid, nullptr, nullptr));
    }
    Side* side = nullptr;
    if (train && (kFast || p.tc3)) SVAE_TRY(side_get(&side));
    if (out.stats_sum != nullptr) finalize_stats_sum_k<<<1, 1024, 0, st>>>(out.stats, s.B, out.stats_sum);
    else finalize_stats_k<<<ceil_div(s.B, 128), 128, 0, st>>>(out.stats, s.B);
    SVAE_LAUNCH_CHECK();
    if (train) {
        if (p.opt) {
            SVAE_TRY(first_layer_param_grads_opt(s, c, p, dp, *gd, d.f(p.S), d.f(p.zs), c.z_scale,
                                                 s.Z > 0 ? d.f(p.dz) : nullptr, st));
            SVAE_TRY(latent_coord_grad(p.F, s.B, s.H, p.Hp, d.l0_w(dp), p.w_img_stride, d.f(p.S), d.f(p.img),
                                       d.f(p.dlat), st));
            SVAE_TRY(latent_backward(s, c, nullptr, p.Hp, d.f(p.img), nullptr, s.Z > 0 ? d.f(p.dz) : nullptr, zo,
                                     eps, d.f(p.g_zo), st, d.f(p.dlat)));
        } else {
            SVAE_TRY(first_layer_param_grads(s, c, p, dp, *gd, d.f(p.S), d.f(p.img), d.f(p.zs), c.z_scale, 0,
                                             s.Z > 0 ? d.f(p.dz) : nullptr, st, side));
            SVAE_TRY(latent_backward(s, c, d.f(p.S), p.Hp, d.f(p.img), dp.coord_w, s.Z > 0 ? d.f(p.dz) : nullptr, zo,
                                     eps, d.f(p.g_zo), st));
        }
        if (in.decoder_grads_event != nullptr) {
            // every decoder gradient is final once the parameter-gradient chain above is: it was forked from `st`
            // after the last decoder kernel, so its stream (or `st` itself without one) is where the event goes
            cudaStream_t s_ev = (side != nullptr && side->forked) ? side->aux : st;
            SVAE_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(in.decoder_grads_event), s_ev));
        }
        if (gq) {
            enc_tc.side = side;
            if (enc_on_tc) SVAE_TRY(encoder_backward_tc(enc_tc, c.activation, qp, x_enc, d.f(p.g_zo), *gq, d.f(p.enc_scratch)));
            else SVAE_TRY(encoder_backward_impl(s, c.activation, qp, x_enc, d.f(p.enc_acts), d.f(p.g_zo), *gq, nullptr,
                                                d.f(p.enc_scratch), st, c.resid != 0));
        }
    }
    SVAE_TRY(side_join(side, st));      // nothing is left running on the auxiliary stream
    return SVAE_OK;
}

// module-level decoder (explicit coordinates)
template <typename T>
static int decoder_impl(const SvaeShape& s, const SvaeConfig& c, const Plan& p, const SvaeDecoderParams& dp,
                        const float* x, const float* z, float* y_hat, const float* g_y, SvaeDecoderParams* gd,
                        float* g_x, float* g_z, char* ws, cudaStream_t st);

__global__ void scale_rows_k(const float* z, float* zs, long n, float sc) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) zs[i] = z[i] * sc;
}
// g_o = g_y * d(y_hat)/d(o) from the stored logits
__global__ void logit_grad_k(const float* __restrict__ o, const float* __restrict__ g_y, float* __restrict__ g_o,
                             long n, int C, int softplus) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sg = 1.f / (1.f + expf(-o[i]));
    float dv = sg * (1.f - sg);
    if (softplus && (i % C) == 0) dv *= 1.f / (1.f + expf(-sg));
    g_o[i] = g_y[i] * dv;
}

template <typename T>
static int decoder_impl(const SvaeShape& s, const SvaeConfig& c, const Plan& p, const SvaeDecoderParams& dp,
                        const float* x, const float* z, float* y_hat, const float* g_y, SvaeDecoderParams* gd,
                        float* g_x, float* g_z, char* ws, cudaStream_t st) {
    DecoderCtx<T> d{&s, &c, &p, ws, st};
    d.bwd_follows = (gd != nullptr);
    if (s.B == 0) return SVAE_OK;
    if (s.Z > 0) {
        scale_rows_k<<<ceil_div((long)s.B * s.Z, 256), 256, 0, st>>>(z, d.f(p.zs), (long)s.B * s.Z, 1.f);
        SVAE_LAUNCH_CHECK();
    }
    SVAE_TRY(latent_projection(s, p, dp, d.f(p.zs), d.f(p.hz), st));
    if (c.bilinear) SVAE_TRY(bilinear_weights(s, p, dp, d.f(p.zs), d.f(p.weff), st));
    if (!std::is_same<T, float>::value) SVAE_TRY(prepare_bf16_weights(s, p, dp, ws, st));
    if (p.tc3) SVAE_TRY(prepare_split_weights(s, p, dp, ws, st));
    for (int b0 = 0; b0 < s.B; b0 += p.chunk) {
        const int nb = (s.B - b0 < p.chunk) ? (s.B - b0) : p.chunk;
        SVAE_TRY(decoder_chunk_forward<T>(d, dp, b0, nb, nullptr, x, y_hat));
        if (gd) {
            const long n = (long)nb * s.P * s.C;
            logit_grad_k<<<ceil_div(n, 256), 256, 0, st>>>(d.logits(), g_y + (size_t)b0 * s.P * s.C, d.g_logits(), n, s.C,
                                                           c.softplus);
            SVAE_LAUNCH_CHECK();
            SVAE_TRY(decoder_chunk_backward<T>(d, dp, *gd, b0, nb, nullptr, x, g_x));
        }
    }
    if (gd) {
        if (p.opt) SVAE_TRY(first_layer_param_grads_opt(s, c, p, dp, *gd, d.f(p.S), d.f(p.zs), 1.f, g_z, st));
        else SVAE_TRY(first_layer_param_grads(s, c, p, dp, *gd, d.f(p.S), nullptr, d.f(p.zs), 1.f, 1, g_z, st));
    }
    return SVAE_OK;
}

}  // namespace svae

using namespace svae;

extern "C" {

int svae_version(void) { return 100; }

int svae_last_error(char* buf, int n) {
    if (buf && n > 0) {
        strncpy(buf, g_err, (size_t)n - 1);
        buf[n - 1] = '\0';
    }
    return (int)strlen(g_err);
}

unsigned long long svae_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int svae_device_sm_count(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SVAE_ECUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SVAE_ECUDA;
    return sms;
}

int svae_workspace_bytes(const SvaeShape* shape, const SvaeConfig* cfg, size_t* bytes) {
    SVAE_REQUIRE(shape && cfg && bytes, SVAE_EINVAL, "null argument");
    SVAE_TRY(validate(*shape, *cfg));
    Plan p;
    SVAE_TRY(make_plan(*shape, *cfg, p));
    *bytes = p.total;
    return SVAE_OK;
}

int svae_encoder_forward(const SvaeShape* shape, int activation, const SvaeEncoderParams* params, const float* x,
                         float* out, float* acts, void* stream) {
    SVAE_REQUIRE(shape && params && x && out && acts, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(shape->Lq >= 1 && shape->Lq <= SVAE_MAX_LAYERS, SVAE_EINVAL, "bad encoder depth %d", shape->Lq);
    if (shape->B == 0) return SVAE_OK;
    return encoder_forward_impl(*shape, activation & 0xff, *params, x, out, acts, (cudaStream_t)stream,
                                (activation & SVAE_ENC_RESID) != 0);
}

int svae_encoder_backward(const SvaeShape* shape, int activation, const SvaeEncoderParams* params, const float* x,
                          const float* acts, float* g_out, SvaeEncoderParams* grads, float* g_x, float* scratch,
                          void* stream) {
    SVAE_REQUIRE(shape && params && x && acts && g_out && grads && scratch, SVAE_EINVAL, "null argument");
    if (shape->B == 0) return SVAE_OK;
    return encoder_backward_impl(*shape, activation & 0xff, *params, x, acts, g_out, *grads, g_x, scratch,
                                 (cudaStream_t)stream, (activation & SVAE_ENC_RESID) != 0);
}

static int decoder_shape_check(const SvaeShape& s, const SvaeConfig& c) {
    SVAE_REQUIRE(s.B >= 0 && s.P > 0 && s.H > 0 && s.L >= 1 && s.L <= SVAE_MAX_LAYERS && s.C >= 1 && s.C <= 4,
                 SVAE_EINVAL, "bad decoder shape");
    SVAE_REQUIRE(c.activation >= 0 && c.activation <= 3, SVAE_EINVAL, "unknown activation %d", c.activation);
    SVAE_REQUIRE(!c.bilinear || s.Z > 0, SVAE_EINVAL, "bilinear needs a latent (models.py:73-75)");
    return SVAE_OK;
}

int svae_decoder_forward(const SvaeShape* shape, const SvaeConfig* cfg, const SvaeDecoderParams* params,
                         const float* x, const float* z, float* y_hat, void* workspace, size_t workspace_bytes,
                         void* stream) {
    SVAE_REQUIRE(shape && cfg && params && x && y_hat && workspace, SVAE_EINVAL, "null argument");
    SVAE_TRY(decoder_shape_check(*shape, *cfg));
    SVAE_REQUIRE(!cfg->bilinear || (params->bilinear_w && z), SVAE_EINVAL, "cfg.bilinear needs bilinear_w and z");
    Plan p;
    SVAE_TRY(make_plan(*shape, *cfg, p));
    SVAE_REQUIRE(workspace_bytes >= p.total, SVAE_ENOSPACE, "workspace %zu < %zu bytes", workspace_bytes, p.total);
    if (use_fast(*cfg))
        return decoder_impl<__nv_bfloat16>(*shape, *cfg, p, *params, x, z, y_hat, nullptr, nullptr, nullptr, nullptr,
                                           (char*)workspace, (cudaStream_t)stream);
    return decoder_impl<float>(*shape, *cfg, p, *params, x, z, y_hat, nullptr, nullptr, nullptr, nullptr,
                               (char*)workspace, (cudaStream_t)stream);
}

int svae_decoder_backward(const SvaeShape* shape, const SvaeConfig* cfg, const SvaeDecoderParams* params,
                          const float* x, const float* z, const float* g_y, SvaeDecoderParams* grads, float* g_x,
                          float* g_z, void* workspace, size_t workspace_bytes, void* stream) {
    SVAE_REQUIRE(shape && cfg && params && x && g_y && grads && workspace, SVAE_EINVAL, "null argument");
    SVAE_TRY(decoder_shape_check(*shape, *cfg));
    SVAE_REQUIRE(!cfg->bilinear || (params->bilinear_w && grads->bilinear_w && z), SVAE_EINVAL,
                 "cfg.bilinear needs bilinear_w (parameters and gradients) and z");
    Plan p;
    SVAE_TRY(make_plan(*shape, *cfg, p));
    SVAE_REQUIRE(workspace_bytes >= p.total, SVAE_ENOSPACE, "workspace %zu < %zu bytes", workspace_bytes, p.total);
    if (use_fast(*cfg))
        return decoder_impl<__nv_bfloat16>(*shape, *cfg, p, *params, x, z, nullptr, g_y, grads, g_x, g_z,
                                           (char*)workspace, (cudaStream_t)stream);
    return decoder_impl<float>(*shape, *cfg, p, *params, x, z, nullptr, g_y, grads, g_x, g_z, (char*)workspace,
                               (cudaStream_t)stream);
}

int svae_step(const SvaeShape* shape, const SvaeConfig* cfg, const SvaeDecoderParams* dec,
              const SvaeEncoderParams* enc, const SvaeStepInputs* in, const SvaeStepOutputs* out,
              SvaeDecoderParams* dec_grads, SvaeEncoderParams* enc_grads, void* workspace, size_t workspace_bytes,
              void* stream) {
    SVAE_REQUIRE(shape && cfg && dec && enc && in && out && workspace, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(in->grid && in->y && out->stats, SVAE_EINVAL, "grid, y and stats are required");
    SVAE_REQUIRE(in->eps || in->rng_step, SVAE_EINVAL, "eps or rng_step (in-kernel draw) is required");
    SVAE_TRY(validate(*shape, *cfg));
    SVAE_REQUIRE(shape->Lq >= 1 && shape->Lq <= SVAE_MAX_LAYERS, SVAE_EINVAL, "bad encoder depth %d", shape->Lq);
    SVAE_REQUIRE((shape->k_ctf > 0) == (in->ctf != nullptr), SVAE_EINVAL, "k_ctf and the ctf pointer disagree");
    SVAE_REQUIRE((dec_grads == nullptr) == (enc_grads == nullptr), SVAE_EINVAL,
                 "pass both gradient structs or neither");
    SVAE_REQUIRE(!cfg->bilinear || (dec->bilinear_w && (dec_grads == nullptr || dec_grads->bilinear_w)), SVAE_EINVAL,
                 "cfg.bilinear needs bilinear_w (parameters and gradients)");
    if (shape->B == 0) return SVAE_OK;
    Plan p;
    SVAE_TRY(make_plan(*shape, *cfg, p));
    SVAE_REQUIRE(workspace_bytes >= p.total, SVAE_ENOSPACE, "workspace %zu < %zu bytes", workspace_bytes, p.total);
    if (use_fast(*cfg))
        return step_impl<__nv_bfloat16>(*shape, *cfg, p, *dec, *enc, *in, *out, dec_grads, enc_grads,
                                        (char*)workspace, (cudaStream_t)stream);
    return step_impl<float>(*shape, *cfg, p, *dec, *enc, *in, *out, dec_grads, enc_grads, (char*)workspace,
                            (cudaStream_t)stream);
}

int svae_adam_step(float* param, float* grad, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                   float eps, int t, int zero_grad, void* stream) {
    SVAE_REQUIRE(param && grad && m && v, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(t >= 1, SVAE_EINVAL, "Adam step count starts at 1");
    return adam(param, grad, m, v, n, lr, beta1, beta2, eps, t, zero_grad, nullptr, (cudaStream_t)stream);
}

int svae_adam_tick(int32_t* t_dev, float* bias_corr_dev, float beta1, float beta2, void* stream) {
    SVAE_REQUIRE(t_dev && bias_corr_dev, SVAE_EINVAL, "null argument");
    return adam_tick(t_dev, bias_corr_dev, beta1, beta2, (cudaStream_t)stream);
}

int svae_adam_step_graph(float* param, float* grad, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                         float eps, const float* bias_corr_dev, int zero_grad, void* stream) {
    SVAE_REQUIRE(param && grad && m && v && bias_corr_dev, SVAE_EINVAL, "null argument");
    return adam(param, grad, m, v, n, lr, beta1, beta2, eps, 0, zero_grad, bias_corr_dev, (cudaStream_t)stream);
}

int svae_gather_rows(const float* src, const int64_t* index, float* dst, int64_t n_rows, int64_t row_len,
                     void* stream) {
    if (n_rows == 0) return SVAE_OK;       // an empty slice (a rank with no images in a ragged minibatch)
    SVAE_REQUIRE(src && index && dst, SVAE_EINVAL, "null argument");
    return gather_rows(src, index, dst, n_rows, row_len, (cudaStream_t)stream);
}

int svae_rotate_bicubic(const float* src, float* dst, const double* inv_affine, const int32_t* mode, int B, int n_rows,
                        int n_cols, int channels, int quantize_u8, void* stream) {
    if (B == 0) return SVAE_OK;
    SVAE_REQUIRE(src && dst && inv_affine && mode, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(B >= 0 && n_rows > 0 && n_cols > 0 && channels >= 1 && channels <= 4, SVAE_EINVAL, "bad image shape");
    return rotate_bicubic(src, dst, inv_affine, mode, B, n_rows, n_cols, channels, quantize_u8, (cudaStream_t)stream);
}

// Python's round(x, 15): correctly rounded decimal -> nearest double (glibc printf/strtod are exact)
static double round15(double x) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.15f", x);
    return strtod(buf, nullptr);
}

int svae_rotation_matrices(const double* angles_deg, int B, int n_rows, int n_cols, double* inv_affine, int32_t* mode) {
    if (B == 0) return SVAE_OK;
    SVAE_REQUIRE(angles_deg && inv_affine && mode && B >= 0, SVAE_EINVAL, "null argument");
    const double cx = n_cols / 2.0, cy = n_rows / 2.0;
    for (int b = 0; b < B; ++b) {
        double ang = fmod(angles_deg[b], 360.0);              // Python's float %: result takes the divisor's sign
        if (ang < 0.0) ang += 360.0;
        double* m = inv_affine + (size_t)b * 6;
        mode[b] = 0;
        if (ang == 0.0) mode[b] = 1;
        else if (ang == 180.0) mode[b] = 2;
        else if (n_rows == n_cols && ang == 90.0) mode[b] = 3;
        else if (n_rows == n_cols && ang == 270.0) mode[b] = 4;
        const double a = -(ang * (M_PI / 180.0));               // -math.radians(angle)
        m[0] = round15(cos(a)); m[1] = round15(sin(a)); m[3] = round15(-sin(a)); m[4] = round15(cos(a));
        m[2] = (m[0] * -cx + m[1] * -cy + 0.0) + cx;            // transform(-cx, -cy) + centre  (PIL/Image.py)
        m[5] = (m[3] * -cx + m[4] * -cy + 0.0) + cy;
    }
    return SVAE_OK;
}

int svae_ctf_filter(const double* params, int n_particles, int n, int m, double scale, float* out, void* stream) {
    SVAE_REQUIRE(n_particles >= 0 && n > 0 && m > 0 && scale > 0, SVAE_EINVAL, "bad CTF kernel shape");
    if (n_particles == 0) return SVAE_OK;
    SVAE_REQUIRE(params && out, SVAE_EINVAL, "null argument");
    return ctf_filter(params, n_particles, n, m, scale, out, (cudaStream_t)stream);
}

int svae_sm_clock_probe(float* out_mhz, void* stream) {
    SVAE_REQUIRE(out_mhz != nullptr, SVAE_EINVAL, "null argument");
    return clock_probe(out_mhz, (cudaStream_t)stream);
}

int svae_resid_linear_forward(const float* x, const float* w, const float* b, float* out, int rows, int n, int activation,
                              void* stream) {
    SVAE_REQUIRE(x && w && b && out && rows >= 0 && n > 0, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(activation >= 0 && activation <= 3, SVAE_EINVAL, "unknown activation %d", activation);
    if (rows == 0) return SVAE_OK;
    SgemmArgs a{};
    a.A = x; a.sAm = n; a.sAk = 1;
    a.B = w; a.sBk = 1; a.sBn = n;
    a.C = out; a.ldc = n;
    a.M = rows; a.N = n; a.K = n;
    a.bias = b; a.act = activation;
    a.add = x; a.ld_add = n;                         // the skip connection rides in the GEMM epilogue
    return sgemm(a, (cudaStream_t)stream);
}

int svae_resid_linear_backward(const float* x, const float* w, const float* out, const float* g_out, float* g_pre,
                               float* g_x, float* g_w, float* g_b, int rows, int n, int activation, void* stream) {
    SVAE_REQUIRE(x && w && out && g_out && g_pre && g_x && g_w && g_b && rows >= 0 && n > 0, SVAE_EINVAL, "null argument");
    SVAE_REQUIRE(activation >= 0 && activation <= 3, SVAE_EINVAL, "unknown activation %d", activation);
    if (rows == 0) return SVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // g_pre = g_out .* act'(out)   (an identity "GEMM" is avoided: scale rows through the dX epilogue of K = 0)
    SVAE_TRY(act_backward(out, g_out, g_pre, (long)rows * n, activation, st));
    // dW (n, n) += g_pre^T x ; db += colsum g_pre
    SgemmArgs wg{};
    wg.A = g_pre; wg.sAm = 1; wg.sAk = n;
    wg.B = x; wg.sBk = n; wg.sBn = 1;
    wg.C = g_w; wg.ldc = n;
    wg.M = n; wg.N = n; wg.K = rows;
    wg.accumulate = 1; wg.split_k = rows >= 2048 ? 8 : (rows >= 512 ? 4 : 1);
    SVAE_TRY(sgemm(wg, st));
    SVAE_TRY(col_sum<float>(g_pre, rows, n, n, g_b, st));
    // g_x = g_pre W + g_pre
    SgemmArgs xg{};
    xg.A = g_pre; xg.sAm = n; xg.sAk = 1;
    xg.B = w; xg.sBk = n; xg.sBn = 1;
    xg.C = g_x; xg.ldc = n;
    xg.M = rows; xg.N = n; xg.K = n;
    xg.add = g_pre; xg.ld_add = n;
    return sgemm(xg, st);
}

int svae_gemm_dx_moments(int rows, int H, int Hp, const void* delta, int ldd, const void* W, int ldw, int activation,
                         const float* grid, const float* img, const float* coord_w, const float* hz, float* S, int P,
                         void* stream) {
    SVAE_REQUIRE(delta && W && grid && img && coord_w && hz && S, SVAE_EINVAL, "null argument");
    TcMoments m;
    m.grid = grid; m.img = img; m.coord_w = coord_w; m.hz = hz; m.S = S; m.P = P; m.b0 = 0;
    return tc_dx_moments(rows, H, Hp, delta, ldd, W, ldw, activation, m, (cudaStream_t)stream);
}

int svae_gemm_dw_top(int rows, int H, int Hp, const void* h_top, const void* h_prev, int activation, const float* g_o,
                     int C, const float* out_w, float* d_out_w, float* d_out_b, float* d_b, float* dW, void* delta_out,
                     void* stream) {
    SVAE_REQUIRE(h_top && h_prev && g_o && out_w && d_out_w && d_out_b && dW, SVAE_EINVAL, "null argument");
    TcTop t;
    t.g_o = g_o; t.C = C; t.out_w = out_w; t.d_out_w = d_out_w; t.d_out_b = d_out_b; t.d_b = d_b;
    return tc_dw_top(rows, H, Hp, h_top, h_prev, activation, t, dW, H, delta_out, (cudaStream_t)stream);
}

int svae_gemm_bf16(int mode, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias,
                   const void* aux, int ldaux, int activation, void* out, int ldo, void* stream) {
    SVAE_REQUIRE(A && W && out, SVAE_EINVAL, "null argument");
    return tc_gemm(mode, M, N, K, A, lda, W, ldw, bias, N, aux, ldaux, activation, out, ldo, (cudaStream_t)stream);
}

}  // extern "C"
