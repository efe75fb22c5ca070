// fp32 FFMA GEMM with fused epilogues.  Used by the encoder (InferenceNetwork, models.py:46-54),
// by the small per-image projections of the decoder (latent_linear, models.py:111) and, in
// PARITY precision, by the decoder hidden layers themselves (models.py:82,126) and their
// backward.  Register-tiled, shared-memory staged, global->register prefetch of the next K tile.
#include "common.cuh"

namespace svae {

template <int BM, int BN, int BK, int TM, int TN, bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(SgemmArgs g, int k_chunk) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile/thread mismatch");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be float4 multiples");
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_chunk;
    const int kend = min(g.K, kbeg + k_chunk);

    constexpr int A_PER = BM * BK / NT, B_PER = BN * BK / NT;
    float ra[A_PER], rb[B_PER];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int e = tid + i * NT;
            int m, k;
            if (A_KCONTIG) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
            int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < g.M && gk < kend) ? __ldg(g.A + (long)gm * g.sAm + (long)gk * g.sAk) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int e = tid + i * NT;
            int n, k;
            if (B_KCONTIG) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
            int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < g.N && gk < kend) ? __ldg(g.B + (long)gk * g.sBk + (long)gn * g.sBn) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int e = tid + i * NT;
            int m, k;
            if (A_KCONTIG) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int e = tid + i * NT;
            int n, k;
            if (B_KCONTIG) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
            Bs[k][n] = rb[i];
        }
    };

    // micro-tile rows: TM/4 groups of 4 rows, group q at ty*4 + q*(BM/(TM/4)); same for columns
    constexpr int GM = TM / 4, GN = TN / 4;
    constexpr int SM_STRIDE = BM / GM, SN_STRIDE = BN / GN;

    if (kbeg < kend) load_tiles(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int q = 0; q < GM; ++q) {
                float4 v = *reinterpret_cast<const float4*>(&As[k][ty * 4 + q * SM_STRIDE]);
                a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < GN; ++q) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + q * SN_STRIDE]);
                b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool first_split = (blockIdx.z == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * 4 + (i / 4) * SM_STRIDE + (i % 4);
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * 4 + (j / 4) * SN_STRIDE + (j % 4);
            if (gn >= g.N) continue;
            float v = g.alpha * acc[i][j];
            if (g.bias != nullptr && first_split) v += __ldg(g.bias + gn);
            if (g.add != nullptr) v += __ldg(g.add + (long)gm * g.ld_add + gn);
            if (g.act >= 0) v = act_apply<false>(g.act, v);
            if (g.dact >= 0) v *= act_deriv_from_out(g.dact, __ldg(g.dsrc + (long)gm * g.ld_dsrc + gn));
            float* c = g.C + (long)gm * g.ldc + gn;
            if (g.accumulate) atomicAdd(c, v); else *c = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_cfg(const SgemmArgs& a, cudaStream_t st) {
    constexpr int BK = 16;
    int split = a.split_k < 1 ? 1 : a.split_k;
    int k_chunk = (int)round_up(ceil_div(a.K, split), BK);
    split = ceil_div(a.K, k_chunk);
    if (split < 1) split = 1;
    dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM), split);
    const bool ak = (a.sAk == 1), bk = (a.sBk == 1);
    if (!ak && a.sAm != 1) { set_error("sgemm: A has no unit stride"); return SVAE_EINVAL; }
    if (!bk && a.sBn != 1) { set_error("sgemm: B has no unit stride"); return SVAE_EINVAL; }
    constexpr int NT = (BM / TM) * (BN / TN);
    if (ak && bk)        sgemm_kernel<BM, BN, BK, TM, TN, true, true><<<grid, NT, 0, st>>>(a, k_chunk);
    else if (ak && !bk)  sgemm_kernel<BM, BN, BK, TM, TN, true, false><<<grid, NT, 0, st>>>(a, k_chunk);
    else if (!ak && bk)  sgemm_kernel<BM, BN, BK, TM, TN, false, true><<<grid, NT, 0, st>>>(a, k_chunk);
    else                 sgemm_kernel<BM, BN, BK, TM, TN, false, false><<<grid, NT, 0, st>>>(a, k_chunk);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

int sgemm(const SgemmArgs& a_in, cudaStream_t st) {
    SgemmArgs a = a_in;
    if (a.M <= 0 || a.N <= 0) return SVAE_OK;
    if (a.K <= 0) { set_error("sgemm: K must be positive"); return SVAE_EINVAL; }
    if (a.split_k > 1) {
        if (a.act >= 0 || a.dact >= 0 || a.add != nullptr) { set_error("sgemm: split-K cannot carry a nonlinear epilogue"); return SVAE_EINVAL; }
        a.accumulate = 1;
    }
    const long big_tiles = (long)ceil_div(a.M, 128) * ceil_div(a.N, 128) * (a.split_k > 1 ? a.split_k : 1);
    if (big_tiles >= 120) return launch_cfg<128, 128, 8, 8>(a, st);
    // few 64x64 tiles: these GEMMs are latency bound (one CTA per SM walking K), so cut the tile to get more,
    // shorter CTAs in flight
    const long mid_tiles = (long)ceil_div(a.M, 64) * ceil_div(a.N, 64) * (a.split_k > 1 ? a.split_k : 1);
    if (mid_tiles >= 296) return launch_cfg<64, 64, 4, 4>(a, st);
    return launch_cfg<32, 32, 4, 4>(a, st);
}

}  // namespace svae
