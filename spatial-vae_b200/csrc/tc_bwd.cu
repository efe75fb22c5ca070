// Fused backward kernels of the decoder hidden stack on sm_100a tensor cores (tcgen05 / TMEM / TMA), the part of
// loss.backward() (train_mnist.py:147-148) that autograd spends on SpatialGenerator.layers (models.py:77-87,126).
//
//   dx_red_kernel   delta_0 = (delta_1 W_1) .* act'(h_0) is never written: the GEMM is computed TRANSPOSED
//                   (D[n, row] = sum_j W_1[j, n] delta_1[row, j], hidden column n along the TMEM lanes, pixel rows
//                   along the TMEM columns), so every epilogue thread owns ONE hidden column and walks the pixel rows
//                   serially.  It recomputes h_0[row, n] = act(Wc[n,0] x' + Wc[n,1] y' + hz[b, n]) in registers
//                   (the first layer, models.py:104-124), multiplies by act', and accumulates the three per-image
//                   column moments S[b, {1, c0, c1}, n] with no cross-thread communication.
//                   W_1 stays resident in shared memory (128 KB per CTA at Hp <= 512); only delta_1 streams.
//                   Replaces: the dX GEMM's store of delta_0, image_col_reduce's pass over it and the h_0 aux stream.
//
//   dw_xf_kernel    dW_l += delta_l^T h_{l-1} for the TOP hidden layer with delta_l produced on the fly:
//                   delta[row, m] = (sum_c g_o[row, c] W_o[c, m]) * act'(h_l[row, m]) is computed by transform warps
//                   IN the shared-memory operand stage that TMA filled with h_l, then fed to tcgen05.mma; the same
//                   warps accumulate db_l, dW_o and db_o, and the transformed stage is written out once by TMA as
//                   delta_l for the dX GEMM.  Replaces out_backward's two passes over (rows x Hp) matrices.
#include "tc_ptx.cuh"

namespace svae {

namespace {

constexpr int KB = 64;                       // reduction elements per pipeline stage
constexpr int BOX_BYTES = 64 * 64 * 2;       // one 64 x 64 bf16 TMA box (MN-major operand block)
constexpr int OP_BYTES = 128 * KB * 2;       // 16 KB: 128 (rows | columns) x 64 of one operand of one CTA
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quadrant
constexpr int BW_THREADS = 128 + 32 * EPI_WARPS;

// ---------------------------------------------------------------------------------------------------------------
// dx_red_kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr int DR_TILE_N = 256;               // hidden columns per CTA-pair tile (MMA M; 128 TMEM lanes per CTA)
constexpr int DR_TILE_R = 256;               // pixel rows per tile (MMA N; each CTA stages 128 of them)
constexpr int DR_SLAB_BYTES = 8 * OP_BYTES;  // resident weight slab: 128 columns x 512 reduction elements
constexpr int DR_TABLE_BYTES = DR_TILE_R * 8;      // per row pair: (g0, g0') and (g1, g1') as two float2 arrays

#ifndef DR_STAGES_RES
#define DR_STAGES_RES 4
#endif
__host__ __device__ constexpr int dr_stages(bool resident) { return resident ? DR_STAGES_RES : 6; }
__host__ __device__ constexpr int dr_stage_bytes(bool resident) { return resident ? OP_BYTES : 2 * OP_BYTES; }
__host__ __device__ constexpr int dr_off_ring(bool resident) { return resident ? DR_SLAB_BYTES : 0; }
__host__ __device__ constexpr int dr_off_table(bool resident) {
    return dr_off_ring(resident) + dr_stages(resident) * dr_stage_bytes(resident);
}
__host__ __device__ constexpr int dr_off_bars(bool resident) { return dr_off_table(resident) + 2 * DR_TABLE_BYTES; }
__host__ __device__ constexpr int dr_smem_bytes(bool resident) { return dr_off_bars(resident) + 256 + 1024; }
static_assert(dr_smem_bytes(true) <= 232448 && dr_smem_bytes(false) <= 232448, "shared memory budget");

struct DxRedParams {
    int M, H, Hp;                // pixel rows of this pass, hidden width, padded hidden width
    int n_tiles, r_tiles, k_blocks;
    int n_groups;                // CTA pairs per column tile: pair i owns column tile i % n_tiles and a contiguous
                                 // range of row tiles, so its running per-image sums are flushed once per image
    int P, b0;                   // pixel rows per image; index of the pass's first image in img / hz / S
    const float* grid;           // (P, 2)
    const float* img;            // (B, 4) cos, sin, dx0, dx1
    const float* coord_w;        // (H, 2)
    const float* hz;             // (B, Hp)  W_z z + b_c
    float* S;                    // (B, 3, Hp)
};

// Per-thread state of the moment epilogue: hidden column n of image b_cur.  The first-layer pre-activation is affine
// in the raw grid coordinate, a = A g0 + B g1 + C with A = Wc0 cos + Wc1 sin, B = Wc1 cos - Wc0 sin,
// C = Wc0 dx0 + Wc1 dx1 + hz[b, n]  (train_mnist.py:50-59,70-74 folded into models.py:104), so the row table only
// holds the grid.  Two rows are processed per packed (f32x2) instruction.
struct MomentState {
    float2 A, B, C;              // splats
    float2 s, m0, m1;            // running sums of the even / odd row of each pair
};

template <int ACT>
__device__ __forceinline__ void moment_pair(float2 v, float2 gx, float2 gy, MomentState& st) {
    const float2 a = __ffma2_rn(st.A, gx, __ffma2_rn(st.B, gy, st.C));
    float2 h, g;
    h.x = act_const<ACT>(a.x); h.y = act_const<ACT>(a.y);
    if (ACT == SVAE_ACT_TANH) {
        g = __ffma2_rn(make_float2(-h.x, -h.y), h, make_float2(1.f, 1.f));
    } else {
        g.x = act_deriv_const<ACT>(h.x); g.y = act_deriv_const<ACT>(h.y);
    }
    const float2 d = __fmul2_rn(v, g);
    st.s = __fadd2_rn(st.s, d);
    st.m0 = __ffma2_rn(gx, d, st.m0);
    st.m1 = __ffma2_rn(gy, d, st.m1);
}
template <int ACT>
__device__ __forceinline__ void moment_one(float v, float gx, float gy, MomentState& st) {
    const float h = act_const<ACT>(fmaf(st.A.x, gx, fmaf(st.B.x, gy, st.C.x)));
    const float d = v * act_deriv_const<ACT>(h);
    st.s.x += d;
    st.m0.x = fmaf(gx, d, st.m0.x);
    st.m1.x = fmaf(gy, d, st.m1.x);
}

template <int ACT, bool RESIDENT>
__global__ void __launch_bounds__(BW_THREADS, 1)
dx_red_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD, const DxRedParams p) {
    constexpr int STAGES = dr_stages(RESIDENT);
    constexpr int STAGE_BYTES = dr_stage_bytes(RESIDENT);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // pointer arithmetic (not an integer round trip) keeps the shared state space: LDS / STS instead of generic LD / ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* slab = smem;                                            // RESIDENT only
    uint8_t* ring = smem + dr_off_ring(RESIDENT);
    float2* table = reinterpret_cast<float2*>(smem + dr_off_table(RESIDENT));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + dr_off_bars(RESIDENT));
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], slab_full, then the tmem base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader_cta = (cta_rank == 0);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
    const uint32_t slabfull = smem_u32(bars + 2 * STAGES + 4);
    const uint32_t full0_leader = map_to_cta(full0, 0), tempty0_leader = map_to_cta(tempty0, 0);
    const uint32_t slabfull_leader = map_to_cta(slabfull, 0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmD);
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS); }
        mbar_init(slabfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this pair's work: column tile nt, row tiles [rt0, rt1)
    const int pair = blockIdx.x / 2;
    const int nt = pair % p.n_tiles, grp = pair / p.n_tiles;
    const int rt0 = (int)((long)grp * p.r_tiles / p.n_groups), rt1 = (int)((long)(grp + 1) * p.r_tiles / p.n_groups);
    const int n0 = nt * DR_TILE_N + (int)cta_rank * 128;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: the weight slab once (RESIDENT), then the delta rows (and weight blocks) per K block =====
        if (RESIDENT && rt0 < rt1) {
            if (is_leader_cta) mbar_expect_tx(slabfull, 2 * p.k_blocks * OP_BYTES);
            for (int kb = 0; kb < p.k_blocks; ++kb)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    tma_load_2d_pair(smem_u32(slab + kb * OP_BYTES + i * BOX_BYTES), &tmW, slabfull_leader, n0 + i * 64,
                                     kb * KB);
        }
        int stage = 0; uint32_t phase = 0;
        for (int rt = rt0; rt < rt1; ++rt) {
            const int r0 = rt * DR_TILE_R + (int)cta_rank * 128;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t fb = full0_leader + 8 * stage;
                if (is_leader_cta) mbar_expect_tx(full0 + 8 * stage, 2 * STAGE_BYTES);
                uint8_t* st = ring + stage * STAGE_BYTES;
                if (!RESIDENT) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        tma_load_2d_pair(smem_u32(st + OP_BYTES + i * BOX_BYTES), &tmW, fb, n0 + i * 64, kb * KB);
                }
                tma_load_2d_pair(smem_u32(st), &tmD, fb, kb * KB, r0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader_cta) {
        // ===== MMA issuer: D[n, row] (+)= W[j, n]^T (MN-major A) x delta[row, j] (K-major B) =====
        constexpr uint32_t idesc = make_idesc(1, 0, DR_TILE_N, DR_TILE_R);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        if (RESIDENT && rt0 < rt1) { mbar_wait(slabfull, 0); tc_fence_after(); }
        for (int rt = rt0; rt < rt1; ++rt) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * DR_TILE_R;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sb = smem_u32(ring + stage * STAGE_BYTES);
                const uint32_t sa = RESIDENT ? smem_u32(slab + kb * OP_BYTES) : sb + OP_BYTES;
#pragma unroll
                for (int k = 0; k < KB / 16; ++k) {
                    const uint64_t ad = make_smem_desc(sa + k * 2048, BOX_BYTES, 1024);
                    const uint64_t bd = make_smem_desc(sb + k * 32, 0, 1024);
                    umma_bf16_pair(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit_pair(empty0 + 8 * stage);
                if (kb == p.k_blocks - 1) umma_commit_pair(tfull0 + 8 * acc);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4 && rt0 < rt1) {
        // ===== epilogue: thread = one hidden column n (TMEM lane), 128 of the tile's 256 pixel rows (TMEM columns) =====
        const int q = warp & 3;                        // TMEM lane quadrant
        const int ch = (warp - 4) >> 2;                // which half of the tile's rows
        const int etid = threadIdx.x - 128;            // 0..255
        const int n = n0 + q * 32 + lane;
        const bool n_ok = n < p.Hp, n_live = n < p.H;
        const float wc0 = n_live ? __ldg(p.coord_w + 2 * n) : 0.f;
        const float wc1 = n_live ? __ldg(p.coord_w + 2 * n + 1) : 0.f;
        int acc = 0; uint32_t acc_phase = 0;
        MomentState ms;
        int b_cur, next_b;                             // image of the running sums; first row of the next image
        auto load_image = [&](int b) {
            b_cur = b;
            next_b = (b + 1) * p.P;
            if (next_b >= p.M) next_b = 0x7fffffff;
            const float4 im = __ldg(reinterpret_cast<const float4*>(p.img) + p.b0 + b);      // cos, sin, dx0, dx1
            const float hzb = n_live ? __ldg(p.hz + (size_t)(p.b0 + b) * p.Hp + n) : 0.f;
            const float A = fmaf(wc0, im.x, wc1 * im.y), B = fmaf(wc1, im.x, -wc0 * im.y);
            const float C = fmaf(wc0, im.z, fmaf(wc1, im.w, hzb));
            ms.A = make_float2(A, A); ms.B = make_float2(B, B); ms.C = make_float2(C, C);
            ms.s = ms.m0 = ms.m1 = make_float2(0.f, 0.f);
        };
        auto flush = [&]() {
            if (n_ok) {
                float* sp = p.S + (size_t)(p.b0 + b_cur) * 3 * p.Hp + n;
                atomicAdd(sp, ms.s.x + ms.s.y);
                atomicAdd(sp + p.Hp, ms.m0.x + ms.m0.y);
                atomicAdd(sp + 2 * p.Hp, ms.m1.x + ms.m1.y);
            }
        };
        {
            const int R0 = rt0 * DR_TILE_R + ch * 128;
            load_image((R0 < p.M ? R0 : p.M - 1) / p.P);
        }
        // grid coordinate of the row this thread contributes to the table of the next tile (loaded one tile ahead)
        auto grid_of = [&](int rt) {
            const int row = rt * DR_TILE_R + etid;
            float2 g = make_float2(0.f, 0.f);
            if (rt < rt1 && row < p.M) g = __ldg(reinterpret_cast<const float2*>(p.grid) + row % p.P);
            return g;
        };
        float2 g_mine = grid_of(rt0);
        uint32_t tile_it = 0;
        for (int rt = rt0; rt < rt1; ++rt, ++tile_it) {
            // row table, pair-interleaved: tabx[i] = (g0[2i], g0[2i+1]), taby[i] = (g1[2i], g1[2i+1])
            float* tabf = reinterpret_cast<float*>(table + (tile_it & 1) * DR_TILE_R);
            tabf[(etid >> 1) * 2 + (etid & 1)] = g_mine.x;
            tabf[DR_TILE_R + (etid >> 1) * 2 + (etid & 1)] = g_mine.y;
            g_mine = grid_of(rt + 1);
            const float2* tabx = reinterpret_cast<const float2*>(tabf) + ch * 64;
            const float2* taby = tabx + DR_TILE_R / 2;
            const int R0 = rt * DR_TILE_R + ch * 128;                  // first pixel row of this thread's columns
            epi_bar_sync_all(32 * EPI_WARPS);                          // the table is complete
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            if (R0 < p.M) {
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * DR_TILE_R + ch * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int rc = R0 + 32 * c;
                    if (rc >= p.M) break;
                    uint32_t v[32];
                    tmem_ld32(t_row + 32 * c, v);
                    while (next_b <= rc) { flush(); load_image(b_cur + 1); }
                    tmem_ld_wait();
                    if (next_b >= rc + 32) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            moment_pair<ACT>(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                             tabx[16 * c + j], taby[16 * c + j], ms);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (rc + j == next_b) { flush(); load_image(b_cur + 1); }
                            const float2 gx = tabx[16 * c + (j >> 1)], gy = taby[16 * c + (j >> 1)];
                            moment_one<ACT>(__uint_as_float(v[j]), (j & 1) ? gx.y : gx.x, (j & 1) ? gy.y : gy.x, ms);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0_leader + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        flush();
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

template <int ACT, bool RESIDENT>
int launch_dx_red(const CUtensorMap& w, const CUtensorMap& d, const DxRedParams& p, int grid, cudaStream_t st) {
    static bool configured = false;
    auto kern = dx_red_kernel<ACT, RESIDENT>;
    if (!configured) {
        SVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dr_smem_bytes(RESIDENT)));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = dr_smem_bytes(RESIDENT);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    SVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, w, d, p));
    return SVAE_OK;
}

template <bool RESIDENT>
int launch_dx_red_act(int act, const CUtensorMap& w, const CUtensorMap& d, const DxRedParams& p, int grid,
                      cudaStream_t st) {
    switch (act) {
        case SVAE_ACT_TANH: return launch_dx_red<SVAE_ACT_TANH, RESIDENT>(w, d, p, grid, st);
        case SVAE_ACT_LEAKYRELU: return launch_dx_red<SVAE_ACT_LEAKYRELU, RESIDENT>(w, d, p, grid, st);
        case SVAE_ACT_RELU: return launch_dx_red<SVAE_ACT_RELU, RESIDENT>(w, d, p, grid, st);
        case SVAE_ACT_SIGMOID: return launch_dx_red<SVAE_ACT_SIGMOID, RESIDENT>(w, d, p, grid, st);
        default: set_error("tc_dx_moments: unknown activation %d", act); return SVAE_EINVAL;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// dw_xf_kernel: dW (M x N, fp32) += delta^T h_prev over the data rows, delta built in shared memory from h_top
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_TILE_M = 256;                                // delta columns per CTA-pair tile (128 TMEM lanes per CTA)
constexpr int DW_TILE_N = 512;                                // h_prev columns per tile: two N = 256 MMAs, all 512 TMEM columns
constexpr int DW_STAGES = 4;
constexpr int DW_G_BYTES = 1024;                              // g_o of the stage's 64 rows (up to 3 channels)
constexpr int DW_STAGE_BYTES = 3 * OP_BYTES + DW_G_BYTES;     // A (delta <- h_top) 16 KB + B (h_prev) 32 KB + g_o, per CTA
constexpr int DW_OFF_BARS = DW_STAGES * DW_STAGE_BYTES;
constexpr int DW_SMEM_BYTES = DW_OFF_BARS + 512 + 1024;
static_assert(DW_SMEM_BYTES <= 232448, "shared memory budget");
constexpr int DW_MAX_C = 3;

struct DwXfParams {
    int M, N, K;                 // dW rows (= delta columns), dW columns (= h_prev columns), data rows
    int m_tiles, n_tiles, k_blocks, k_splits, k_blocks_per_split;
    float* out; int ldo; int vec_red;
    const float* g_o;            // (K, C) gradient w.r.t. the logits
    const float* out_w; int out_w_ld;     // (C, H)
    float* d_out_w;              // (C, H)   += g_o^T h_top
    float* d_out_b;              // (C)      += column sums of g_o
    float* d_b;                  // (H)      += column sums of delta, may be NULL
    int store_delta;             // write delta (K x Hp, bf16) through tmD
};

// The tile is 256 x 512 so that every delta element is transformed by ONE CTA pair per column tile of h_prev and the
// transform warps have two MMAs (1024 tensor-core cycles) per 64-row stage to hide behind.
template <int ACT, int C>
__global__ void __launch_bounds__(BW_THREADS, 1)
dw_xf_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmD, const DwXfParams p) {
    constexpr int STAGES = DW_STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_OFF_BARS);
    // bars: afull[S] (this CTA's h_top boxes), bfull[S] (leader: both CTAs' h_prev boxes), xfull[S] (leader: both CTAs'
    // transform warps), empty[S] (MMA retired + every transform warp's delta store has read the stage), tfull
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader_cta = (cta_rank == 0);
    const uint32_t afull0 = smem_u32(bars), bfull0 = smem_u32(bars + STAGES), xfull0 = smem_u32(bars + 2 * STAGES);
    const uint32_t empty0 = smem_u32(bars + 3 * STAGES), tfull = smem_u32(bars + 4 * STAGES);
    const uint32_t bfull0_leader = map_to_cta(bfull0, 0), xfull0_leader = map_to_cta(xfull0, 0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.store_delta) tma_prefetch_desc(&tmD);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(afull0 + 8 * i, 1);
            mbar_init(bfull0 + 8 * i, 1);
            mbar_init(xfull0 + 8 * i, 2 * EPI_WARPS);
            mbar_init(empty0 + 8 * i, 1 + EPI_WARPS);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.m_tiles * p.n_tiles * p.k_splits;
    const int first_tile = blockIdx.x / 2, tile_stride = gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int stage = 0; uint32_t phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
            const int ks = tile / (p.m_tiles * p.n_tiles), mn = tile % (p.m_tiles * p.n_tiles);
            const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
            const int kb0 = ks * p.k_blocks_per_split, kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            const int m0 = mt * DW_TILE_M + (int)cta_rank * 128;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                uint8_t* st = smem + stage * DW_STAGE_BYTES;
                // g_o rows of the block ride on the same barrier (bulk copy; sizes are multiples of 16 bytes, the last block
                // may read up to 12 bytes past row K-1: callers keep g_o 16-byte padded)
                const long rows_left = (long)p.K - (long)kb * KB;
                const uint32_t g_bytes = (uint32_t)(((rows_left < KB ? rows_left : KB) * C * 4 + 15) & ~15L);
                mbar_expect_tx(afull0 + 8 * stage, OP_BYTES + g_bytes);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    tma_load_2d(smem_u32(st + i * BOX_BYTES), &tmA, afull0 + 8 * stage, m0 + i * 64, kb * KB);
                bulk_load(smem_u32(st + 3 * OP_BYTES), p.g_o + (size_t)kb * KB * C, g_bytes, afull0 + 8 * stage);
                if (is_leader_cta) mbar_expect_tx(bfull0 + 8 * stage, 4 * OP_BYTES);
#pragma unroll
                for (int j = 0; j < 2; ++j)          // MMA j covers h_prev columns [nt 512 + j 256, +256), 128 per CTA
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        tma_load_2d_pair(smem_u32(st + OP_BYTES + (2 * j + i) * BOX_BYTES), &tmB, bfull0_leader + 8 * stage,
                                         nt * DW_TILE_N + j * 256 + (int)cta_rank * 128 + i * 64, kb * KB);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader_cta) {
        // ===== MMA issuer: both operands MN-major (64 x 64 boxes, rows = reduction index) =====
        constexpr uint32_t idesc = make_idesc(1, 1, DW_TILE_M, 256);
        int stage = 0; uint32_t phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
            const int ks = tile / (p.m_tiles * p.n_tiles);
            const int kb0 = ks * p.k_blocks_per_split, kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            // the transform warps drained the previous tile's accumulator before they produced this tile's first stage
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bfull0 + 8 * stage, phase);
                mbar_wait(xfull0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * DW_STAGE_BYTES), sb = sa + OP_BYTES;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k) {
                        const uint64_t ad = make_smem_desc(sa + k * 2048, BOX_BYTES, 1024);
                        const uint64_t bd = make_smem_desc(sb + j * 2 * BOX_BYTES + k * 2048, BOX_BYTES, 1024);
                        umma_bf16_pair(tmem_base + j * 256, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                umma_commit_pair(empty0 + 8 * stage);
                if (kb == kb1 - 1) umma_commit_pair(tfull);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== transform warps (delta in place of h_top), then the epilogue of the tile =====
        const int w = warp - 4;                        // 0..7: rows [8w, 8w+8) of every 64-row stage
        const int cchunk = lane & 15;                  // 16-byte chunk (8 columns) of this CTA's 128 delta columns
        const int box = cchunk >> 3, cc = cchunk & 7;
        const int rsub = lane >> 4;                    // rows 8w + rsub + 2i, i = 0..3
        const int q = warp & 3, ch = w >> 2;           // epilogue: TMEM lane quadrant, column half
        int stage = 0; uint32_t phase = 0;
        uint32_t tile_phase = 0;
        int prev_stage = -1;                           // lane 0: stage whose delta store may still be reading
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
            const int ks = tile / (p.m_tiles * p.n_tiles), mn = tile % (p.m_tiles * p.n_tiles);
            const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
            const int kb0 = ks * p.k_blocks_per_split, kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            const int m0 = mt * DW_TILE_M + (int)cta_rank * 128;
            const int mc = m0 + cchunk * 8;                            // first of this thread's 8 delta columns
            const bool sums = (nt == 0);                               // one column tile per delta column does the sums
            const bool store = sums && p.store_delta;
            const bool gsum = sums && mc == 0;                         // db_o rides with delta column 0
            float2 wo[C][4], dwo[C][4], db[4];
            float gs[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gs[c] = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    wo[c][e].x = (mc + 2 * e < p.M) ? __ldg(p.out_w + (size_t)c * p.out_w_ld + mc + 2 * e) : 0.f;
                    wo[c][e].y = (mc + 2 * e + 1 < p.M) ? __ldg(p.out_w + (size_t)c * p.out_w_ld + mc + 2 * e + 1) : 0.f;
                    dwo[c][e] = make_float2(0.f, 0.f);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) db[e] = make_float2(0.f, 0.f);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(afull0 + 8 * stage, phase);
                uint8_t* sa = smem + stage * DW_STAGE_BYTES + box * BOX_BYTES;
                const float* sg = reinterpret_cast<const float*>(smem + stage * DW_STAGE_BYTES + 3 * OP_BYTES);
                uint4 hv[4];
                float g[4][C];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 8 * w + rsub + 2 * i;
                    hv[i] = *reinterpret_cast<const uint4*>(sa + r * 128 + ((cc ^ (r & 7)) << 4));
                    const bool live = (long)kb * KB + r < p.K;
#pragma unroll
                    for (int c = 0; c < C; ++c) g[i][c] = live ? sg[r * C + c] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 8 * w + rsub + 2 * i;
                    const uint32_t hw[4] = {hv[i].x, hv[i].y, hv[i].z, hv[i].w};
                    uint32_t dv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 h = make_float2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u));
                        float2 t = make_float2(0.f, 0.f);
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float2 gc = make_float2(g[i][c], g[i][c]);
                            t = (c == 0) ? __fmul2_rn(gc, wo[c][e]) : __ffma2_rn(gc, wo[c][e], t);
                            if (sums) dwo[c][e] = __ffma2_rn(gc, h, dwo[c][e]);
                        }
                        float2 a;
                        if (ACT == SVAE_ACT_TANH) {
                            a = __ffma2_rn(make_float2(-h.x, -h.y), h, make_float2(1.f, 1.f));
                        } else {
                            a.x = act_deriv_const<ACT>(h.x); a.y = act_deriv_const<ACT>(h.y);
                        }
                        const float2 d = __fmul2_rn(t, a);
                        if (sums) db[e] = __fadd2_rn(db[e], d);
                        dv[e] = pack_bf16(d.x, d.y);
                    }
                    *reinterpret_cast<uint4*>(sa + r * 128 + ((cc ^ (r & 7)) << 4)) = make_uint4(dv[0], dv[1], dv[2], dv[3]);
                    if (gsum) {
#pragma unroll
                        for (int c = 0; c < C; ++c) gs[c] += g[i][c];
                    }
                }
                fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core and to TMA
                __syncwarp();
                if (lane == 0) {
                    if (store) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            tma_store_2d(&tmD, smem_u32(smem + stage * DW_STAGE_BYTES + i * BOX_BYTES + w * 1024), m0 + i * 64,
                                         kb * KB + 8 * w);
                    }
                    tma_store_commit();
                    if (prev_stage >= 0) { tma_store_wait_read<1>(); mbar_arrive(empty0 + 8 * prev_stage); }
                    prev_stage = stage;
                    mbar_arrive_cluster(xfull0_leader + 8 * stage);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (lane == 0 && prev_stage >= 0) { tma_store_wait_read<0>(); mbar_arrive(empty0 + 8 * prev_stage); prev_stage = -1; }
            // ---- column sums: lanes l and l^16 share their columns; the 8 warps are combined through global atomics
            if (sums) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    db[e].x += __shfl_xor_sync(0xffffffffu, db[e].x, 16);
                    db[e].y += __shfl_xor_sync(0xffffffffu, db[e].y, 16);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        dwo[c][e].x += __shfl_xor_sync(0xffffffffu, dwo[c][e].x, 16);
                        dwo[c][e].y += __shfl_xor_sync(0xffffffffu, dwo[c][e].y, 16);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) gs[c] += __shfl_xor_sync(0xffffffffu, gs[c], 16);
                if (lane < 16) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            const int m = mc + 2 * e + hlf;
                            if (m < p.M) {
                                if (p.d_b) atomicAdd(p.d_b + m, hlf ? db[e].y : db[e].x);
#pragma unroll
                                for (int c = 0; c < C; ++c)
                                    atomicAdd(p.d_out_w + (size_t)c * p.out_w_ld + m, hlf ? dwo[c][e].y : dwo[c][e].x);
                            }
                        }
                    }
                    if (gsum) {
#pragma unroll
                        for (int c = 0; c < C; ++c) atomicAdd(p.d_out_b + c, gs[c]);
                    }
                }
            }
            // ---- epilogue: fp32 partial dW tile -> global atomics; warp = (lane quadrant q, column half ch)
            mbar_wait(tfull, tile_phase);
            tile_phase ^= 1;
            tc_fence_after();
            const int m = mt * DW_TILE_M + (int)cta_rank * 128 + q * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c = ch * 256; c < ch * 256 + 256; c += 32) {
                const int n = nt * DW_TILE_N + c;
                if (n >= p.N) break;
                uint32_t v[32];
                tmem_ld32(t_row + c, v);
                tmem_ld_wait();
                if (m < p.M) {
                    float* op = p.out + (size_t)m * p.ldo + n;
                    if (p.vec_red && n + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"l"(op + 4 * j), "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                                           "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3])) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < p.N) atomicAdd(op + j, __uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();     // ordered before the next tile's first xfull arrive, which the MMA issuer waits on
        }
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

template <int ACT, int C>
int launch_dw_xf(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& d, const DwXfParams& p, int grid,
                 cudaStream_t st) {
    static bool configured = false;
    auto kern = dw_xf_kernel<ACT, C>;
    if (!configured) {
        SVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = DW_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    SVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b, d, p));
    return SVAE_OK;
}

template <int C>
int launch_dw_xf_act(int act, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& d, const DwXfParams& p,
                     int grid, cudaStream_t st) {
    switch (act) {
        case SVAE_ACT_TANH: return launch_dw_xf<SVAE_ACT_TANH, C>(a, b, d, p, grid, st);
        case SVAE_ACT_LEAKYRELU: return launch_dw_xf<SVAE_ACT_LEAKYRELU, C>(a, b, d, p, grid, st);
        case SVAE_ACT_RELU: return launch_dw_xf<SVAE_ACT_RELU, C>(a, b, d, p, grid, st);
        case SVAE_ACT_SIGMOID: return launch_dw_xf<SVAE_ACT_SIGMOID, C>(a, b, d, p, grid, st);
        default: set_error("tc_dw_top: unknown activation %d", act); return SVAE_EINVAL;
    }
}

}  // namespace

int tc_dx_moments(int rows, int H, int Hp, const void* delta, int ldd, const void* W, int ldw, int act,
                  const TcMoments& r, cudaStream_t st) {
    if (rows <= 0) return SVAE_OK;
    SVAE_REQUIRE(Hp % 64 == 0 && H <= Hp, SVAE_EINVAL, "tc_dx_moments: the padded width must be a multiple of 64");
    SVAE_REQUIRE(r.grid && r.img && r.coord_w && r.hz && r.S && r.P > 0, SVAE_EINVAL, "tc_dx_moments: null argument");
    DxRedParams p{};
    p.M = rows; p.H = H; p.Hp = Hp;
    p.n_tiles = ceil_div(Hp, DR_TILE_N);
    p.r_tiles = ceil_div(rows, DR_TILE_R);
    p.k_blocks = Hp / KB;
    p.P = r.P; p.b0 = r.b0; p.grid = r.grid; p.img = r.img; p.coord_w = r.coord_w; p.hz = r.hz; p.S = r.S;
    CUtensorMap mw, md;
    SVAE_TRY(make_map(&mw, W, Hp, Hp, ldw, 64, 64));            // W[j, n]: boxes of 64 n x 64 j
    SVAE_TRY(make_map(&md, delta, rows, Hp, ldd, 64, 128));     // delta[row, j]: boxes of 64 j x 128 rows
    // CTA pairs = n_tiles column tiles x n_groups contiguous ranges of row tiles
    const int max_pairs = sm_count() / 2;
    SVAE_REQUIRE(p.n_tiles <= max_pairs, SVAE_EINVAL, "tc_dx_moments: hidden width %d too large", Hp);
    p.n_groups = max_pairs / p.n_tiles;
    if (p.n_groups > p.r_tiles) p.n_groups = p.r_tiles;
    const int pairs = p.n_groups * p.n_tiles;
    // the weight slab of the pair's column tile stays in shared memory when it fits
    if (p.k_blocks * OP_BYTES <= DR_SLAB_BYTES) return launch_dx_red_act<true>(act, mw, md, p, 2 * pairs, st);
    return launch_dx_red_act<false>(act, mw, md, p, 2 * pairs, st);
}

int tc_dw_top(int rows, int H, int Hp, const void* h_top, const void* h_prev, int act, const TcTop& t, float* dW, int ldo,
              void* delta_out, cudaStream_t st) {
    if (rows <= 0) return SVAE_OK;
    SVAE_REQUIRE(Hp % 64 == 0 && H <= Hp, SVAE_EINVAL, "tc_dw_top: the padded width must be a multiple of 64");
    SVAE_REQUIRE(t.C >= 1 && t.C <= DW_MAX_C, SVAE_EINVAL, "tc_dw_top: n_out = %d not supported (1..%d)", t.C, DW_MAX_C);
    SVAE_REQUIRE(t.g_o && t.out_w && t.d_out_w && t.d_out_b && dW, SVAE_EINVAL, "tc_dw_top: null argument");
    DwXfParams p{};
    p.M = H; p.N = H; p.K = rows;
    p.m_tiles = ceil_div(H, DW_TILE_M); p.n_tiles = ceil_div(H, DW_TILE_N);
    p.k_blocks = ceil_div(rows, KB);
    const int mn = p.m_tiles * p.n_tiles, pairs_max = sm_count() / 2;
    int splits = pairs_max / mn;
    if (splits < 1) splits = 1;
    if (splits > p.k_blocks) splits = p.k_blocks;
    p.k_blocks_per_split = ceil_div(p.k_blocks, splits);
    p.k_splits = ceil_div(p.k_blocks, p.k_blocks_per_split);
    p.out = dW; p.ldo = ldo;
    p.vec_red = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
    p.g_o = t.g_o; p.out_w = t.out_w; p.out_w_ld = H; p.d_out_w = t.d_out_w; p.d_out_b = t.d_out_b; p.d_b = t.d_b;
    p.store_delta = delta_out != nullptr;
    CUtensorMap ma, mb, md;
    memset(&md, 0, sizeof(md));
    SVAE_TRY(make_map(&ma, h_top, rows, Hp, Hp, 64, 64));
    SVAE_TRY(make_map(&mb, h_prev, rows, Hp, Hp, 64, 64));
    if (delta_out) SVAE_TRY(make_map(&md, delta_out, rows, Hp, Hp, 64, 8));
    const int tiles = mn * p.k_splits;
    const int grid = 2 * (tiles < pairs_max ? tiles : pairs_max);
    switch (t.C) {
        case 1: return launch_dw_xf_act<1>(act, ma, mb, md, p, grid, st);
        case 2: return launch_dw_xf_act<2>(act, ma, mb, md, p, grid, st);
        default: return launch_dw_xf_act<3>(act, ma, mb, md, p, grid, st);
    }
}

}  // namespace svae
