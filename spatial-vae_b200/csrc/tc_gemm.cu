// bf16 tensor-core GEMMs for the decoder hidden layers (models.py:82,126 and their backward) on
// sm_100a: TMA (cp.async.bulk.tensor) feeds a shared-memory ring, one elected thread issues
// tcgen05.mma with the fp32 accumulator in TMEM (two 128x256 accumulator stages = all 512 columns),
// epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue while the next tile's
// MMAs run.  Persistent CTAs, one per SM.  CG = 2 pairs two SMs (cta_group::2, cluster of 2): the pair
// computes a 256 x 256 tile, each CTA stages its own 128 A rows and HALF of the B tile, so the weight /
// operand traffic through L2 and shared memory per FLOP halves.
//
//   mode 0  FWD : out[M,N]  = act(A[M,K] W[N,K]^T + bias)              A K-major,  B K-major
//                 (+ optional fused output layer: o[m,c] += sum_n h[m,n] W_o[c,n])
//                 (RES: ResidLinear, out = act(A W^T + bias + R) with the residual tile R streamed into the epilogue)
//   mode 1  DX  : out[M,N]  = (A[M,K] W[K,N]) .* act'(aux[M,N])        A K-major,  B MN-major
//   mode 2  DW  : outf[M,N] += A[Kr,M]^T Bm[Kr,N]  (split over Kr)     A MN-major, B MN-major
#include "tc_ptx.cuh"

namespace svae {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB: this CTA's 128 rows (or 128 M-columns) x 64 K
constexpr int BOX_BYTES = 64 * 64 * 2;       // one 64x64 bf16 TMA box (MN-major operands)
constexpr int MAX_DOT_C = 3;                  // output channels the fused output-layer dot supports
#ifndef TC_STAGES_WIDE
#define TC_STAGES_WIDE 4
#endif
constexpr int EPI_BLOCK_BYTES = 128 * 128;    // one 128-row x 128-byte epilogue block (SWIZZLE_128B)
// Epilogue warp-groups per CTA (TC_FWD_EPI_GROUPS, forward pair kernel only).  Two groups were measured in rounds 1
// and 2 (fwd at C2: 397 us against 385-390 us with one group), as was a 5-stage operand ring (TC_STAGES_WIDE: 385 vs
// 390 us) and an L2-resident problem (same TFLOP/s): the kernel is bound by neither epilogue latency, nor TMA
// latency, nor HBM -- it sits at 80 % of what cuBLAS sustains under the 1 kW power cap -- so one group and four
// stages, the configuration validated on hardware, are kept.
#ifndef TC_FWD_EPI_GROUPS
#define TC_FWD_EPI_GROUPS 1
#endif
__host__ __device__ constexpr int epi_groups(int cg, int mode) { return (cg == 2 && mode == 0) ? TC_FWD_EPI_GROUPS : 1; }
__host__ __device__ constexpr int num_threads(int cg, int mode) { return 128 + 128 * epi_groups(cg, mode); }
__host__ __device__ constexpr int b_stage_bytes(int cg) { return (BN / cg) * BK * 2; }        // 32 KB | 16 KB
__host__ __device__ constexpr int stage_bytes(int cg) { return A_STAGE_BYTES + b_stage_bytes(cg); }
// Layout index of a kernel variant: its MODE, or 3 for the forward GEMM with a residual stream (mode 0 + the aux
// staging blocks of mode 1).
constexpr int LAYOUT_FWD_RES = 3;
// ring depth: whatever the staging blocks leave (mode 1 needs aux blocks too)
__host__ __device__ constexpr int stages_of(int cg, int mode) {
    return cg == 2 ? ((mode == 0 || mode == 2) ? TC_STAGES_WIDE : 4) : 3;
}
// shared-memory map after the operand ring:
//   2 output staging blocks per epilogue group | (mode 1) 2 aux blocks per group |
//   tables: mode 0: 2 x (bias[BN] + W_o[3][BN]) floats | barriers
__host__ __device__ constexpr int off_out_stage(int cg, int mode) { return stages_of(cg, mode) * stage_bytes(cg); }
__host__ __device__ constexpr int off_aux_stage(int cg, int mode) {
    return off_out_stage(cg, mode) + (mode == 2 ? 0 : 2 * epi_groups(cg, mode) * EPI_BLOCK_BYTES);
}
__host__ __device__ constexpr int off_tables(int cg, int mode) {
    return off_aux_stage(cg, mode) + ((mode == 1 || mode == LAYOUT_FWD_RES) ? 2 * epi_groups(cg, mode) * EPI_BLOCK_BYTES : 0);
}
__host__ __device__ constexpr int table_bytes(int mode) {
    return (mode == 0 || mode == LAYOUT_FWD_RES) ? 2 * (1 + MAX_DOT_C) * BN * 4 : 0;
}
__host__ __device__ constexpr int off_bars(int cg, int mode) { return off_tables(cg, mode) + table_bytes(mode); }
// The dynamic shared memory is declared 1024-byte aligned; the pair dX kernel has no room for alignment slack
// (it traps if the base ever comes back misaligned), the others keep 1 KB of slack and align by hand.
__host__ __device__ constexpr int align_slack(int cg, int mode) {
    return (cg == 2 && mode == 1 && epi_groups(cg, mode) == 2) ? 0 : 1024;
}
__host__ __device__ constexpr int smem_bytes(int cg, int mode) { return off_bars(cg, mode) + 256 + align_slack(cg, mode); }
static_assert(smem_bytes(1, 0) <= 232448 && smem_bytes(1, 1) <= 232448 && smem_bytes(2, 0) <= 232448 &&
              smem_bytes(2, 1) <= 232448 && smem_bytes(2, 2) <= 232448 && smem_bytes(1, LAYOUT_FWD_RES) <= 232448 &&
              smem_bytes(2, LAYOUT_FWD_RES) <= 232448, "shared memory budget");

struct TcParams {
    int M, N, K;              // modes 0/1: rows, Hp, Hp;  mode 2: out rows, out cols, reduction rows
    int m_tiles, n_tiles, k_blocks, k_splits, k_blocks_per_split;
    const float* bias; int bias_n;
    const __nv_bfloat16* aux; int ldaux;
    int act;
    void* out; int ldo;
    int vec_red;              // mode 2: 16-byte aligned rows -> red.global.add.v4.f32
    int out_f32;              // modes 0/1: fp32 output (and fp32 aux) instead of bf16
    int res;                  // mode 0: a bf16 residual tile is added before the activation (ResidLinear)
    int raw;                  // mode 0, fp32 output: K is split over CTAs and the RAW partial sums are added to `out`
                              // with atomics (no bias, no activation): the caller zeroes `out` and finishes the layer
    // mode 0, optional: fused output layer (models.py:84): o_accum[m, c] += sum_n h[m,n] * out_w[c, n]
    const float* out_w; int out_w_ld; int dot_c; float* o_accum;
};

// MODE: 0 fwd, 1 dX, 2 dW.  ACT: activation of the epilogue (compile time so the per-element code
// is branch free).  DOTC: output channels of the fused output-layer dot product (mode 0), 0 = none.
// CG: 1 = one CTA per 128 x 256 tile; 2 = CTA pair (cluster of 2, cta_group::2) per 256 x 256 tile.
// OUT32: modes 0/1 write fp32 (and read an fp32 aux matrix) instead of bf16: used by the encoder, whose
// fp32 GEMMs run as three bf16 MMAs on hi/lo splits of the operands (error-compensated, ~fp32 accuracy).
// RES: mode 0, bf16 only: out = act(acc + bias + R), R (M x N, bf16) streamed by TMA into the epilogue like dX's aux.
template <int MODE, int ACT, int DOTC, int CG, bool OUT32, bool RES = false>
__global__ void __launch_bounds__(num_threads(CG, RES ? LAYOUT_FWD_RES : MODE), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, const TcParams p) {
    static_assert(!RES || (MODE == 0 && !OUT32 && DOTC == 0), "the residual stream exists for the plain bf16 forward");
    constexpr int LM = RES ? LAYOUT_FWD_RES : MODE;      // shared-memory layout of this variant
    constexpr bool AUX = (MODE == 1) || RES;             // an M x N tile is streamed into the epilogue
    constexpr int STAGES = stages_of(CG, LM);
    constexpr int B_STAGE_BYTES = b_stage_bytes(CG);
    constexpr int STAGE_BYTES = stage_bytes(CG);
    constexpr int BN_CTA = BN / CG;                // B-tile rows (N) staged by this CTA
    constexpr int TILE_M = BM * CG;                // rows of the output tile computed by the CTA (pair)
    constexpr int EG = epi_groups(CG, LM);         // epilogue warp-groups (4 warps each)
    constexpr int NT = num_threads(CG, LM);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles (same offset in both CTAs of a pair)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared state space (LDS/STS, not generic LD/ST)
    if (align_slack(CG, LM) == 0 && smem != smem_raw) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    float* s_tab = reinterpret_cast<float*>(smem + off_tables(CG, LM));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off_bars(CG, LM));
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], aux_full[2*EG], then the tmem base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * EG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool is_leader_cta = (cta_rank == 0);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
    const uint32_t auxfull0 = smem_u32(bars + 2 * STAGES + 4);
    // barriers that live in the leader CTA of a pair (operand "full", accumulator "empty")
    const uint32_t full0_leader = (CG == 2) ? map_to_cta(full0, 0) : full0;
    const uint32_t tempty0_leader = (CG == 2) ? map_to_cta(tempty0, 0) : tempty0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (MODE != 2) tma_prefetch_desc(&tmOut);
        if (AUX) tma_prefetch_desc(&tmAux);
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 4 * EG * CG);    // every epilogue warp of the CTA (pair)
        }
        for (int i = 0; i < 2 * EG; ++i) mbar_init(auxfull0 + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair(smem_u32(tmem_slot), 512); else tmem_alloc(smem_u32(tmem_slot), 512);
    }
    constexpr bool fuse_dot = (MODE == 0) && (DOTC > 0);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) { __syncwarp(); cluster_sync_all(); }   // peer barriers initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.m_tiles * p.n_tiles * p.k_splits;
    const int first_tile = blockIdx.x / CG, tile_stride = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (every CTA: its own A rows and its share of B) =====
        int stage = 0; uint32_t phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
            const int ks = tile / (p.m_tiles * p.n_tiles);
            const int mn = tile % (p.m_tiles * p.n_tiles);
            const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
            const int kb0 = ks * p.k_blocks_per_split;
            const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            const int a0 = mt * TILE_M + (int)cta_rank * BM;          // first A row / M column of this CTA
            const int b0 = nt * BN + (int)cta_rank * BN_CTA;          // first B row / N column of this CTA
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t fb = full0_leader + 8 * stage;
                if (is_leader_cta) mbar_expect_tx(full0 + 8 * stage, STAGE_BYTES * CG);
                const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                const uint32_t sb = smem_u32(smem_b + stage * B_STAGE_BYTES);
                auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
                    if (CG == 2) tma_load_2d_pair(dst, m, fb, c0, c1); else tma_load_2d(dst, m, fb, c0, c1);
                };
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < BM / 64; ++i) load(sa + i * BOX_BYTES, &tmA, a0 + i * 64, kb * BK);
                } else {
                    load(sa, &tmA, kb * BK, a0);
                }
                if (MODE == 0) {
                    load(sb, &tmB, kb * BK, b0);
                } else {
#pragma unroll
                    for (int i = 0; i < BN_CTA / 64; ++i) load(sb + i * BOX_BYTES, &tmB, b0 + i * 64, kb * BK);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader_cta) {
        // ===== MMA issuer (leader CTA of the pair) =====
        constexpr uint32_t idesc = make_idesc(MODE == 2 ? 1 : 0, MODE == 0 ? 0 : 1, TILE_M, BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
            const int ks = tile / (p.m_tiles * p.n_tiles);
            const int kb0 = ks * p.k_blocks_per_split;
            const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                const uint32_t sb = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // K-major: 16 bf16 = 32 bytes along the swizzled 128-byte row; 8-row groups 1024 B apart.
                    // MN-major: 16 k-rows = 2048 bytes; 64-element MN blocks BOX_BYTES apart (LBO).
                    const uint64_t ad = (MODE == 2) ? make_smem_desc(sa + k * 2048, BOX_BYTES, 1024)
                                                    : make_smem_desc(sa + k * 32, 0, 1024);
                    const uint64_t bd = (MODE == 0) ? make_smem_desc(sb + k * 32, 0, 1024)
                                                    : make_smem_desc(sb + k * 2048, BOX_BYTES, 1024);
                    const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
                    if (CG == 2) umma_bf16_pair(d_tmem, ad, bd, idesc, accum); else umma_bf16(d_tmem, ad, bd, idesc, accum);
                }
                // frees the smem slot (in both CTAs) once the MMAs retire; last K block: accumulator complete
                if (CG == 2) umma_commit_pair(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
                if (kb == kb1 - 1) { if (CG == 2) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc); }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (swizzled smem -> TMA store | reductions) =====
        const int eg = (warp - 4) >> 2;        // epilogue warp-group of this warp
        const int q = warp & 3;                // TMEM lane quadrant of this warp
        const int row = q * 32 + lane;         // row of this CTA's 128-row slab owned by this thread
        const int gtid = threadIdx.x - 128 - eg * 128;        // 0..127 within the group
        const int etid = threadIdx.x - 128;                   // 0..128*EG-1 over all epilogue threads
        const bool leader = (gtid == 0);
        int acc = 0; uint32_t acc_phase = 0;
        auto release_accumulator = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tempty0_leader + 8 * acc); else mbar_arrive(tempty0 + 8 * acc);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        };
        if (MODE == 2 || (MODE == 0 && OUT32 && p.raw)) {
            // fp32 partial tile (dW, or a K split of an fp32-output forward) -> global atomics; group g takes the
            // 32-column chunks c/32 == g (mod EG)
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
                const int mn = tile % (p.m_tiles * p.n_tiles);
                const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
                const int m = mt * TILE_M + (int)cta_rank * BM + row;
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
                for (int c = eg * 32; c < BN; c += 32 * EG) {
                    const int n = nt * BN + c;
                    if (n >= p.N) break;
                    uint32_t v[32];
                    tmem_ld32(t_row + c, v);
                    tmem_ld_wait();
                    if (m < p.M) {
                        float* op = reinterpret_cast<float*>(p.out) + (size_t)m * p.ldo + n;
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
                release_accumulator();
            }
        } else {
            // 128-byte-wide column blocks: registers -> swizzled staging block -> TMA store (or reduction).
            // Group g takes blocks jb == g (mod EG) of every tile.  dX additionally streams the matching
            // act[l-1] block in by TMA, two of the group's blocks ahead.
            constexpr int BCOLS = OUT32 ? 32 : 64;         // columns per 128-byte-wide staging block
            constexpr int HALVES = OUT32 ? 1 : 2;          // 32-column TMEM loads per block
            uint8_t* out_stage = smem + off_out_stage(CG, LM) + eg * 2 * EPI_BLOCK_BYTES;
            uint8_t* aux_stage = smem + off_aux_stage(CG, LM) + eg * 2 * EPI_BLOCK_BYTES;
            const uint32_t auxfull_g = auxfull0 + 8 * (2 * eg);
            uint32_t blk = 0;                              // this group's running block counter (buffer = blk & 1)
            uint32_t tile_it = 0;                          // tiles processed (parity selects the table copy)
            auto blocks_in_tile = [&](int tile) {
                const int nt = (tile % (p.m_tiles * p.n_tiles)) % p.n_tiles;
                const int left = (p.N - nt * BN) / BCOLS;
                return left < BN / BCOLS ? left : BN / BCOLS;
            };
            // prefetch cursor over this group's aux blocks (leader thread of the group only)
            int pf_tile = first_tile, pf_jb = eg; uint32_t pf_blk = 0;
            auto prefetch_aux = [&]() {
                while (pf_tile < num_tiles && pf_jb >= blocks_in_tile(pf_tile)) { pf_jb = eg; pf_tile += tile_stride; }
                if (pf_tile >= num_tiles) return;
                const int mn = pf_tile % (p.m_tiles * p.n_tiles);
                const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
                const uint32_t bar = auxfull_g + 8 * (pf_blk & 1);
                mbar_expect_tx(bar, EPI_BLOCK_BYTES);
                tma_load_2d(smem_u32(aux_stage + (pf_blk & 1) * EPI_BLOCK_BYTES), &tmAux, bar, nt * BN + pf_jb * BCOLS,
                            mt * TILE_M + (int)cta_rank * BM);
                ++pf_blk;
                pf_jb += EG;
            };
            if (AUX && leader) { prefetch_aux(); prefetch_aux(); }
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++tile_it) {
                const int mn = tile % (p.m_tiles * p.n_tiles);
                const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
                const int m_cta = mt * TILE_M + (int)cta_rank * BM;
                const int m = m_cta + row;
                const int par = tile_it & 1;
                // per-tile tables, double buffered by tile parity (a fast group may already be one tile ahead)
                float* s_bias = s_tab + par * (1 + MAX_DOT_C) * BN;
                float* s_wo = s_bias + BN;
                if (MODE == 0) {
                    for (int i = etid; i < BN; i += 128 * EG) {
                        const int n = nt * BN + i;
                        s_bias[i] = (p.bias != nullptr && n < p.bias_n) ? p.bias[n] : 0.f;
                        if (fuse_dot) {
#pragma unroll
                            for (int c = 0; c < DOTC; ++c)
                                s_wo[c * BN + i] = (n < p.bias_n) ? p.out_w[(size_t)c * p.out_w_ld + n] : 0.f;
                        }
                    }
                    epi_bar_sync_all(128 * EG);
                }
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
                const int nblk = blocks_in_tile(tile);
                float2 dot2[DOTC > 0 ? DOTC : 1];          // even / odd column partial sums of the fused output dot
#pragma unroll
                for (int c = 0; c < (DOTC > 0 ? DOTC : 1); ++c) dot2[c] = make_float2(0.f, 0.f);
#pragma unroll 1
                for (int jb = eg; jb < nblk; jb += EG, ++blk) {
                    const uint32_t buf = blk & 1;
                    uint8_t* ostage = out_stage + buf * EPI_BLOCK_BYTES;
                    const uint8_t* astage = aux_stage + buf * EPI_BLOCK_BYTES;
                    if (AUX) mbar_wait(auxfull_g + 8 * buf, (blk >> 1) & 1);
                    // the TMA store issued two blocks ago must have finished READING this staging buffer
                    if (leader) tma_store_wait_read<1>();
                    epi_bar_sync(eg);
#pragma unroll
                    for (int half = 0; half < HALVES; ++half) {
                        const int tc = jb * BCOLS + half * 32;             // first tile-local column of this chunk
                        uint32_t v[32];
                        tmem_ld32(t_row + tc, v);
                        if (OUT32) {
                            // fp32 in / fp32 out: the block is 32 columns = 8 chunks of 4 floats
                            uint4 auxv[8];
                            if (MODE == 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    auxv[j] = *reinterpret_cast<const uint4*>(astage + row * 128 + ((j ^ (row & 7)) << 4));
                            }
                            tmem_ld_wait();
                            const float* af = reinterpret_cast<const float*>(auxv);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float r[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float x = __uint_as_float(v[4 * j + e]);
                                    // fp32 outputs are the error-compensated (3-term) GEMMs: accurate tanhf / expf here
                                    r[e] = (MODE == 0) ? act_apply<false>(ACT, x + s_bias[tc + 4 * j + e])
                                                       : x * act_deriv_const<ACT>(af[4 * j + e]);
                                }
                                *reinterpret_cast<float4*>(ostage + row * 128 + ((j ^ (row & 7)) << 4)) =
                                    make_float4(r[0], r[1], r[2], r[3]);
                            }
                        } else {
                            uint4 auxv[4];
                            if (AUX) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    auxv[j] = *reinterpret_cast<const uint4*>(astage + row * 128 + (((half * 4 + j) ^ (row & 7)) << 4));
                            }
                            tmem_ld_wait();
                            uint32_t packed[16];
                            if (MODE == 0) {
#pragma unroll
                                for (int j4 = 0; j4 < 8; ++j4) {          // 4 columns at a time: 128-bit table reads
                                    float4 bv = *reinterpret_cast<const float4*>(s_bias + tc + 4 * j4);
                                    if (RES) {                             // + the layer input (skip connection)
                                        const uint32_t* aw = reinterpret_cast<const uint32_t*>(auxv);
                                        const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&aw[2 * j4]);
                                        const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&aw[2 * j4 + 1]);
                                        bv.x += __low2float(r01); bv.y += __high2float(r01);
                                        bv.z += __low2float(r23); bv.w += __high2float(r23);
                                    }
                                    // packed f32x2 arithmetic: half the issue slots for the bias add and the output dot
                                    const float2 a01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j4 + 0]), __uint_as_float(v[4 * j4 + 1])),
                                                                  make_float2(bv.x, bv.y));
                                    const float2 a23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])),
                                                                  make_float2(bv.z, bv.w));
                                    const float2 h01 = make_float2(act_const<ACT>(a01.x), act_const<ACT>(a01.y));
                                    const float2 h23 = make_float2(act_const<ACT>(a23.x), act_const<ACT>(a23.y));
                                    packed[2 * j4] = pack_bf16(h01.x, h01.y);
                                    packed[2 * j4 + 1] = pack_bf16(h23.x, h23.y);
                                    if (fuse_dot) {
#pragma unroll
                                        for (int c = 0; c < DOTC; ++c) {
                                            const float4 wv = *reinterpret_cast<const float4*>(s_wo + c * BN + tc + 4 * j4);
                                            dot2[c] = __ffma2_rn(h01, make_float2(wv.x, wv.y),
                                                                 __ffma2_rn(h23, make_float2(wv.z, wv.w), dot2[c]));
                                        }
                                    }
                                }
                            } else {
                                const uint32_t* aw = reinterpret_cast<const uint32_t*>(auxv);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
                                    const float d0 = __uint_as_float(v[2 * j]) * act_deriv_const<ACT>(__low2float(hv));
                                    const float d1 = __uint_as_float(v[2 * j + 1]) * act_deriv_const<ACT>(__high2float(hv));
                                    packed[j] = pack_bf16(d0, d1);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(ostage + row * 128 + (((half * 4 + j) ^ (row & 7)) << 4)) =
                                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                        }
                    }
                    fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
                    epi_bar_sync(eg);
                    if (leader) {
                        tma_store_2d(&tmOut, smem_u32(ostage), nt * BN + jb * BCOLS, m_cta);   // clips rows >= M
                        tma_store_commit();
                        if (AUX) prefetch_aux();           // aux buffer `buf` is free again
                    }
                }
                release_accumulator();
                if (fuse_dot && m < p.M) {
#pragma unroll
                    for (int c = 0; c < DOTC; ++c) atomicAdd(p.o_accum + (size_t)m * DOTC + c, dot2[c].x + dot2[c].y);
                }
            }
            if (leader) tma_store_wait_read<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) { __syncwarp(); cluster_sync_all(); }   // the peer may still be reading our operands / barriers
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side: launch plumbing ---------------------------------------------------------------------------
template <int MODE, int ACT, int DOTC, int CG, bool OUT32, bool RES = false>
int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& x, const TcParams& p,
           int grid, cudaStream_t st) {
    static bool configured = false;
    constexpr int LM = RES ? LAYOUT_FWD_RES : MODE;
    auto kern = tc_gemm_kernel<MODE, ACT, DOTC, CG, OUT32, RES>;
    if (!configured) {
        SVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(CG, LM)));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(num_threads(CG, LM));
    cfg.dynamicSmemBytes = smem_bytes(CG, LM);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    SVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b, o, x, p));
    return SVAE_OK;
}

// dispatch on the run-time options; only the combinations the library uses are instantiated
template <int MODE, int ACT>
int launch_variant(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& x,
                   const TcParams& p, int cg, int grid, cudaStream_t st) {
#define SVAE_TC_LAUNCH(M_, D_, O_)                                                             \
    (cg == 2 ? launch<M_, ACT, D_, 2, O_>(a, b, o, x, p, grid, st)                              \
             : launch<M_, ACT, D_, 1, O_>(a, b, o, x, p, grid, st))
    if (MODE == 0) {
        if (p.res)
            return cg == 2 ? launch<0, ACT, 0, 2, false, true>(a, b, o, x, p, grid, st)
                           : launch<0, ACT, 0, 1, false, true>(a, b, o, x, p, grid, st);
        if (p.out_f32) return SVAE_TC_LAUNCH(0, 0, true);
        if (p.o_accum == nullptr) return SVAE_TC_LAUNCH(0, 0, false);
        if (p.dot_c == 1) return SVAE_TC_LAUNCH(0, 1, false);
        if (p.dot_c == 2) return SVAE_TC_LAUNCH(0, 2, false);
        return SVAE_TC_LAUNCH(0, 3, false);
    }
    if (p.out_f32) return SVAE_TC_LAUNCH(1, 0, true);
    return SVAE_TC_LAUNCH(1, 0, false);
#undef SVAE_TC_LAUNCH
}

template <int MODE>
int launch_act(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& x,
               const TcParams& p, int cg, int grid, cudaStream_t st) {
    switch (p.act) {
        case SVAE_ACT_TANH: return launch_variant<MODE, SVAE_ACT_TANH>(a, b, o, x, p, cg, grid, st);
        case SVAE_ACT_LEAKYRELU: return launch_variant<MODE, SVAE_ACT_LEAKYRELU>(a, b, o, x, p, cg, grid, st);
        case SVAE_ACT_RELU: return launch_variant<MODE, SVAE_ACT_RELU>(a, b, o, x, p, cg, grid, st);
        case SVAE_ACT_SIGMOID: return launch_variant<MODE, SVAE_ACT_SIGMOID>(a, b, o, x, p, cg, grid, st);
        default: set_error("tc_gemm: unknown activation %d", p.act); return SVAE_EINVAL;
    }
}

}  // namespace

// how many ways the K loop of an M x N forward can be split so that every CTA (pair) gets a tile
int tc_split_k_factor(int M, int N) {
    const int cg = cta_group_size();
    const int tiles = ceil_div(M, BM * cg) * ceil_div(N, BN);
    return tiles > 0 ? (sm_count() / cg) / tiles : 1;
}

int tc_gemm(int mode, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias,
            int bias_n, const void* aux, int ldaux, int act, void* out, int ldo, cudaStream_t st, const TcExtra& ex) {
    SVAE_REQUIRE(mode >= 0 && mode <= 2, SVAE_EINVAL, "tc_gemm: unknown mode %d", mode);
    if (M <= 0 || N <= 0 || K <= 0) return SVAE_OK;
    TcParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.bias_n = bias_n; p.aux = reinterpret_cast<const __nv_bfloat16*>(aux); p.ldaux = ldaux;
    p.act = act; p.out = out; p.ldo = ldo;
    p.out_w = ex.out_w; p.out_w_ld = ex.out_w_ld; p.dot_c = ex.dot_c; p.o_accum = ex.o_accum;
    p.out_f32 = (ex.out_f32 && mode != 2) ? 1 : 0;
    p.res = (mode == 0 && ex.resid != nullptr) ? 1 : 0;
    p.raw = (mode == 0 && p.out_f32 && ex.raw_split_k) ? 1 : 0;
    const bool f32 = p.out_f32 != 0;
    const uint32_t ebox = f32 ? 32 : 64;          // epilogue block: 128 bytes of columns
    const int cg = cta_group_size();
    p.m_tiles = ceil_div(M, BM * cg);
    p.n_tiles = ceil_div(N, BN);
    p.k_blocks = ceil_div(K, BK);
    p.k_splits = 1;
    p.k_blocks_per_split = p.k_blocks;
    CUtensorMap ma, mb, mo, mx;
    memset(&mo, 0, sizeof(mo));
    memset(&mx, 0, sizeof(mx));
    const int sms = sm_count();
    if (mode == 0) {
        SVAE_REQUIRE(N % 64 == 0 && K % 64 == 0, SVAE_EINVAL, "tc_gemm fwd: N, K must be multiples of 64");
        SVAE_REQUIRE(!(f32 && ex.o_accum != nullptr), SVAE_EINVAL, "tc_gemm fwd: fp32 output cannot fuse the output dot");
        SVAE_REQUIRE(ex.o_accum == nullptr || (ex.dot_c >= 1 && ex.dot_c <= MAX_DOT_C && ex.out_w != nullptr), SVAE_EINVAL,
                     "tc_gemm fwd: fused output dot supports 1..%d channels", MAX_DOT_C);
        SVAE_TRY(make_map(&ma, A, M, K, lda, 64, 128));
        SVAE_TRY(make_map(&mb, W, N, K, ldw, 64, 256 / cg));
        SVAE_TRY(make_map(&mo, out, M, N, ldo, ebox, 128, f32));
        if (p.res) {
            SVAE_REQUIRE(!f32 && ex.o_accum == nullptr, SVAE_EINVAL,
                         "tc_gemm fwd: the residual stream exists for the plain bf16 forward (no fp32 output, no fused dot)");
            SVAE_TRY(make_map(&mx, ex.resid, M, N, ex.ld_resid, 64, 128));
        }
    } else if (mode == 1) {
        SVAE_REQUIRE(N % 64 == 0 && K % 64 == 0, SVAE_EINVAL, "tc_gemm dx: N, K must be multiples of 64");
        SVAE_REQUIRE(aux != nullptr, SVAE_EINVAL, "tc_gemm dx: aux is required");
        SVAE_TRY(make_map(&ma, A, M, K, lda, 64, 128));
        SVAE_TRY(make_map(&mb, W, K, N, ldw, 64, 64));
        SVAE_TRY(make_map(&mo, out, M, N, ldo, ebox, 128, f32));
        SVAE_TRY(make_map(&mx, aux, M, N, ldaux, ebox, 128, f32));
    } else {
        // A: (K rows) x (lda cols) with M <= lda logical columns; Bm: (K rows) x (ldw cols), N <= ldw
        SVAE_TRY(make_map(&ma, A, K, round_up(M, 64) <= lda ? round_up(M, 64) : lda, lda, 64, 64));
        SVAE_TRY(make_map(&mb, W, K, round_up(N, 64) <= ldw ? round_up(N, 64) : ldw, ldw, 64, 64));
        const int mn = p.m_tiles * p.n_tiles;
        int splits = (sms / cg) / mn;
        if (splits < 1) splits = 1;
        if (splits > p.k_blocks) splits = p.k_blocks;
        p.k_blocks_per_split = ceil_div(p.k_blocks, splits);
        p.k_splits = ceil_div(p.k_blocks, p.k_blocks_per_split);
        p.vec_red = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    }
    if (p.raw) {
        // few output tiles walking a long K (the encoder's small-batch GEMMs): split K over the idle CTA pairs
        const int mn = p.m_tiles * p.n_tiles;
        int splits = (sms / cg) / mn;
        if (splits < 1) splits = 1;
        if (splits > p.k_blocks) splits = p.k_blocks;
        p.k_blocks_per_split = ceil_div(p.k_blocks, splits);
        p.k_splits = ceil_div(p.k_blocks, p.k_blocks_per_split);
        p.vec_red = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    }
    const int tiles = p.m_tiles * p.n_tiles * p.k_splits;
    const int groups = sms / cg;                                   // CTAs (cg=1) or CTA pairs (cg=2) resident at once
    const int grid = (tiles < groups ? tiles : groups) * cg;
    if (mode == 0) return launch_act<0>(ma, mb, mo, mx, p, cg, grid, st);
    if (mode == 1) return launch_act<1>(ma, mb, mo, mx, p, cg, grid, st);
    if (cg == 2) return launch<2, SVAE_ACT_TANH, 0, 2, false>(ma, mb, mo, mx, p, grid, st);
    return launch<2, SVAE_ACT_TANH, 0, 1, false>(ma, mb, mo, mx, p, grid, st);
}

}  // namespace svae
