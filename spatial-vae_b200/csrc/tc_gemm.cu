// bf16 tensor-core GEMMs for the decoder hidden layers (models.py:82,126 and their backward) on
// sm_100a: TMA (cp.async.bulk.tensor) feeds a 4-stage shared-memory ring, one elected thread issues
// tcgen05.mma with the fp32 accumulator in TMEM (two 128x256 accumulator stages = all 512 columns),
// four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue while the next tile's
// MMAs run.  Persistent CTAs, one per SM.
//
//   mode 0  FWD : out[M,N]  = act(A[M,K] W[N,K]^T + bias)              A K-major,  B K-major
//   mode 1  DX  : out[M,N]  = (A[M,K] W[K,N]) .* act'(aux[M,N])        A K-major,  B MN-major
//   mode 2  DW  : outf[M,N] += A[Kr,M]^T Bm[Kr,N]  (split over Kr)     A MN-major, B MN-major
#include <cuda.h>

#include "common.cuh"

namespace svae {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int BOX_BYTES = 64 * 64 * 2;       // one 64x64 bf16 TMA box (MN-major operands)
constexpr int MAX_BIAS = 2048;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + MAX_BIAS * 4 + 256 + 1024;  // + barriers + align slack
constexpr int NUM_THREADS = 256;
constexpr unsigned long long WAIT_TIMEOUT_CYCLES = 4000000000ull;  // ~2 s: trap instead of hanging the GPU

struct TcParams {
    int M, N, K;              // modes 0/1: rows, Hp, Hp;  mode 2: out rows, out cols, reduction rows
    int m_tiles, n_tiles, k_blocks, k_splits, k_blocks_per_split;
    const float* bias; int bias_n;
    const __nv_bfloat16* aux; int ldaux;
    int act;
    void* out; int ldo;
    int vec_red;              // mode 2: 16-byte aligned rows -> red.global.add.v4.f32
};

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if ((++spins & 0xfff) == 0) {
            const unsigned long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (tcgen05): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1),
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn, int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + MAX_BIAS * 4);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the tmem base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    if (MODE == 0) {
        for (int i = threadIdx.x; i < MAX_BIAS; i += NUM_THREADS)
            s_bias[i] = (p.bias != nullptr && i < p.bias_n) ? p.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.m_tiles * p.n_tiles * p.k_splits;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int ks = tile / (p.m_tiles * p.n_tiles);
            const int mn = tile % (p.m_tiles * p.n_tiles);
            const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
            const int kb0 = ks * p.k_blocks_per_split;
            const int kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                mbar_expect_tx(fb, STAGE_BYTES);
                const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                const uint32_t sb = smem_u32(smem_b + stage * B_STAGE_BYTES);
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * BOX_BYTES, &tmA, fb, mt * BM + i * 64, kb * BK);
                } else {
                    tma_load_2d(sa, &tmA, fb, kb * BK, mt * BM);
                }
                if (MODE == 0) {
                    tma_load_2d(sb, &tmB, fb, kb * BK, nt * BN);
                } else {
#pragma unroll
                    for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * BOX_BYTES, &tmB, fb, nt * BN + i * 64, kb * BK);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(MODE == 2 ? 1 : 0, MODE == 0 ? 0 : 1, BM, BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
                    umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * stage);                 // frees the smem slot once the MMAs retire
                if (kb == kb1 - 1) umma_commit(tfull0 + 8 * acc);  // accumulator complete
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                // TMEM lane quadrant of this warp
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mn = tile % (p.m_tiles * p.n_tiles);
            const int mt = mn / p.n_tiles, nt = mn % p.n_tiles;
            const int m = mt * BM + q * 32 + lane;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                const int n = nt * BN + c;
                if (n >= p.N) break;       // N is a multiple of 64 in modes 0/1; mode 2 masks per element
                uint32_t v[32];
                uint4 auxv[4];
                if (MODE == 1 && m < p.M) {
                    const uint4* ap = reinterpret_cast<const uint4*>(p.aux + (size_t)m * p.ldaux + n);
#pragma unroll
                    for (int j = 0; j < 4; ++j) auxv[j] = __ldg(ap + j);
                }
                tmem_ld32(t_row + c, v);
                tmem_ld_wait();
                if (MODE == 0) {
                    if (m < p.M) {
                        uint32_t packed[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float a0 = __uint_as_float(v[2 * j]) + s_bias[n + 2 * j];
                            const float a1 = __uint_as_float(v[2 * j + 1]) + s_bias[n + 2 * j + 1];
                            packed[j] = pack_bf16(act_apply<true>(p.act, a0), act_apply<true>(p.act, a1));
                        }
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.ldo + n);
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    }
                } else if (MODE == 1) {
                    if (m < p.M) {
                        uint32_t packed[16];
                        const uint32_t* aw = reinterpret_cast<const uint32_t*>(auxv);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
                            const float d0 = __uint_as_float(v[2 * j]) * act_deriv_from_out(p.act, __low2float(hv));
                            const float d1 = __uint_as_float(v[2 * j + 1]) * act_deriv_from_out(p.act, __high2float(hv));
                            packed[j] = pack_bf16(d0, d1);
                        }
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.ldo + n);
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    }
                } else {
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
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D bf16 row-major tensor (rows x cols, leading dimension ld elements), box = box_cols x box_rows
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
             uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    SVAE_REQUIRE(fn != nullptr, SVAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    SVAE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0, SVAE_EALIGN,
                 "bf16 matrices need 16-byte aligned base and leading dimension %% 8 == 0");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVAE_REQUIRE(r == CUDA_SUCCESS, SVAE_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return SVAE_OK;
}

template <int MODE>
int launch(const CUtensorMap& a, const CUtensorMap& b, const TcParams& p, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        SVAE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    tc_gemm_kernel<MODE><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, b, p);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

}  // namespace

int tc_gemm(int mode, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias,
            int bias_n, const void* aux, int ldaux, int act, void* out, int ldo, cudaStream_t st) {
    SVAE_REQUIRE(mode >= 0 && mode <= 2, SVAE_EINVAL, "tc_gemm: unknown mode %d", mode);
    if (M <= 0 || N <= 0 || K <= 0) return SVAE_OK;
    TcParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.bias_n = bias_n; p.aux = reinterpret_cast<const __nv_bfloat16*>(aux); p.ldaux = ldaux;
    p.act = act; p.out = out; p.ldo = ldo;
    p.m_tiles = ceil_div(M, BM);
    p.n_tiles = ceil_div(N, BN);
    p.k_blocks = ceil_div(K, BK);
    p.k_splits = 1;
    p.k_blocks_per_split = p.k_blocks;
    CUtensorMap ma, mb;
    const int sms = sm_count();
    if (mode == 0) {
        SVAE_REQUIRE(N % 64 == 0 && K % 64 == 0 && N <= MAX_BIAS, SVAE_EINVAL, "tc_gemm fwd: N, K must be multiples of 64");
        SVAE_REQUIRE(ldo % 8 == 0, SVAE_EALIGN, "tc_gemm: output leading dimension %% 8 != 0");
        SVAE_TRY(make_map(&ma, A, M, K, lda, 64, 128));
        SVAE_TRY(make_map(&mb, W, N, K, ldw, 64, 256));
    } else if (mode == 1) {
        SVAE_REQUIRE(N % 64 == 0 && K % 64 == 0, SVAE_EINVAL, "tc_gemm dx: N, K must be multiples of 64");
        SVAE_REQUIRE(ldo % 8 == 0 && ldaux % 8 == 0 && aux != nullptr, SVAE_EALIGN, "tc_gemm dx: aux/out alignment");
        SVAE_TRY(make_map(&ma, A, M, K, lda, 64, 128));
        SVAE_TRY(make_map(&mb, W, K, N, ldw, 64, 64));
    } else {
        // A: (K rows) x (lda cols) with M <= lda logical columns; Bm: (K rows) x (ldw cols), N <= ldw
        SVAE_TRY(make_map(&ma, A, K, round_up(M, 64) <= lda ? round_up(M, 64) : lda, lda, 64, 64));
        SVAE_TRY(make_map(&mb, W, K, round_up(N, 64) <= ldw ? round_up(N, 64) : ldw, ldw, 64, 64));
        const int mn = p.m_tiles * p.n_tiles;
        int splits = sms / mn;
        if (splits < 1) splits = 1;
        if (splits > p.k_blocks) splits = p.k_blocks;
        p.k_blocks_per_split = ceil_div(p.k_blocks, splits);
        p.k_splits = ceil_div(p.k_blocks, p.k_blocks_per_split);
        p.vec_red = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    }
    const int tiles = p.m_tiles * p.n_tiles * p.k_splits;
    const int grid = tiles < sms ? tiles : sms;
    if (mode == 0) return launch<0>(ma, mb, p, grid, st);
    if (mode == 1) return launch<1>(ma, mb, p, grid, st);
    return launch<2>(ma, mb, p, grid, st);
}

}  // namespace svae
