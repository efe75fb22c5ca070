// Plain-loop stand-in for tc_gemm (spatial-vae_b200/csrc/tc_gemm.cu: tcgen05 / TMEM / TMA, cannot run on a host)
// with the same contract -- bf16 operands, fp32 accumulation, the epilogues of modes 0 / 1 / 2 and the fused
// output-layer dot product -- so the host orchestration of the FAST precision path (api.cu) can be exercised in
// tests/simt_emu.  Test infrastructure only; says nothing about the real kernel.
#include "common.cuh"

namespace svae {

static float act_f(int act, float a) {
    switch (act) {
        case SVAE_ACT_TANH: return tanhf(a);
        case SVAE_ACT_LEAKYRELU: return a > 0.f ? a : 0.01f * a;
        case SVAE_ACT_RELU: return a > 0.f ? a : 0.f;
        default: return 1.f / (1.f + expf(-a));
    }
}
static float dact_f(int act, float h) {
    switch (act) {
        case SVAE_ACT_TANH: return 1.f - h * h;
        case SVAE_ACT_LEAKYRELU: return h > 0.f ? 1.f : 0.01f;
        case SVAE_ACT_RELU: return h > 0.f ? 1.f : 0.f;
        default: return h * (1.f - h);
    }
}

int tc_gemm(int mode, int M, int N, int K, const void* A_, int lda, const void* W_, int ldw, const float* bias,
            int bias_n, const void* aux, int ldaux, int act, void* out, int ldo, cudaStream_t, const TcExtra& ex) {
    SVAE_REQUIRE(mode >= 0 && mode <= 2, SVAE_EINVAL, "tc_gemm: unknown mode %d", mode);
    if (M <= 0 || N <= 0 || K <= 0) return SVAE_OK;
    count_launch();
    const __nv_bfloat16* A = (const __nv_bfloat16*)A_;
    const __nv_bfloat16* W = (const __nv_bfloat16*)W_;
    const bool f32 = ex.out_f32 && mode != 2;
    SVAE_REQUIRE(mode == 2 || (N % 64 == 0 && K % 64 == 0), SVAE_EINVAL, "tc_gemm: N, K must be multiples of 64");
    SVAE_REQUIRE(ex.red_S == nullptr, SVAE_EINVAL, "tc_gemm emulation: fused reduction not emulated");
    SVAE_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, SVAE_EALIGN, "TMA operands need a 16-byte row pitch");
    if (mode == 2) {
        float* o = (float*)out;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                float acc = 0.f;
                for (int r = 0; r < K; ++r)
                    acc += __bfloat162float(A[(long)r * lda + i]) * __bfloat162float(W[(long)r * ldw + j]);
                o[(long)i * ldo + j] += acc;
            }
        return SVAE_OK;
    }
    SVAE_REQUIRE(mode == 0 || aux != nullptr, SVAE_EINVAL, "tc_gemm dx: aux is required");
    for (int m = 0; m < M; ++m) {
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) {
                const float w = mode == 0 ? __bfloat162float(W[(long)n * ldw + k]) : __bfloat162float(W[(long)k * ldw + n]);
                acc += __bfloat162float(A[(long)m * lda + k]) * w;
            }
            float r;
            if (mode == 0) {
                r = act_f(act, acc + ((bias != nullptr && n < bias_n) ? bias[n] : 0.f));
                if (ex.o_accum != nullptr && n < ex.out_w_ld)
                    for (int c = 0; c < ex.dot_c; ++c) dot[c] += r * ex.out_w[(long)c * ex.out_w_ld + n];
            } else {
                const float h = f32 ? ((const float*)aux)[(long)m * ldaux + n]
                                    : __bfloat162float(((const __nv_bfloat16*)aux)[(long)m * ldaux + n]);
                r = acc * dact_f(act, h);
            }
            if (f32) ((float*)out)[(long)m * ldo + n] = r;
            else ((__nv_bfloat16*)out)[(long)m * ldo + n] = __float2bfloat16_rn(r);
        }
        if (mode == 0 && ex.o_accum != nullptr)
            for (int c = 0; c < ex.dot_c; ++c) ex.o_accum[(long)m * ex.dot_c + c] += dot[c];
    }
    return SVAE_OK;
}

}  // namespace svae
