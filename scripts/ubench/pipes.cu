// Micro-benchmarks of SM pipe throughputs that bound the fused epilogues (B200, sm_100a): MUFU.TANH vs EX2 / RCP,
// 3-register FFMA vs packed fma.rn.f32x2, broadcast LDS.32/64/128.  One CTA of 512 threads per SM, clock64 timed.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    __shared__ float4 tab[256];
    if (threadIdx.x < 256) tab[threadIdx.x] = make_float4(seed * threadIdx.x, seed, 1.f, 2.f);
    __syncthreads();
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
    float2 p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
            if (OP == 4 && i < 4) {
                unsigned long long x, y, z;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(p[i].x), "f"(p[i].y));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(p[(i + 1) & 3].x), "f"(p[(i + 1) & 3].y));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(seed));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(p[i].x), "=f"(p[i].y) : "l"(x));
            }
            if (OP == 5) { float v = tab[(it + i) & 255].x; a[i] += v; }
            if (OP == 6) { float2 v = *reinterpret_cast<float2*>(&tab[(it + i) & 255]); a[i] += v.x + v.y; }
            if (OP == 7) { float4 v = tab[(it + i) & 255]; a[i] += v.x + v.y + v.z + v.w; }
            if (OP == 8) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&a[i])));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const char* names[] = {"tanh.approx.f32", "ex2.approx", "rcp.approx", "fma.rn.f32 (3 reg)", "fma.rn.f32x2", "LDS.32 bcast(+1 FADD)",
                           "LDS.64 bcast(+2 FADD)", "LDS.128 bcast(+4 FADD)", "tanh.approx.bf16x2"};
    for (int op = 0; op < 9; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
                case 0: k<0><<<148, 512>>>(out, cyc, 0.001f); break; case 1: k<1><<<148, 512>>>(out, cyc, 0.001f); break;
                case 2: k<2><<<148, 512>>>(out, cyc, 0.001f); break; case 3: k<3><<<148, 512>>>(out, cyc, 0.001f); break;
                case 4: k<4><<<148, 512>>>(out, cyc, 0.001f); break; case 5: k<5><<<148, 512>>>(out, cyc, 0.001f); break;
                case 6: k<6><<<148, 512>>>(out, cyc, 0.001f); break; case 7: k<7><<<148, 512>>>(out, cyc, 0.001f); break;
                case 8: k<8><<<148, 512>>>(out, cyc, 0.001f); break;
            }
            cudaDeviceSynchronize();
        }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        const double ops = (double)ITERS * (op == 4 ? 4 : 8) * 512;        // thread-level instructions per SM
        printf("%-26s %8.0f cycles  %6.2f thread-instr/clk/SM  (%5.2f clk per warp-instr per SMSP)\n", names[op], c, ops / c, c / (ops / 32 / 4));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
