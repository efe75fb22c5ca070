// Dependent-launch latency inside a CUDA graph: a chain of N small kernels, plain stream order against
// programmatic dependent launch (each kernel starts with griddepcontrol.wait).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl pdl.cu && ./pdl
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <bool PDL>
__global__ void link_k(const float* __restrict__ in, float* __restrict__ out, int n) {
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * 1.0001f + 1.f;
}

template <bool PDL>
int run(int chain, int blocks, int reps) {
    const int n = blocks * 256;
    float *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < chain; ++k) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, link_k<PDL>, (const float*)((k & 1) ? b : a), (k & 1) ? a : b, n));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), (chain & 1) ? b : a, n * 4, cudaMemcpyDeviceToHost));
    printf("%s chain %d x %4d blocks: %.2f us per kernel (check %.4f)\n", PDL ? "pdl  " : "plain", chain, blocks,
           ms * 1e3 / reps / chain, h[0]);
    cudaFree(a); cudaFree(b);
    return 0;
}

int main() {
    for (int blocks : {1, 128, 1024, 4096}) {
        if (run<false>(40, blocks, 200)) return 1;
        if (run<true>(40, blocks, 200)) return 1;
    }
    return 0;
}
