// SIMT-on-host shim for UNIT-TESTING the library's CUDA C++ kernel sources without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  tests/simt_emu/build.py rewrites the launch syntax of spatial-vae_b200/csrc/*.cu
// (everything except the tcgen05 file) and compiles the SAME kernel bodies with g++ against this header into
// tests/simt_emu/_build/libsvae_emu.so, which tests/ load through the same C ABI to check indexing, reductions and
// the host orchestration on tiny inputs.  Nothing in the package, bench.py or __graft_entry__ knows about it; the
// product has no CPU path.
//
// Execution model: one OS thread.  A launch runs its blocks one after another; the threads of a block are
// ucontext fibers scheduled round-robin, yielding only at __syncthreads() (block barrier) and at warp shuffles
// (per-warp barrier), so static __shared__ arrays are simply function-local statics and atomics are plain adds.
// Finished threads count as arrived at every barrier (the CUDA rule for exited threads).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <functional>
#include <type_traits>
#include <vector>

// ---- qualifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define SVAE_SIMT_EMU 1

// ---- vector types ---------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// packed f32x2 arithmetic of sm_100 (crt/sm_100_rt.h)
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }

// ---- bf16 (round to nearest even, like the _rn intrinsics) --------------------------------------------------
struct __nv_bfloat16 { uint16_t bits; };
struct alignas(4) __nv_bfloat162 { __nv_bfloat16 x, y; };   // x = low half
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = (uint16_t)((u >> 16) | 0x40); return r; }   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    r.bits = (uint16_t)(u >> 16);
    return r;
}
inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t u = (uint32_t)h.bits << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
    return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)};
}
inline float __low2float(__nv_bfloat162 v) { return __bfloat162float(v.x); }
inline float __high2float(__nv_bfloat162 v) { return __bfloat162float(v.y); }

// ---- runtime types ---------------------------------------------------------------------------------
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// SM count of the emulated device: 148 like a B200, or SVAE_EMU_SMS (a small value makes persistent kernels walk
// several tiles per CTA, which is where accumulator double-buffering and barrier phases get exercised)
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) {
    const char* e = getenv("SVAE_EMU_SMS");
    *v = (e != nullptr && atoi(e) > 0) ? atoi(e) : 148;
    return cudaSuccess;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// streams and events: launches execute synchronously in program order, which is one of the orders any stream / event
// dependency graph allows, so these only hand out distinct handles
typedef struct CUevent_st* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { static long n = 0; *s = (cudaStream_t)(++n * 64); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { static long n = 0; *e = (cudaEvent_t)(++n * 64); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

// ---- built-in variables and the fiber scheduler ------------------------------------------------------
namespace svae_emu {

// wait reasons of a fiber
enum { RUNNABLE, WAIT_BLOCK, WAIT_WARP, WAIT_NAMED, WAIT_CLUSTER, WAIT_COND, DONE };

struct State {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<int> status;
    std::vector<int> named_id;        // WAIT_NAMED: barrier id
    std::vector<int> named_count;     // WAIT_NAMED: participants
    std::vector<std::function<bool()>> cond;   // WAIT_COND: runnable again once this returns true
    std::vector<const char*> why;     // WAIT_COND: what the fiber waits for (deadlock report)
    std::vector<char*> stacks;
    std::vector<float> shfl;          // one exchange slot per thread
    std::function<void()>* body = nullptr;
    int n = 0;                        // fibers of the running cluster = cluster_size * threads_per_cta
    int per_cta = 0, cluster = 1;     // threads per CTA, CTAs per cluster
    int cur = 0;                      // running fiber
    unsigned first_block = 0;         // blockIdx.x of CTA 0 of the running cluster
    std::vector<void*> dyn;           // dynamic shared memory arena of each CTA of the cluster
    size_t dyn_cap = 0;               // allocated bytes per arena
    size_t dyn_req = 0;               // bytes the running launch asked for
    int cta() const { return cur / per_cta; }
    int tid() const { return cur % per_cta; }
};
State& state();
void yield(int why);
void wait_named(int id, int count);
void wait_until(std::function<bool()> cond, const char* why);
void run_cluster(std::function<void()>& body, int threads_per_cta, int cluster_size, unsigned first_block, size_t smem);
void* dyn_smem();                     // arena of the running fiber's CTA
void* dyn_smem_of(int cta_rank);      // arena of another CTA of the cluster
unsigned long long globaltimer();

}  // namespace svae_emu

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

inline void __syncthreads() { svae_emu::yield(svae_emu::WAIT_BLOCK); }
inline void __syncwarp(unsigned = 0xffffffffu) { svae_emu::yield(svae_emu::WAIT_WARP); }
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    svae_emu::State& s = svae_emu::state();
    const int me = s.cur;
    s.shfl[me] = v;
    svae_emu::yield(svae_emu::WAIT_WARP);
    const int partner = (me & ~31) | ((me & 31) ^ lane_mask);
    const float r = (partner < s.n) ? s.shfl[partner] : v;
    svae_emu::yield(svae_emu::WAIT_WARP);
    return r;
}
inline float __shfl_sync(unsigned, float v, int src_lane) {
    svae_emu::State& s = svae_emu::state();
    const int me = s.cur;
    s.shfl[me] = v;
    svae_emu::yield(svae_emu::WAIT_WARP);
    const int partner = (me & ~31) | (src_lane & 31);
    const float r = (partner < s.n) ? s.shfl[partner] : v;
    svae_emu::yield(svae_emu::WAIT_WARP);
    return r;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline void __trap() { fprintf(stderr, "svae_emu: __trap()\n"); abort(); }
inline void sincospi(double x, double* s, double* c) {       // exact at multiples of 1/2 like CUDA's, else libm
    const double r = x - 2.0 * floor(x / 2.0);               // [0, 2)
    if (r == 0.0) { *s = 0.0; *c = 1.0; }
    else if (r == 0.5) { *s = 1.0; *c = 0.0; }
    else if (r == 1.0) { *s = 0.0; *c = -1.0; }
    else if (r == 1.5) { *s = -1.0; *c = 0.0; }
    else { *s = sin(M_PI * r); *c = cos(M_PI * r); }
}
inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }
inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; return old; }
inline long long clock64() { return (long long)(svae_emu::globaltimer() * 2); }   // "2 GHz"

// round-to-nearest single operations (the build uses -ffp-contract=off, so plain operators are exactly these)
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }

template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a > (T)b ? (T)a : (T)b;
}

// ---- launches: kernel<<<grid, block, smem, stream>>>(args...) is rewritten by build.py into
//      svae_emu::Launcher(grid, block, smem, stream).run(kernel, args...) ---------------------------------------
namespace svae_emu {
struct Launcher {
    dim3 g, b;
    size_t smem;
    unsigned cluster;
    Launcher(dim3 g_, dim3 b_, size_t smem_ = 0, cudaStream_t = nullptr, unsigned cluster_ = 1)
        : g(g_), b(b_), smem(smem_), cluster(cluster_) {}
    template <typename... P, typename... A>
    void run(void (*k)(P...), A&&... a) {
        std::function<void()> body = [&]() { k(a...); };
        gridDim = g;
        blockDim = b;
        for (unsigned z = 0; z < g.z; ++z)
            for (unsigned y = 0; y < g.y; ++y)
                for (unsigned x = 0; x < g.x; x += cluster) {
                    blockIdx = uint3{x, y, z};
                    run_cluster(body, (int)(b.x * b.y * b.z), (int)cluster, x, smem);
                }
    }
};
}  // namespace svae_emu
