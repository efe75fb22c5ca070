// One-off ingest work moved to the device: per-particle real-space CTF kernels.
//
// The reference builds them on the host, one particle at a time in a Python loop (spatial_vae/ctf.py:33-56):
//   c[a,b]   = CTF at the spatial frequency (fftfreq(n)[a], fftfreq(m)[b]) / (apix * scale)      (ctf.py:7-24, 43-52)
//   out      = -fftshift(ifft2(c)).real                                                          (ctf.py:54)
// Here one thread block does one particle: the CTF is evaluated into shared memory, then the inverse DFT is taken
// as two separable passes with tabulated twiddles (only the real part of the second pass is needed because c is
// real), the shift is folded into the output index.  Everything is fp64 like numpy, the result is cast to fp32 at
// the end like the reference.  `defocus` is used for both astigmatism axes, as the reference does (ctf.py:45-46).
#include "kernels.cuh"

namespace svae {

namespace {
__device__ __forceinline__ double fft_freq(int i, int n) {       // numpy.fft.fftfreq(n)[i]
    return (double)(i < (n + 1) / 2 ? i : i - n) / (double)n;
}
}  // namespace

__global__ void __launch_bounds__(256) ctf_filter_k(const double* __restrict__ params, int n, int m, double scale,
                                                    float* __restrict__ out) {
    extern __shared__ double sd[];
    double* c = sd;                    // n x m   CTF in frequency space
    double* re = c + n * m;            // n x m   row pass, real part
    double* im = re + n * m;           // n x m   row pass, imaginary part
    double* cm = im + n * m;           // m       cos(2 pi k / m)
    double* sm_ = cm + m;              // m       sin(2 pi k / m)
    double* cn = sm_ + m;              // n
    double* sn = cn + n;               // n
    const double* p = params + (long)blockIdx.x * 8;   // defocus cs voltage apix bfactor ampcont dfdiff dfang
    const double PI = 3.14159265358979323846;           // numpy.pi
    const double defocus = p[0] * 10000.0, cs = p[1] * 1e7, volt = p[2] * 1000.0, apix = p[3] * scale;
    const double bfactor = p[4], w = p[5] / 100.0;
    const double lam = 12.2639 / sqrt(volt + 0.97845e-6 * volt * volt);
    for (int k = threadIdx.x; k < m; k += blockDim.x) sincospi(2.0 * k / m, &sm_[k], &cm[k]);
    for (int k = threadIdx.x; k < n; k += blockDim.x) sincospi(2.0 * k / n, &sn[k], &cn[k]);
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) {
        const double x = fft_freq(i / m, n) / apix, y = fft_freq(i % m, m) / apix;
        const double s2 = x * x + y * y;
        const double gamma = 2.0 * PI * (-0.5 * defocus * lam * s2 + 0.25 * cs * lam * lam * lam * s2 * s2);
        c[i] = (sqrt(1.0 - w * w) * sin(gamma) - w * cos(gamma)) * exp(-bfactor / 4.0 * s2);
    }
    __syncthreads();
    // rows: T[a, v] = sum_b c[a, b] exp(+2 pi i b v / m)
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) {
        const int a = i / m, v = i % m;
        double r = 0.0, q = 0.0;
        int k = 0;                                            // (b * v) mod m
        for (int b = 0; b < m; ++b) {
            r = fma(c[a * m + b], cm[k], r);
            q = fma(c[a * m + b], sm_[k], q);
            k += v;
            if (k >= m) k -= m;
        }
        re[i] = r;
        im[i] = q;
    }
    __syncthreads();
    // columns, real part only: R[u, v] = sum_a (Tre[a, v] cos(2 pi a u / n) - Tim[a, v] sin(2 pi a u / n)) / (n m)
    float* o = out + (long)blockIdx.x * n * m;
    const double inv = 1.0 / ((double)n * (double)m);
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) {
        const int u = i / m, v = i % m;
        double r = 0.0;
        int k = 0;                                            // (a * u) mod n
        for (int a = 0; a < n; ++a) {
            r += re[a * m + v] * cn[k] - im[a * m + v] * sn[k];
            k += u;
            if (k >= n) k -= n;
        }
        // fftshift: shifted[(u + n/2) mod n][(v + m/2) mod m] = R[u][v]; the kernel is the negative of that
        const int us = (u + n / 2) % n, vs = (v + m / 2) % m;
        o[us * m + vs] = (float)(-(r * inv));
    }
}

int ctf_filter(const double* params, int N, int n, int m, double scale, float* out, cudaStream_t st) {
    if (N == 0) return SVAE_OK;
    const size_t smem = ((size_t)3 * n * m + 2 * (n + m)) * sizeof(double);
    SVAE_REQUIRE(smem <= 200 * 1024, SVAE_EINVAL, "CTF kernels of %d x %d do not fit in shared memory", n, m);
    if (smem > 48 * 1024)
        SVAE_CUDA(cudaFuncSetAttribute(ctf_filter_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctf_filter_k<<<N, 256, smem, st>>>(params, n, m, scale, out);
    SVAE_LAUNCH_CHECK();
    return SVAE_OK;
}

}  // namespace svae
