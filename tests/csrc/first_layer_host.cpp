// Host build (g++) of spatial-vae_b200/csrc/first_layer.cuh for tests/test_first_layer_math.py: plain loops that
// play the role of the CUDA kernels' thread indexing around the SAME arithmetic the kernels compile.
// Test infrastructure only -- nothing in the product links this.
#include "../../spatial-vae_b200/csrc/first_layer.cuh"

using namespace svae;

template <int F>
static void forward_t(int B, int P, int H, const float* grid, const float* img, const float* w, long sb, long sn,
                      long si, const float* hz, float* pre) {
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            float x0, x1, f[kMaxCoordFeatures];
            transform_coord(grid[p * 2], grid[p * 2 + 1], img[b * 4], img[b * 4 + 1], img[b * 4 + 2], img[b * 4 + 3], x0, x1);
            coord_features<F>(x0, x1, f);
            for (int n = 0; n < H; ++n) {
                float a = hz[(long)b * H + n];
                for (int i = F - 1; i >= 0; --i) a = __builtin_fmaf(w[b * sb + n * sn + i * si], f[i], a);
                pre[((long)b * P + p) * H + n] = a;
            }
        }
}

template <int F>
static void moments_t(int B, int P, int H, const float* grid, const float* img, const float* d0, float* T) {
    for (int b = 0; b < B; ++b) {
        float* Tb = T + (long)b * (F + 1) * H;
        for (long i = 0; i < (long)(F + 1) * H; ++i) Tb[i] = 0.f;
        for (int p = 0; p < P; ++p) {
            float x0, x1, f[kMaxCoordFeatures];
            transform_coord(grid[p * 2], grid[p * 2 + 1], img[b * 4], img[b * 4 + 1], img[b * 4 + 2], img[b * 4 + 3], x0, x1);
            coord_features<F>(x0, x1, f);
            for (int n = 0; n < H; ++n) {
                const float d = d0[((long)b * P + p) * H + n];
                Tb[n] += d;
                for (int i = 0; i < F; ++i) Tb[(1 + i) * H + n] = __builtin_fmaf(f[i], d, Tb[(1 + i) * H + n]);
            }
        }
    }
}

template <int F>
static void latent_t(int B, int H, const float* w, long sb, long sn, long si, const float* T, const float* img,
                     float* out) {
    for (int b = 0; b < B; ++b) {
        float dth = 0.f, d0 = 0.f, d1 = 0.f;
        for (int n = 0; n < H; ++n) {
            float a, b0, b1;
            latent_coord_terms<F>(w + b * sb + n * sn, si, T + (long)b * (F + 1) * H + n, H, img[b * 4 + 2], img[b * 4 + 3],
                                  a, b0, b1);
            dth += a; d0 += b0; d1 += b1;
        }
        out[b * 3] = dth; out[b * 3 + 1] = d0; out[b * 3 + 2] = d1;
    }
}

template <int F>
static void row_grad_t(int B, int P, int H, const float* x, const float* w, long sb, long sn, long si, const float* d0,
                       float* gx) {
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            const long r = (long)b * P + p;
            float a0 = 0.f, a1 = 0.f;
            for (int n = 0; n < H; ++n) {
                float j0, j1;
                feature_jacobian<F>(w + b * sb + n * sn, si, x[r * 2], x[r * 2 + 1], j0, j1);
                a0 = __builtin_fmaf(d0[r * H + n], j0, a0);
                a1 = __builtin_fmaf(d0[r * H + n], j1, a1);
            }
            gx[r * 2] = a0; gx[r * 2 + 1] = a1;
        }
}

extern "C" {
// w[b*sb + n*sn + i*si]: coordinate weight i of hidden unit n for image b (sb = 0: shared by all images)
void fl_forward(int F, int B, int P, int H, const float* grid, const float* img, const float* w, long sb, long sn,
                long si, const float* hz, float* pre) {
    if (F == 5) forward_t<5>(B, P, H, grid, img, w, sb, sn, si, hz, pre); else forward_t<2>(B, P, H, grid, img, w, sb, sn, si, hz, pre);
}
void fl_moments(int F, int B, int P, int H, const float* grid, const float* img, const float* d0, float* T) {
    if (F == 5) moments_t<5>(B, P, H, grid, img, d0, T); else moments_t<2>(B, P, H, grid, img, d0, T);
}
void fl_latent(int F, int B, int H, const float* w, long sb, long sn, long si, const float* T, const float* img,
               float* out) {
    if (F == 5) latent_t<5>(B, H, w, sb, sn, si, T, img, out); else latent_t<2>(B, H, w, sb, sn, si, T, img, out);
}
void fl_row_grad(int F, int B, int P, int H, const float* x, const float* w, long sb, long sn, long si,
                 const float* d0, float* gx) {
    if (F == 5) row_grad_t<5>(B, P, H, x, w, sb, sn, si, d0, gx); else row_grad_t<2>(B, P, H, x, w, sb, sn, si, d0, gx);
}
}
