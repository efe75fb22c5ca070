// Arithmetic of the decoder's first layer with the reference's coordinate options, shared by the
// CUDA kernels (step_kernels.cu) and -- compiled with g++ -- by the CPU checks in tests/ (the same
// source, so the formulas are verified against the oracle's autograd without a GPU).
//
//   models.py:99-102   expand_coords: x -> (x0, x1, x0^2, x1^2, x0*x1)            (F = 5, else F = 2)
//   models.py:104      coord_linear(x)
//   models.py:114-121  bilinear(x, z): sum_ij f_i Wb[n,i,j] z_j  ==  a per-image coordinate weight
//                      W_eff[b][n,i] = Wc[n,i] + sum_j Wb[n,i,j] z[b,j]
//   train_mnist.py:50-59,70-74  x' = R(theta) c + t  with c the untransformed grid coordinate
//
// Backward runs through the per-image FEATURE MOMENTS of delta0 = dL/d(pre-activation of layer 0):
//   T[b][0][n] = sum_p delta0[b,p,n],   T[b][1+i][n] = sum_p delta0[b,p,n] * f_i(x'_bp)
// from which   dW_eff[b][n,i] = T[b][1+i][n],  d(coord_b) = sum_b T[b][0],  and (d theta, d t) follow
// in closed form because d f / d x' is itself linear in the features (latent_coord_terms below).
#pragma once
#if defined(__CUDACC__)
#define SVAE_HD __host__ __device__ __forceinline__
#else
#define SVAE_HD inline
#endif

namespace svae {

constexpr int kMaxCoordFeatures = 5;

// img = (cos theta, sin theta, t0, t1)
SVAE_HD void transform_coord(float c0, float c1, float cs, float sn, float t0, float t1, float& x0, float& x1) {
    x0 = c0 * cs - c1 * sn + t0;
    x1 = c0 * sn + c1 * cs + t1;
}

template <int F>
SVAE_HD void coord_features(float x0, float x1, float (&f)[kMaxCoordFeatures]) {
    f[0] = x0;
    f[1] = x1;
    if (F == 5) {
        f[2] = x0 * x0;
        f[3] = x1 * x1;
        f[4] = x0 * x1;
    }
}

// d(sum_i w_i f_i)/dx0 and /dx1 at (x0, x1); w has F entries, element stride sw
template <int F>
SVAE_HD void feature_jacobian(const float* w, long sw, float x0, float x1, float& j0, float& j1) {
    j0 = w[0];
    j1 = w[sw];
    if (F == 5) {
        j0 += 2.f * x0 * w[2 * sw] + x1 * w[4 * sw];
        j1 += 2.f * x1 * w[3 * sw] + x0 * w[4 * sw];
    }
}

// One hidden unit's contribution to (dL/dtheta, dL/dt0, dL/dt1) of its image.
//   w: the unit's F coordinate weights (stride sw);  T: its F+1 feature moments (stride sT).
// With u = x' - t:  dx0'/dtheta = -u1,  dx1'/dtheta = u0, so
//   dtheta = sum_p delta0 * ( -(x1'-t1) * j0(x') + (x0'-t0) * j1(x') ),   dt_k = sum_p delta0 * j_k(x').
template <int F>
SVAE_HD void latent_coord_terms(const float* w, long sw, const float* T, long sT, float t0, float t1, float& dth,
                                float& d0, float& d1) {
    const float T1 = T[0], Tx0 = T[sT], Tx1 = T[2 * sT];
    const float w0 = w[0], w1 = w[sw];
    d0 = w0 * T1;
    d1 = w1 * T1;
    dth = w0 * (t1 * T1 - Tx1) + w1 * (Tx0 - t0 * T1);
    if (F == 5) {
        const float Tx00 = T[3 * sT], Tx11 = T[4 * sT], Tx01 = T[5 * sT];
        const float w2 = w[2 * sw], w3 = w[3 * sw], w4 = w[4 * sw];
        d0 += 2.f * w2 * Tx0 + w4 * Tx1;
        d1 += 2.f * w3 * Tx1 + w4 * Tx0;
        dth += 2.f * w2 * (t1 * Tx0 - Tx01) + w4 * (t1 * Tx1 - Tx11)      // -(x1'-t1) * (2 x0' w2 + x1' w4)
             + 2.f * w3 * (Tx01 - t0 * Tx1) + w4 * (Tx00 - t0 * Tx0);     //  (x0'-t0) * (2 x1' w3 + x0' w4)
    }
}

}  // namespace svae
