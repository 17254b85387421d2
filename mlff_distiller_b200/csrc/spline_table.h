// Host-side construction of the per-model radial filter splines (FP64, once per mlffd_model_create).
//
// The PaiNN filter of layer l,  f_l(d) = W2 SiLU(W1 phi~(d) + b1) + b2  in R^{3H}  with
// phi~_k(d) = exp(-gamma_k (d - mu_k)^2) * 0.5 (cos(pi d / rc) + 1) [d < rc]
// (reference src/mlff_distiller/models/student_model.py:249-255, 285-292, 318-322, 350), depends on
// the scalar distance only: no atom feature and no structure enters it.  It is therefore a property
// of the MODEL, not of a step.  Each of its 3H component functions is interpolated on [0, rc] by a
// uniform quintic B-spline with kSplineIntervals intervals; the message kernels
// (message_spline.cuh) evaluate value and d-derivative from the same six coefficients, so the
// forces stay the exact gradient of the (interpolated) energy.
//
// Stated bound (tests/test_spline_host.py, tests/test_gpu_parity.py:test_filter_spline_matches_oracle):
// with 192 intervals the interpolation error of the trained Original / Tiny / Ultra-tiny filters is
// <= 2e-8 in value and <= 2e-6 per Angstrom in derivative in exact arithmetic (max |f| ~ 2.5, max |f'| ~ 5),
// below the FP32 rounding of the stored coefficients, which dominates the total: <= 1e-7 / 6e-6 measured
// for any interval count between 160 and 256 (fewer intervals = less amplification of the coefficient
// rounding in the derivative, 1 / h; more = less truncation).
//
// Collocation: n + 5 coefficients are fixed by f at the n + 1 knots plus the midpoints of the first
// two and last two intervals (Schoenberg-Whitney holds: site s lies inside the support of basis s).
// The banded system (bandwidth <= 5 either side) is solved by Gaussian elimination with partial
// pivoting restricted to the band.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace mlffd {

#ifndef MLFFD_SPLINE_INTERVALS
#define MLFFD_SPLINE_INTERVALS 192
#endif
constexpr int kSplineIntervals = MLFFD_SPLINE_INTERVALS;
constexpr int kSplineDegree = 5;
constexpr int kSplineRows = kSplineIntervals + kSplineDegree;   // coefficients per component function
constexpr int kSliceChannels = 32;                              // channels of one shared-memory table slice

// Uniform B-spline basis of degree 5 at local parameter u in [0,1]: b[j] multiplies coefficient
// seg + j; db[j] = d b[j] / du.  Cox-de Boor recursion on integer knots:
//   b^p_j = ((u + p - j) b^{p-1}_{j-1} + (j + 1 - u) b^{p-1}_j) / p,   d b^p_j / du = b^{p-1}_{j-1} - b^{p-1}_j.
inline void quintic_basis_host(double u, double* b, double* db) {
    double prev[6] = {1, 0, 0, 0, 0, 0}, cur[6];
    for (int p = 1; p <= 5; ++p) {
        if (p == 5 && db) {
            for (int j = 0; j <= 5; ++j) db[j] = (j >= 1 ? prev[j - 1] : 0.0) - (j <= 4 ? prev[j] : 0.0);
        }
        for (int j = 0; j <= p; ++j) {
            const double left = (j >= 1) ? prev[j - 1] : 0.0;
            const double right = (j <= p - 1) ? prev[j] : 0.0;
            cur[j] = ((u + p - j) * left + (j + 1 - u) * right) / p;
        }
        for (int j = 0; j <= p; ++j) prev[j] = cur[j];
    }
    for (int j = 0; j <= 5; ++j) b[j] = prev[j];
}

// Filter of one layer at distance d (FP64): out[3H].  W1 [H][K], b1 [H], W2 [3H][H], b2 [3H] are the
// reference's rbf_to_scalar.{0,2}.{weight,bias}; gamma as the reference computes it, in FP32
// (student_model.py:252).
inline void filter_value_host(double d, int H, int K, double rc, const float* centers, const float* gammas,
                              const float* W1, const float* b1, const float* W2, const float* b2,
                              std::vector<double>& hidden, double* out) {
    const double kPi = 3.14159265358979323846;
    const double fc = (d < rc) ? 0.5 * (std::cos(kPi * d / rc) + 1.0) : 0.0;
    double phi[64];
    for (int k = 0; k < K; ++k) {
        const double diff = d - (double)centers[k];
        phi[k] = std::exp(-(double)gammas[k] * diff * diff) * fc;
    }
    hidden.resize(H);
    for (int h = 0; h < H; ++h) {
        double y = b1[h];
        for (int k = 0; k < K; ++k) y += (double)W1[(size_t)h * K + k] * phi[k];
        hidden[h] = y / (1.0 + std::exp(-y));
    }
    for (int c = 0; c < 3 * H; ++c) {
        double y = b2[c];
        const float* w = W2 + (size_t)c * H;
        for (int h = 0; h < H; ++h) y += (double)w[h] * hidden[h];
        out[c] = y;
    }
}

// Coefficients of all 3H component functions of one layer.  Returns them in the device layout
//   [slice = H / 32][row = kSplineRows][comp = a | b | c][32 channels]
// so that one table slice (what a CTA keeps in shared memory) is one contiguous block.
inline std::vector<float> build_filter_spline(int H, int K, float rc, const float* centers, const float* gammas,
                                              const float* W1, const float* b1, const float* W2,
                                              const float* b2) {
    const int n = kSplineIntervals, R = kSplineRows, C = 3 * H;
    const double h = (double)rc / n;
    // collocation sites, ascending
    std::vector<double> site;
    site.reserve(R);
    for (int i = 0; i <= n; ++i) {
        site.push_back(i * h);
        if (i == 0 || i == 1 || i == n - 2 || i == n - 1) site.push_back((i + 0.5) * h);
    }
    // dense storage, banded access
    std::vector<double> A((size_t)R * R, 0.0), rhs((size_t)R * C, 0.0), hidden;
    for (int s = 0; s < R; ++s) {
        const double x = site[s] / h;
        int seg = std::min((int)x, n - 1);
        double b[6];
        quintic_basis_host(x - seg, b, nullptr);
        for (int j = 0; j <= 5; ++j) A[(size_t)s * R + seg + j] = b[j];
        filter_value_host(std::min(site[s], (double)rc), H, K, rc, centers, gammas, W1, b1, W2, b2, hidden,
                          rhs.data() + (size_t)s * C);
    }
    const int bw = 6;   // nonzeros of row s lie within [s - bw, s + bw]; pivoting widens the upper band to 2 bw
    for (int k = 0; k < R; ++k) {
        int piv = k;
        const int r_hi = std::min(R - 1, k + bw);
        for (int r = k + 1; r <= r_hi; ++r)
            if (std::fabs(A[(size_t)r * R + k]) > std::fabs(A[(size_t)piv * R + k])) piv = r;
        const int c_hi = std::min(R - 1, k + 2 * bw);
        if (piv != k) {
            for (int c = k; c <= c_hi; ++c) std::swap(A[(size_t)k * R + c], A[(size_t)piv * R + c]);
            for (int c = 0; c < C; ++c) std::swap(rhs[(size_t)k * C + c], rhs[(size_t)piv * C + c]);
        }
        const double inv = 1.0 / A[(size_t)k * R + k];
        for (int r = k + 1; r <= r_hi; ++r) {
            const double f = A[(size_t)r * R + k] * inv;
            if (f == 0.0) continue;
            for (int c = k; c <= c_hi; ++c) A[(size_t)r * R + c] -= f * A[(size_t)k * R + c];
            double* rr = rhs.data() + (size_t)r * C;
            const double* rk = rhs.data() + (size_t)k * C;
            for (int c = 0; c < C; ++c) rr[c] -= f * rk[c];
        }
    }
    for (int k = R - 1; k >= 0; --k) {
        double* rk = rhs.data() + (size_t)k * C;
        const int c_hi = std::min(R - 1, k + 2 * bw);
        for (int c2 = k + 1; c2 <= c_hi; ++c2) {
            const double a = A[(size_t)k * R + c2];
            if (a == 0.0) continue;
            const double* rc2 = rhs.data() + (size_t)c2 * C;
            for (int c = 0; c < C; ++c) rk[c] -= a * rc2[c];
        }
        const double inv = 1.0 / A[(size_t)k * R + k];
        for (int c = 0; c < C; ++c) rk[c] *= inv;
    }
    // device layout
    const int slices = H / kSliceChannels;
    std::vector<float> out((size_t)slices * R * 3 * kSliceChannels);
    for (int sl = 0; sl < slices; ++sl)
        for (int r = 0; r < R; ++r)
            for (int comp = 0; comp < 3; ++comp)
                for (int ch = 0; ch < kSliceChannels; ++ch)
                    out[(((size_t)sl * R + r) * 3 + comp) * kSliceChannels + ch] =
                        (float)rhs[(size_t)r * C + comp * H + sl * kSliceChannels + ch];
    return out;
}

}  // namespace mlffd
