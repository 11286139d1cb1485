// EPnP (Lepetit, Moreno-Noguer, Fua 2009) as OpenCV's `cv::epnp` runs it, in double precision,
// written as __host__ __device__ building blocks so the math can be unit-tested on the host
// (tests/ compiles this header with g++) while the product path only ever calls it from CUDA
// kernels (csrc/pnp_ransac.cu).
//
// Reference call site: pix2pose_model/recognition.py:216-217
//   cv2.solvePnPRansac(obj, img, camK, None, flags=cv2.SOLVEPNP_EPNP, reprojectionError=5, iterationsCount=100)
// OpenCV itself is a third-party dependency that is not vendored in the reference
// (requirements.txt:3 pins opencv-python 3.4.2.17; this image has 4.13.0).  This file restates the
// published algorithm of modules/calib3d/src/epnp.cpp: control points from PCA, barycentric
// coordinates, M^T M null space, three beta approximations (N = 4 / 2 / 3 linearisations), five
// Gauss-Newton steps each, Horn-style R|t from the SVD of the 3x3 correlation, best reprojection
// error wins -- and, because the 5-point case depends on every rounding of it, OpenCV's linear algebra
// (Jacobi SVD, SVD back-substitution, mulTransposed) operation by operation; both are pinned against the
// image's cv2 by tests/test_epnp_host.py (bitwise).  Compile WITHOUT mul+add contraction (-fmad=false /
// -ffp-contract=off): an FMA where OpenCV has a separate multiply and add changes the result.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define P2P_HD __host__ __device__
#else
#define P2P_HD
#endif

namespace p2p {
namespace epnp {

struct Cam { double fu, fv, uc, vc; };

P2P_HD inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
P2P_HD inline double dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

// Symmetric eigen-decomposition by Householder tridiagonalisation + implicit-shift QL (the classic
// tred2 / tqli pair): ~10x fewer operations than cyclic Jacobi for N = 12.  Used by the LARGE-n refit only (one per
// problem, serial tail of epnp_refit_kernel): for n >= 6 noisy points the four smallest eigenvalues of M^T M are well
// separated and any accurate solver agrees with OpenCV's Jacobi to ~1e-13; the 5-point hypotheses and the refits on small
// consensus sets need OpenCV's exact Jacobi (cv_jacobi_svd12_ut below).  `A` (row-major, symmetric) is destroyed, w
// descending, row i of `Vt` = unit eigenvector of w[i].  FIRST > 0 keeps only the eigenvectors FIRST .. N-1 (the
// smallest eigenvalues): row i of the full Vt lands in row i - FIRST of `Vt` ((N - FIRST) x N); EPnP reads the last four.
template <int N, int FIRST = 0>
P2P_HD inline void tridiag_eig_sym(double* A, double* Vt, double* w) {
    double d[N], e[N];
    // ---- tred2: A -> tridiagonal (d, e), A overwritten by the orthogonal transformation Q
    for (int i = N - 1; i > 0; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += fabs(A[i * N + k]);
            if (scale == 0.0) {
                e[i] = A[i * N + l];
            } else {
                for (int k = 0; k <= l; ++k) {
                    A[i * N + k] /= scale;
                    h += A[i * N + k] * A[i * N + k];
                }
                double f = A[i * N + l];
                double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A[i * N + l] = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A[j * N + i] = A[i * N + j] / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A[j * N + k] * A[i * N + k];
                    for (int k = j + 1; k <= l; ++k) g += A[k * N + j] * A[i * N + k];
                    e[j] = g / h;
                    f += e[j] * A[i * N + j];
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A[i * N + j];
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A[j * N + k] -= f * e[k] + g * A[i * N + k];
                }
            }
        } else {
            e[i] = A[i * N + l];
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < N; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A[i * N + k] * A[k * N + j];
                for (int k = 0; k <= l; ++k) A[k * N + j] -= g * A[k * N + i];
            }
        }
        d[i] = A[i * N + i];
        A[i * N + i] = 1.0;
        for (int j = 0; j <= l; ++j) A[j * N + i] = A[i * N + j] = 0.0;
    }
    // ---- tqli: eigenvalues in d, eigenvectors in the columns of A
    for (int i = 1; i < N; ++i) e[i - 1] = e[i];
    e[N - 1] = 0.0;
    double anorm = 0.0;
    for (int i = 0; i < N; ++i) anorm = fmax(anorm, fabs(d[i]) + fabs(e[i]));
    for (int l = 0; l < N; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < N - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.220446049250313e-16 * dd || fabs(e[m]) <= 1e-18 * anorm) break;
            }
            if (m != l) {
                if (iter++ == 60) break;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double sn = 1.0, c = 1.0, pp = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = sn * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= pp;
                        e[m] = 0.0;
                        break;
                    }
                    sn = f / r;
                    c = g / r;
                    g = d[i + 1] - pp;
                    r = (d[i] - g) * sn + 2.0 * c * b;
                    d[i + 1] = g + (pp = sn * r);
                    g = c * r - b;
                    for (int k = 0; k < N; ++k) {
                        f = A[k * N + i + 1];
                        A[k * N + i + 1] = sn * A[k * N + i] + c * f;
                        A[k * N + i] = c * A[k * N + i] - sn * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= pp;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // ---- sort descending, eigenvectors as rows of Vt
    int order[N];
    for (int i = 0; i < N; ++i) order[i] = i;
    for (int i = 0; i < N - 1; ++i) {
        int m = i;
        for (int j = i + 1; j < N; ++j)
            if (d[order[j]] > d[order[m]]) m = j;
        const int t = order[i]; order[i] = order[m]; order[m] = t;
    }
    for (int i = 0; i < N; ++i) {
        w[i] = d[order[i]];
        if (i >= FIRST)
            for (int k = 0; k < N; ++k) Vt[(i - FIRST) * N + k] = A[k * N + order[i]];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bit-exact restatement of the OpenCV linear algebra EPnP goes through (modules/core/src/lapack.cpp of the image's
// cv2 4.13.0; its arithmetic was read off the shipped binary and every routine below is checked BITWISE against
// cv2.SVDecomp / cv2.invert(DECOMP_SVD) / cv2.solve(DECOMP_SVD) / cv2.mulTransposed in tests/test_epnp_host.py).
// Why bit-exact: with exactly 5 correspondences (the RANSAC minimal set) M^T M has a 2-D null space and EPnP reads the
// LEFT singular vectors, which for (numerically) zero singular values are the normalised rounding residue of the
// one-sided Jacobi iteration -- a chaotic function of every rounding upstream.  Matching cv2's hypotheses therefore
// needs the same operations in the same order: sequential (non-pairwise) sums, no FMA contraction (this header is
// compiled with -fmad=false / -ffp-contract=off), OpenCV's own hypot, its rotation formulas and its descending
// selection sort.
// Element (i,k) of a matrix lives at A[(i*cols + k) * S]: S = 1 on the host, S = blockDim on the device when the
// matrix is interleaved across the threads of a block in shared memory.

P2P_HD inline double cv_hypot(double a, double b) {   // lapack.cpp's local hypot, NOT libm's
    a = fabs(a);
    b = fabs(b);
    if (a > b) {
        b /= a;
        return a * sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * sqrt(1 + a * a);
    }
    return 0;
}

// JacobiSVDImpl_<double>: one-sided Jacobi on the N rows (length M) of At.  On return W = singular values
// (descending), rows i < n1 of At scaled to unit length (= rows of U^T), Vt (if not null) = V^T.
// One compact copy of the code serves every small instance (3x3, 6x3, 6x4, 6x5): sizes are run-time arguments and the
// matrices live in strided scratch memory (shared memory on the device), because fully unrolled per-size copies made the
// hypothesis kernel 500 KB of straight-line code that spent 44 % of its stall samples waiting for instruction fetches
// (profiles/r02d_hyp_*).
#if defined(__CUDA_ARCH__)
#define P2P_UNROLL _Pragma("unroll")
#define P2P_NOINLINE __noinline__
#else
#define P2P_UNROLL
#define P2P_NOINLINE
#endif
// Row length M is a compile-time constant (3 and 6 on the device: two copies of the code), so the two rows of a pair sit in
// registers while they are rotated; N, the pair loops and the V update stay run-time loops over the strided scratch.
template <int S, int M>
P2P_HD P2P_NOINLINE void cv_jacobi_svd_m(double* At, int N, double* W, double* Vt, int n1) {
    const double eps = 2.220446049250313e-16 * 10, minval = 2.2250738585072014e-308;
#define P2P_AT(i, k) At[((i) * M + (k)) * S]
#define P2P_VT(i, k) Vt[((i) * N + (k)) * S]
#define P2P_W(i) W[(i) * S]
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        P2P_UNROLL
        for (int k = 0; k < M; ++k) { const double t = P2P_AT(i, k); sd += t * t; }
        P2P_W(i) = sd;
        if (Vt)
            for (int k = 0; k < N; ++k) P2P_VT(i, k) = (k == i) ? 1 : 0;
    }
    const int max_iter = M > 30 ? M : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) {
                double ri[M], rj[M];
                double a = P2P_W(i), p = 0, b = P2P_W(j);
                P2P_UNROLL
                for (int k = 0; k < M; ++k) {
                    ri[k] = P2P_AT(i, k);
                    rj[k] = P2P_AT(j, k);
                    p += ri[k] * rj[k];
                }
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cv_hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                P2P_UNROLL
                for (int k = 0; k < M; ++k) {
                    const double t0 = c * ri[k] + s * rj[k];
                    const double t1 = -s * ri[k] + c * rj[k];
                    P2P_AT(i, k) = t0;
                    P2P_AT(j, k) = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                P2P_W(i) = a;
                P2P_W(j) = b;
                changed = true;
                if (Vt)
                    for (int k = 0; k < N; ++k) {
                        const double x = P2P_VT(i, k), y = P2P_VT(j, k);
                        P2P_VT(i, k) = c * x + s * y;
                        P2P_VT(j, k) = -s * x + c * y;
                    }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        P2P_UNROLL
        for (int k = 0; k < M; ++k) { const double t = P2P_AT(i, k); sd += t * t; }
        P2P_W(i) = sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {   // selection sort, descending: j = first maximum of W[i..]
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (P2P_W(j) < P2P_W(k)) j = k;
        if (i != j) {
            const double tw = P2P_W(i); P2P_W(i) = P2P_W(j); P2P_W(j) = tw;
            P2P_UNROLL
            for (int k = 0; k < M; ++k) { const double t = P2P_AT(i, k); P2P_AT(i, k) = P2P_AT(j, k); P2P_AT(j, k) = t; }
            if (Vt)
                for (int k = 0; k < N; ++k) { const double t = P2P_VT(i, k); P2P_VT(i, k) = P2P_VT(j, k); P2P_VT(j, k) = t; }
        }
    }
    // unit rows; a singular value <= DBL_MIN gets a vector regenerated from the fixed-seed RNG and orthogonalised
    // against the previous rows (exactly rank-deficient input, e.g. coplanar points with one coordinate constant)
    unsigned long long rng = 0x12345678ull;
    for (int i = 0; i < n1; ++i) {
        double sd = P2P_W(i);
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
            const double val0 = 1. / M;
            for (int k = 0; k < M; ++k) {
                rng = (unsigned long long)(unsigned)rng * 4164903690ull + (unsigned)(rng >> 32);
                P2P_AT(i, k) = ((unsigned)rng & 256) != 0 ? val0 : -val0;
            }
            for (int it = 0; it < 2; ++it)
                for (int j = 0; j < i; ++j) {
                    sd = 0;
                    for (int k = 0; k < M; ++k) sd += P2P_AT(i, k) * P2P_AT(j, k);
                    double asum = 0;
                    for (int k = 0; k < M; ++k) {
                        const double t = P2P_AT(i, k) - sd * P2P_AT(j, k);
                        P2P_AT(i, k) = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < M; ++k) P2P_AT(i, k) *= asum;
                }
            sd = 0;
            for (int k = 0; k < M; ++k) { const double t = P2P_AT(i, k); sd += t * t; }
            sd = sqrt(sd);
        }
        const double sc = sd > minval ? 1 / sd : 0.;
        P2P_UNROLL
        for (int k = 0; k < M; ++k) P2P_AT(i, k) *= sc;
    }
#undef P2P_AT
#undef P2P_VT
#undef P2P_W
}

// run-time row length: 3 and 6 are what EPnP uses; 12 only in the host tests (the device has cv_jacobi_svd12_ut)
template <int S>
P2P_HD inline void cv_jacobi_svd(double* At, int M, int N, double* W, double* Vt, int n1) {
    if (M == 3) cv_jacobi_svd_m<S, 3>(At, N, W, Vt, n1);
    else if (M == 6) cv_jacobi_svd_m<S, 6>(At, N, W, Vt, n1);
#if !defined(__CUDA_ARCH__)
    else if (M == 12) cv_jacobi_svd_m<S, 12>(At, N, W, Vt, n1);
#endif
}

// The same routine specialised for what EPnP asks of the 12 x 12 M^T M (U^T only), laid out for a GPU thread: row i
// stays in registers while it is paired with rows i+1.., and the squared row norms W[] are not kept in memory at all --
// OpenCV's W[i] is by construction the k-ordered sum of squares of the row as stored, so recomputing it next to the dot
// product (three independent add chains) gives the same bits and removes a dynamically indexed array (= local memory)
// from the inner loop.  Bitwise identical to cv_jacobi_svd<12, 12, false, 12> (tests/test_epnp_host.py).
template <int S = 1>
P2P_HD inline void cv_jacobi_svd12_ut(double* At, double* W) {
    const double eps = 2.220446049250313e-16 * 10, minval = 2.2250738585072014e-308;
    constexpr int N = 12;
#define P2P_AT(i, k) At[((i) * N + (k)) * S]
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i) {
            double ri[N];
P2P_UNROLL
            for (int k = 0; k < N; ++k) ri[k] = P2P_AT(i, k);
            bool dirty = false;
            for (int j = i + 1; j < N; ++j) {
                double rj[N];
P2P_UNROLL
                for (int k = 0; k < N; ++k) rj[k] = P2P_AT(j, k);
                double a = 0, b = 0, p = 0;
P2P_UNROLL
                for (int k = 0; k < N; ++k) {
                    a += ri[k] * ri[k];
                    b += rj[k] * rj[k];
                    p += ri[k] * rj[k];
                }
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cv_hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
P2P_UNROLL
                for (int k = 0; k < N; ++k) {
                    const double x = ri[k], y = rj[k];
                    ri[k] = c * x + s * y;
                    P2P_AT(j, k) = -s * x + c * y;
                }
                dirty = true;
            }
            if (dirty) {
P2P_UNROLL
                for (int k = 0; k < N; ++k) P2P_AT(i, k) = ri[k];
                changed = true;
            }
        }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
P2P_UNROLL
        for (int k = 0; k < N; ++k) { const double t = P2P_AT(i, k); sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < N; ++k) { const double t = P2P_AT(i, k); P2P_AT(i, k) = P2P_AT(j, k); P2P_AT(j, k) = t; }
        }
    }
    unsigned long long rng = 0x12345678ull;
    for (int i = 0; i < N; ++i) {
        double sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {   // exactly zero singular value: see cv_jacobi_svd
            const double val0 = 1. / N;
            for (int k = 0; k < N; ++k) {
                rng = (unsigned long long)(unsigned)rng * 4164903690ull + (unsigned)(rng >> 32);
                P2P_AT(i, k) = ((unsigned)rng & 256) != 0 ? val0 : -val0;
            }
            for (int it = 0; it < 2; ++it)
                for (int j = 0; j < i; ++j) {
                    sd = 0;
                    for (int k = 0; k < N; ++k) sd += P2P_AT(i, k) * P2P_AT(j, k);
                    double asum = 0;
                    for (int k = 0; k < N; ++k) {
                        const double t = P2P_AT(i, k) - sd * P2P_AT(j, k);
                        P2P_AT(i, k) = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < N; ++k) P2P_AT(i, k) *= asum;
                }
            sd = 0;
            for (int k = 0; k < N; ++k) { const double t = P2P_AT(i, k); sd += t * t; }
            sd = sqrt(sd);
        }
        const double sc = sd > minval ? 1 / sd : 0.;
P2P_UNROLL
        for (int k = 0; k < N; ++k) P2P_AT(i, k) *= sc;
    }
#undef P2P_AT
}

// Scratch convention: every routine below that needs an SVD takes `ws`, >= kWs doubles with element stride S.
constexpr int kWs = 60;   // 6x5 solve: at 30 + vt 25 + w 5

// cv::SVD::compute(A, w, u, vt) for a 3x3 A (row-major, dense): results in ws -- ut = U^T at [0,9) (rows = left singular
// vectors), vt = V^T at [9,18), w at [18,21).  want_v = false: cvSVD(A, W, Ut, 0, CV_SVD_U_T).
template <int S>
P2P_HD inline void cv_svd3(const double* A, double* ws, bool want_v = true) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ws[(i * 3 + j) * S] = A[j * 3 + i];
    cv_jacobi_svd<S>(ws, 3, 3, ws + 18 * S, want_v ? ws + 9 * S : nullptr, 3);
}

// cvInvert(A, Ainv, CV_SVD) for 3x3: SVD::compute + SVD::backSubst (SVBkSbImpl_ with an identity right-hand side).
template <int S>
P2P_HD inline void cv_invert3_svd(const double* A, double* x, double* ws) {
    cv_svd3<S>(A, ws);
    const double* ut = ws;
    const double* vt = ws + 9 * S;
    const double* w = ws + 18 * S;
    for (int i = 0; i < 9; ++i) x[i] = 0;
    double thr = 0;
    for (int i = 0; i < 3; ++i) thr += w[i * S];
    thr *= 2.220446049250313e-16 * 2;
    for (int i = 0; i < 3; ++i) {
        double wi = w[i * S];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double buf[3];
        for (int j = 0; j < 3; ++j) buf[j] = ut[(i * 3 + j) * S] * wi;
        for (int r = 0; r < 3; ++r) {
            const double sv = vt[(i * 3 + r) * S];
            for (int j = 0; j < 3; ++j) x[r * 3 + j] = x[r * 3 + j] + sv * buf[j];
        }
    }
}

// cvSolve(A, b, x, CV_SVD) for the 6 x N system (N <= 5) whose column c is column cols[c] of the 6 x 10 matrix `l`
// (element stride S): minimum-norm least squares, singular values <= 2*DBL_EPSILON*sum(w) dropped.
template <int S>
P2P_HD inline void cv_solve6_svd(const double* l, const int* cols, int N, const double* b, double* x, double* ws) {
    double* at = ws;
    double* vt = ws + 6 * N * S;
    double* w = vt + N * N * S;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < 6; ++j) at[(i * 6 + j) * S] = l[(j * 10 + cols[i]) * S];
    cv_jacobi_svd<S>(at, 6, N, w, vt, N);
    for (int i = 0; i < N; ++i) x[i] = 0;
    double thr = 0;
    for (int i = 0; i < N; ++i) thr += w[i * S];
    thr *= 2.220446049250313e-16 * 2;
    for (int i = 0; i < N; ++i) {
        double wi = w[i * S];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double sum = 0;
        for (int j = 0; j < 6; ++j) sum += at[(i * 6 + j) * S] * b[j];
        sum *= wi;
        for (int j = 0; j < N; ++j) x[j] = x[j] + sum * vt[(i * N + j) * S];
    }
}

// Householder QR least squares for the 6x4 Gauss-Newton system (epnp.cpp qr_solve).
P2P_HD inline bool qr_solve_6x4(double* A, double* b, double* x) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    for (int k = 0; k < nc; ++k) {
        // epnp.cpp's pivot scan reads each element BEFORE advancing its pointer, so it covers rows k .. nr-2 only
        double eta = fabs(A[k * nc + k]);
        for (int i = k; i < nr - 1; ++i) eta = fmax(eta, fabs(A[i * nc + k]));
        if (eta == 0.0) return false;  // singular: OpenCV leaves X untouched
        const double inv_eta = 1. / eta;
        double sigma = 0;
        for (int i = k; i < nr; ++i) {
            A[i * nc + k] *= inv_eta;
            sigma += A[i * nc + k] * A[i * nc + k];
        }
        sigma = sqrt(sigma);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; ++j) {
            double sum = 0;
            for (int i = k; i < nr; ++i) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
            for (int i = k; i < nr; ++i) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
    for (int j = 0; j < nc; ++j) {  // b <- Q^T b
        double sum = 0;
        for (int i = j; i < nr; ++i) sum += A[i * nc + j] * b[i];
        const double tau = sum / A1[j];
        for (int i = j; i < nr; ++i) b[i] -= tau * A[i * nc + j];
    }
    x[nc - 1] = b[nc - 1] / A2[nc - 1];  // back substitution
    for (int i = nc - 2; i >= 0; --i) {
        double sum = 0;
        for (int j = i + 1; j < nc; ++j) sum += A[i * nc + j] * x[j];
        x[i] = (b[i] - sum) / A2[i];
    }
    return true;
}

// ---- control points (epnp.cpp choose_control_points) from the centroid and the 3x3 scatter
// PW0^T PW0 of the n reference points.
template <int S>
P2P_HD inline void choose_control_points(const double* centroid, const double* scatter, int n, double cws[4][3], double* ws) {
    cv_svd3<S>(scatter, ws, false);
    for (int j = 0; j < 3; ++j) cws[0][j] = centroid[j];
    for (int i = 1; i < 4; ++i) {
        const double k = sqrt(ws[(18 + i - 1) * S] / n);
        for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * ws[(3 * (i - 1) + j) * S];
    }
}

// cc_inv of compute_barycentric_coordinates: cvInvert(CC, CC_inv, CV_SVD) of [c1-c0 c2-c0 c3-c0].
template <int S>
P2P_HD inline void control_inverse(const double cws[4][3], double* ci, double* ws) {
    double cc[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    cv_invert3_svd<S>(cc, ci, ws);
}

P2P_HD inline void barycentric(const double* ci, const double cws[4][3], const double* p, double* a) {
    for (int j = 0; j < 3; ++j)
        a[1 + j] = ci[3 * j] * (p[0] - cws[0][0]) + ci[3 * j + 1] * (p[1] - cws[0][1]) + ci[3 * j + 2] * (p[2] - cws[0][2]);
    a[0] = 1.0 - a[1] - a[2] - a[3];
}

// ut8 = rows 8..11 of the 12 x 12 eigenvector matrix (4 x 12): the null-space candidates, smallest eigenvalue last.
// `l` (6 x 10, element stride S) receives L_6x10.
template <int S>
P2P_HD inline void compute_L_6x10(const double* ut8, double* l) {
    const double* v[4] = {ut8 + 12 * 3, ut8 + 12 * 2, ut8 + 12 * 1, ut8};
    int a = 0, b = 1;
    for (int i = 0; i < 6; ++i) {
        double dv[4][3];
        for (int q = 0; q < 4; ++q)
            for (int c = 0; c < 3; ++c) dv[q][c] = v[q][3 * a + c] - v[q][3 * b + c];
        double* r = l + 10 * i * S;
        r[0 * S] = dot3(dv[0], dv[0]);
        r[1 * S] = 2.0 * dot3(dv[0], dv[1]);
        r[2 * S] = dot3(dv[1], dv[1]);
        r[3 * S] = 2.0 * dot3(dv[0], dv[2]);
        r[4 * S] = 2.0 * dot3(dv[1], dv[2]);
        r[5 * S] = dot3(dv[2], dv[2]);
        r[6 * S] = 2.0 * dot3(dv[0], dv[3]);
        r[7 * S] = 2.0 * dot3(dv[1], dv[3]);
        r[8 * S] = 2.0 * dot3(dv[2], dv[3]);
        r[9 * S] = dot3(dv[3], dv[3]);
        ++b;
        if (b > 3) { ++a; b = a + 1; }
    }
}

P2P_HD inline void compute_rho(const double cws[4][3], double* rho) {
    rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
    rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
}

// betas10 = [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44]
template <int S>
P2P_HD inline void find_betas_approx_1(const double* l, const double* rho, double* betas, double* ws) {  // [B11 B12 B13 B14]
    const int cols[4] = {0, 1, 3, 6};
    double b4[4];
    cv_solve6_svd<S>(l, cols, 4, rho, b4, ws);
    if (b4[0] < 0) { betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0]; }
    else { betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0]; }
}
template <int S>
P2P_HD inline void find_betas_approx_2(const double* l, const double* rho, double* betas, double* ws) {  // [B11 B12 B22]
    const int cols[3] = {0, 1, 2};
    double b3[3];
    cv_solve6_svd<S>(l, cols, 3, rho, b3, ws);
    if (b3[0] < 0) { betas[0] = sqrt(-b3[0]); betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
    else { betas[0] = sqrt(b3[0]); betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0; betas[3] = 0.0;
}
template <int S>
P2P_HD inline void find_betas_approx_3(const double* l, const double* rho, double* betas, double* ws) {  // [B11 B12 B22 B13 B23]
    const int cols[5] = {0, 1, 2, 3, 4};
    double b5[5];
    cv_solve6_svd<S>(l, cols, 5, rho, b5, ws);
    if (b5[0] < 0) { betas[0] = sqrt(-b5[0]); betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
    else { betas[0] = sqrt(b5[0]); betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
    betas[3] = 0.0;
}

template <int S>
P2P_HD P2P_NOINLINE void gauss_newton(const double* l, const double* rho, double* betas) {
    double x[4] = {0, 0, 0, 0};
    for (int it = 0; it < 5; ++it) {
        double A[24], b[6];
        for (int i = 0; i < 6; ++i) {
            double r[10];
            for (int c = 0; c < 10; ++c) r[c] = l[(i * 10 + c) * S];
            A[i * 4 + 0] = 2 * r[0] * betas[0] + r[1] * betas[1] + r[3] * betas[2] + r[6] * betas[3];
            A[i * 4 + 1] = r[1] * betas[0] + 2 * r[2] * betas[1] + r[4] * betas[2] + r[7] * betas[3];
            A[i * 4 + 2] = r[3] * betas[0] + r[4] * betas[1] + 2 * r[5] * betas[2] + r[8] * betas[3];
            A[i * 4 + 3] = r[6] * betas[0] + r[7] * betas[1] + r[8] * betas[2] + 2 * r[9] * betas[3];
            b[i] = rho[i] - (r[0] * betas[0] * betas[0] + r[1] * betas[0] * betas[1] + r[2] * betas[1] * betas[1] +
                             r[3] * betas[0] * betas[2] + r[4] * betas[1] * betas[2] + r[5] * betas[2] * betas[2] +
                             r[6] * betas[0] * betas[3] + r[7] * betas[1] * betas[3] + r[8] * betas[2] * betas[3] +
                             r[9] * betas[3] * betas[3]);
        }
        qr_solve_6x4(A, b, x);
        for (int i = 0; i < 4; ++i) betas[i] += x[i];
    }
}

// The three refined beta sets from the four null-space candidates ut8 (rows 8..11 of U^T) -- epnp.cpp compute_pose,
// middle part.  `lws`: scratch of 60 + kWs doubles (element stride S): L_6x10, then the SVD workspace.
template <int S>
P2P_HD inline void betas_from_ut(const double* ut8, const double cws[4][3], double betas[3][4], double* lws) {
    double rho[6];
    double* l = lws;
    double* ws = lws + 60 * S;
    compute_L_6x10<S>(ut8, l);
    compute_rho(cws, rho);
    find_betas_approx_1<S>(l, rho, betas[0], ws); gauss_newton<S>(l, rho, betas[0]);
    find_betas_approx_2<S>(l, rho, betas[1], ws); gauss_newton<S>(l, rho, betas[1]);
    find_betas_approx_3<S>(l, rho, betas[2], ws); gauss_newton<S>(l, rho, betas[2]);
}

// Small-n path, bit-exact: cvSVD(MtM, D, Ut, 0, CV_SVD_MODIFY_A | CV_SVD_U_T) in place on `mtm` (144 doubles, element
// stride S); rows 8..11 are copied out and the storage is then reused as the L_6x10 / SVD scratch.
template <int S = 1>
P2P_HD inline void solve_betas_exact(double* mtm, const double cws[4][3], double* ut8, double betas[3][4]) {
    double w[12];
    cv_jacobi_svd12_ut<S>(mtm, w);   // MtM is bitwise symmetric, so A^T = A
    for (int i = 0; i < 48; ++i) ut8[i] = mtm[(96 + i) * S];
    betas_from_ut<S>(ut8, cws, betas, mtm);
}

// cvMulTransposed(M, MtM, 1) for the 2n x 12 matrix fill_M builds, restated on its sparsity: per entry the same
// products in the same (row-sequential) order; the structural zeros contribute exact +-0 and are skipped.
// `mtm` (stride S) receives the full symmetric matrix.
template <int S = 1>
P2P_HD inline void mtm_exact(const double* alphas, const double* us, int n, const Cam& cam, double* mtm) {
    for (int x = 0; x < 4; ++x)
        for (int y = x; y < 4; ++y) {
            double e00 = 0, e02 = 0, e20 = 0, e11 = 0, e12 = 0, e21 = 0, e22 = 0;
            for (int i = 0; i < n; ++i) {
                const double ax = alphas[4 * i + x], ay = alphas[4 * i + y];
                const double du = cam.uc - us[2 * i], dv = cam.vc - us[2 * i + 1];
                const double x0 = ax * cam.fu, x1 = ax * cam.fv, xu = ax * du, xv = ax * dv;
                const double y0 = ay * cam.fu, y1 = ay * cam.fv, yu = ay * du, yv = ay * dv;
                e00 += x0 * y0;
                e02 += x0 * yu;
                e20 += xu * y0;
                e11 += x1 * y1;
                e12 += x1 * yv;
                e21 += xv * y1;
                e22 += xu * yu;
                e22 += xv * yv;
            }
            const double blk[9] = {e00, 0, e02, 0, e11, e12, e20, e21, e22};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    if (x == y && c < r) continue;   // lower triangle of a diagonal block = mirror of its upper one
                    mtm[((3 * x + r) * 12 + 3 * y + c) * S] = blk[r * 3 + c];
                    mtm[((3 * y + c) * 12 + 3 * x + r) * S] = blk[r * 3 + c];
                }
        }
}

P2P_HD inline void compute_ccs(const double* betas, const double* ut8, double ccs[4][3]) {
    for (int i = 0; i < 4; ++i) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double* v = ut8 + 12 * (3 - i);
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[3 * j + k];
    }
}

P2P_HD inline void camera_point(const double* a, const double ccs[4][3], double* pc) {
    for (int j = 0; j < 3; ++j) pc[j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
}

// estimate_R_and_t from the centroids and the 3x3 correlation ABt = sum (pc-pc0)(pw-pw0)^T:
// cvSVD(ABt, D, U, V, CV_SVD_MODIFY_A), R = U V^T with the det < 0 fix on the last row.
template <int S>
P2P_HD inline void rt_from_correlation(const double* abt, const double* pc0, const double* pw0, double R[3][3], double* t, double* ws) {
    cv_svd3<S>(abt, ws);
    // U[i][k] = ut[k][i], V[j][k] = vt[k][j]
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i][j] = ws[i * S] * ws[(9 + j) * S] + ws[(3 + i) * S] * ws[(12 + j) * S] + ws[(6 + i) * S] * ws[(15 + j) * S];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    t[0] = pc0[0] - dot3(R[0], pw0);
    t[1] = pc0[1] - dot3(R[1], pw0);
    t[2] = pc0[2] - dot3(R[2], pw0);
}

P2P_HD inline double reproj_dist(const double R[3][3], const double* t, const double* pw, double u, double v, const Cam& cam) {
    const double Xc = dot3(R[0], pw) + t[0], Yc = dot3(R[1], pw) + t[1], inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
    const double ue = cam.uc + cam.fu * Xc * inv_Zc, ve = cam.vc + cam.fv * Yc * inv_Zc;
    return sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
}

// cv::Rodrigues, matrix -> vector (after projecting R onto SO(3) through its SVD) and back.
template <int S>
P2P_HD inline void rodrigues_to_vec(const double Rin[3][3], double* r, double* ws) {
    double A[9], R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i * 3 + j] = Rin[i][j];
    cv_svd3<S>(A, ws);   // SVD::compute(R, W, U, Vt); R = U*Vt
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i * 3 + j] = ws[i * S] * ws[(9 + j) * S] + ws[(3 + i) * S] * ws[(12 + j) * S] + ws[(6 + i) * S] * ws[(15 + j) * S];
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double sn = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : (c < -1. ? -1. : c);
    double theta = acos(c);
    if (sn < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; }
        else {
            double t;
            t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.));
            t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = theta * rx; r[1] = theta * ry; r[2] = theta * rz;
        }
    } else {
        const double vth = 1 / (2 * sn) * theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

P2P_HD inline void rodrigues_to_mat(const double* r, double R[3][3]) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < 2.220446049250313e-16) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    const double x = r[0] * itheta, y = r[1] * itheta, z = r[2] * itheta;
    R[0][0] = c + c1 * (x * x);       R[0][1] = c1 * (x * y) - s * z; R[0][2] = c1 * (x * z) + s * y;
    R[1][0] = c1 * (x * y) + s * z; R[1][1] = c + c1 * (y * y);       R[1][2] = c1 * (y * z) - s * x;
    R[2][0] = c1 * (x * z) - s * y; R[2][1] = c1 * (y * z) + s * x; R[2][2] = c + c1 * (z * z);
}

// Large-n refit (csrc/pnp_ransac.cu epnp_refit_kernel): everything EPnP needs from the n inliers are 52 sums, taken by a
// block reduction --
//   s[k], s[10+k], s[20+k], s[30+k], k = pair (x <= y) of control points: sum of a_x a_y {1, du, dv, du^2 + dv^2}
//   (du = uc - u, dv = vc - v): M^T M in structured form;   s[40 + 3j + c] = sum a_j (pw_c - c0_c): correlation terms.
// The serial tail turns them into the three candidate poses in two steps: refit_prepare (one thread: eigenvectors, L_6x10,
// rho and the sums every candidate shares) and refit_candidate (independent per candidate c = 0, 1, 2: beta initialisation
// c, Gauss-Newton, R|t -- three threads of the block run them side by side).  pf = first inlier (solve_for_sign looks at
// point 0).
struct RefitShared {
    double ut8[48], l[60], rho[6], af[4], Srow[4], dsum[3];
};

P2P_HD inline void refit_prepare(const double* s, const double cws[4][3], const double* ci, const Cam& cam, const double* pf,
                                 RefitShared& sh) {
    double mtm[144], w[12];
    int k = 0;
    for (int x = 0; x < 4; ++x)
        for (int y = x; y < 4; ++y) {
            const double S = s[k], Su = s[10 + k], Sv = s[20 + k], Sq = s[30 + k];
            const double blk[9] = {cam.fu * cam.fu * S, 0, cam.fu * Su, 0, cam.fv * cam.fv * S, cam.fv * Sv, cam.fu * Su, cam.fv * Sv, Sq};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    mtm[(3 * x + r) * 12 + 3 * y + c] = blk[r * 3 + c];
                    mtm[(3 * y + c) * 12 + 3 * x + r] = blk[r * 3 + c];
                }
            ++k;
        }
    // the eigenvectors of the four smallest eigenvalues of M^T M by tridiagonal QL (see tridiag_eig_sym)
    tridiag_eig_sym<12, 8>(mtm, sh.ut8, w);
    compute_L_6x10<1>(sh.ut8, sh.l);
    compute_rho(cws, sh.rho);
    barycentric(ci, cws, pf, sh.af);
    // S_j = sum_i alpha_ij: alpha sums to 1 per point, so it is the row sum of the pair table
    for (int j = 0; j < 4; ++j) sh.Srow[j] = 0;
    k = 0;
    for (int x = 0; x < 4; ++x)
        for (int y = x; y < 4; ++y) {
            sh.Srow[x] += s[k];
            if (y != x) sh.Srow[y] += s[k];
            ++k;
        }
    for (int q = 0; q < 3; ++q) sh.dsum[q] = 0;   // sum (pw - c0): ~0 (rounding only), kept for fidelity
    for (int j = 0; j < 4; ++j)
        for (int q = 0; q < 3; ++q) sh.dsum[q] += s[40 + j * 3 + q];
}

P2P_HD inline void refit_candidate(int c, const double* s, int m, const double* c0, const RefitShared& sh, double R[3][3], double* t) {
    double betas[4], ws[kWs];
    if (c == 0) find_betas_approx_1<1>(sh.l, sh.rho, betas, ws);
    else if (c == 1) find_betas_approx_2<1>(sh.l, sh.rho, betas, ws);
    else find_betas_approx_3<1>(sh.l, sh.rho, betas, ws);
    gauss_newton<1>(sh.l, sh.rho, betas);
    double ccs[4][3], pc[3];
    compute_ccs(betas, sh.ut8, ccs);
    camera_point(sh.af, ccs, pc);
    const double sg = pc[2] < 0.0 ? -1.0 : 1.0;
    // alphas are affine in pw:  pc0 = sum_j (S_j / m) ccs_j,  pw0 = c0,  ABt = sum_j ccs_j T_j^T - pc0 (x) sum(pw - c0)
    double pc0[3] = {0, 0, 0}, abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 3; ++r) ccs[j][r] *= sg;
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 3; ++r) pc0[r] += sh.Srow[j] / m * ccs[j][r];
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) abt[3 * r + q] += ccs[j][r] * s[40 + j * 3 + q];
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) abt[3 * r + q] -= pc0[r] * sh.dsum[q];
    rt_from_correlation<1>(abt, pc0, c0, R, t, ws);
}

// Whole EPnP for a small point set held by one thread (the 5-point RANSAC hypotheses, and refits on <= MAXN inliers),
// bit-exact with cv2.solvePnP(SOLVEPNP_EPNP) up to libm's sin/cos/acos in Rodrigues.
// pws: n x 3 object points, us: n x 2 pixel coordinates.  `mtm` = 144-element scratch with element stride S (shared
// memory on the device).  Returns the winning R, t.
template <int MAXN, int S = 1>
P2P_HD inline void solve_small(const double* pws, const double* us, int n, const Cam& cam, double* mtm, double R[3][3], double* t) {
    double cws[4][3], c0[3] = {0, 0, 0}, sc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) c0[j] += pws[3 * i + j];
    for (int j = 0; j < 3; ++j) c0[j] /= n;
    for (int i = 0; i < n; ++i) {   // cvMulTransposed(PW0, PW0tPW0, 1)
        const double d[3] = {pws[3 * i] - c0[0], pws[3 * i + 1] - c0[1], pws[3 * i + 2] - c0[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sc[a * 3 + b] += d[a] * d[b];
    }
    choose_control_points<S>(c0, sc, n, cws, mtm);   // `mtm` doubles as the SVD scratch until M^T M is built
    double ci[9], alphas[MAXN * 4];
    control_inverse<S>(cws, ci, mtm);
    for (int i = 0; i < n; ++i) barycentric(ci, cws, pws + 3 * i, alphas + 4 * i);
    mtm_exact<S>(alphas, us, n, cam, mtm);
    double ut[48], betas[3][4];   // rows 8..11 of U^T only
    solve_betas_exact<S>(mtm, cws, ut, betas);
    double best = 0;
    for (int k = 0; k < 3; ++k) {
        double ccs[4][3], pcs[MAXN * 3];
        compute_ccs(betas[k], ut, ccs);
        for (int i = 0; i < n; ++i) camera_point(alphas + 4 * i, ccs, pcs + 3 * i);
        if (pcs[2] < 0.0)  // solve_for_sign
            for (int i = 0; i < 3 * n; ++i) pcs[i] = -pcs[i];
        double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0}, abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < 3; ++j) { pc0[j] += pcs[3 * i + j]; pw0[j] += pws[3 * i + j]; }
        for (int j = 0; j < 3; ++j) { pc0[j] /= n; pw0[j] /= n; }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < 3; ++j)
                for (int m = 0; m < 3; ++m) abt[3 * j + m] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i + m] - pw0[m]);
        double Rk[3][3], tk[3];
        rt_from_correlation<S>(abt, pc0, pw0, Rk, tk, mtm + 60 * S);
        double err = 0;
        for (int i = 0; i < n; ++i) err += reproj_dist(Rk, tk, pws + 3 * i, us[2 * i], us[2 * i + 1], cam);
        err /= n;
        // compute_pose: N=1; if (e2 < e1) N=2; if (e3 < e[N]) N=3  (NaN never wins)
        if (k == 0 || err < best) {
            best = err;
            for (int i = 0; i < 3; ++i) { t[i] = tk[i]; for (int j = 0; j < 3; ++j) R[i][j] = Rk[i][j]; }
        }
    }
}

}  // namespace epnp

// OpenCV's RNG (cv::RNG: multiply-with-carry) as RANSACPointSetRegistrator seeds it: RNG rng((uint64)-1).
struct CvRng {
    unsigned long long state;
    P2P_HD CvRng() : state(0xFFFFFFFFFFFFFFFFull) {}
    P2P_HD unsigned next() {
        state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    P2P_HD int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// RANSACPointSetRegistrator::getSubset for modelPoints = 5: distinct indices by rejection.
P2P_HD inline void ransac_subset5(CvRng& rng, int count, int* idx) {
    for (int i = 0; i < 5; ++i) {
        int v;
        for (;;) {
            v = rng.uniform(0, count);
            bool dup = false;
            for (int j = 0; j < i; ++j) dup = dup || (idx[j] == v);
            if (!dup) break;
        }
        idx[i] = v;
    }
}

// cv::RANSACUpdateNumIters
P2P_HD inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, 2.2250738585072014e-308);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

}  // namespace p2p
