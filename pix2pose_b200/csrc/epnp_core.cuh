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
// error wins.
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
// tred2 / tqli pair): ~10x fewer operations than cyclic Jacobi for N = 12, which matters because the RANSAC
// kernel runs one of these per hypothesis per thread.  Same contract as jacobi_eig_sym: `A` (row-major) is
// destroyed, w descending, row i of `Vt` = unit eigenvector of w[i].
// FIRST > 0 keeps only the eigenvectors FIRST .. N-1 (the smallest eigenvalues): row i of the full Vt lands in row
// i - FIRST of `Vt` ((N - FIRST) x N).  EPnP only ever reads the last four, and the RANSAC kernel is bound by the
// per-thread local-memory footprint, so the 5-point solver does not materialise the other eight.
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

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (row-major `A`, destroyed).
// On return w[0] >= w[1] >= ... and row i of `Vt` is the unit eigenvector of w[i]
// (the layout of cvSVD(..., CV_SVD_U_T) that epnp.cpp indexes as ut + 12*i).
template <int N>
P2P_HD inline void jacobi_eig_sym(double* A, double* Vt, double* w) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Vt[i * N + j] = (i == j) ? 1.0 : 0.0;
    // Rotations whose off-diagonal element is below eps * ||A||_F are skipped; a sweep without any
    // rotation ends the iteration (rank-deficient 5-point systems would otherwise spin on noise).
    double fro = 0.0;
    for (int i = 0; i < N * N; ++i) fro += A[i] * A[i];
    const double tiny = 1e-17 * sqrt(fro);
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (fabs(apq) <= tiny) continue;
                rotated = true;
                const double app = A[p * N + p], aqq = A[q * N + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {  // columns p,q
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {  // rows p,q
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {  // eigenvectors (stored as rows)
                    const double vpk = Vt[p * N + k], vqk = Vt[q * N + k];
                    Vt[p * N + k] = c * vpk - s * vqk;
                    Vt[q * N + k] = s * vpk + c * vqk;
                }
            }
        if (!rotated) break;
    }
    for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
    for (int i = 0; i < N - 1; ++i) {  // selection sort, descending
        int m = i;
        for (int j = i + 1; j < N; ++j)
            if (w[j] > w[m]) m = j;
        if (m != i) {
            const double tw = w[i]; w[i] = w[m]; w[m] = tw;
            for (int k = 0; k < N; ++k) { const double tv = Vt[i * N + k]; Vt[i * N + k] = Vt[m * N + k]; Vt[m * N + k] = tv; }
        }
    }
}

// One-sided (Hestenes) Jacobi SVD of an M x N matrix (row-major, N <= M <= 6): A = U diag(s) V^T.
// `A` is overwritten by U*diag(s) column-wise; returns s (descending) and V (N x N, row-major,
// columns = right singular vectors) with U's columns normalised in place where s > 0.
template <int M, int N>
P2P_HD inline void jacobi_svd(double* A, double* s, double* V) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < M; ++k) {
                    a += A[k * N + p] * A[k * N + p];
                    b += A[k * N + q] * A[k * N + q];
                    g += A[k * N + p] * A[k * N + q];
                }
                if (g == 0.0 || fabs(g) <= 1e-16 * sqrt(a * b)) continue;
                rotated = true;
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int k = 0; k < M; ++k) {
                    const double x = A[k * N + p], y = A[k * N + q];
                    A[k * N + p] = c * x - sn * y;
                    A[k * N + q] = sn * x + c * y;
                }
                for (int k = 0; k < N; ++k) {
                    const double x = V[k * N + p], y = V[k * N + q];
                    V[k * N + p] = c * x - sn * y;
                    V[k * N + q] = sn * x + c * y;
                }
            }
        if (!rotated) break;
    }
    for (int j = 0; j < N; ++j) {
        double n2 = 0;
        for (int k = 0; k < M; ++k) n2 += A[k * N + j] * A[k * N + j];
        s[j] = sqrt(n2);
    }
    for (int i = 0; i < N - 1; ++i) {  // sort descending, permuting columns of A and V
        int m = i;
        for (int j = i + 1; j < N; ++j)
            if (s[j] > s[m]) m = j;
        if (m != i) {
            const double ts = s[i]; s[i] = s[m]; s[m] = ts;
            for (int k = 0; k < M; ++k) { const double t = A[k * N + i]; A[k * N + i] = A[k * N + m]; A[k * N + m] = t; }
            for (int k = 0; k < N; ++k) { const double t = V[k * N + i]; V[k * N + i] = V[k * N + m]; V[k * N + m] = t; }
        }
    }
    for (int j = 0; j < N; ++j)
        if (s[j] > 0.0)
            for (int k = 0; k < M; ++k) A[k * N + j] /= s[j];
}

// Minimum-norm least squares x = pinv(A) b through the SVD, singular values below
// 2*DBL_EPSILON*sum(s) dropped (cv::solve(..., DECOMP_SVD) / SVD::backSubst).
template <int M, int N>
P2P_HD inline void svd_solve(const double* A_in, const double* b, double* x) {
    double A[M * N], s[N], V[N * N];
    for (int i = 0; i < M * N; ++i) A[i] = A_in[i];
    jacobi_svd<M, N>(A, s, V);
    double thr = 0;
    for (int j = 0; j < N; ++j) thr += s[j];
    thr *= 2.0 * 2.220446049250313e-16;
    for (int i = 0; i < N; ++i) x[i] = 0.0;
    for (int j = 0; j < N; ++j) {
        if (!(s[j] > thr)) continue;
        double ub = 0;
        for (int k = 0; k < M; ++k) ub += A[k * N + j] * b[k];
        ub /= s[j];
        for (int i = 0; i < N; ++i) x[i] += V[i * N + j] * ub;
    }
}

// Householder QR least squares for the 6x4 Gauss-Newton system (epnp.cpp qr_solve).
P2P_HD inline bool qr_solve_6x4(double* A, double* b, double* x) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    for (int k = 0; k < nc; ++k) {
        double eta = 0;
        for (int i = k; i < nr; ++i) eta = fmax(eta, fabs(A[i * nc + k]));
        if (eta == 0.0) return false;  // singular: OpenCV leaves X untouched
        double sigma = 0;
        for (int i = k; i < nr; ++i) {
            A[i * nc + k] /= eta;
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

// Left singular vectors of a 3x3 matrix exactly as OpenCV's JacobiSVDImpl_ (modules/core/src/lapack.cpp)
// produces them for cvSVD(A, W, Ut, 0, CV_SVD_U_T): one-sided Jacobi on the ROWS of A^T with its
// rotation formulas, descending sort, rows normalised.  EPnP's control points are c0 + k*u_i, so the
// SIGN convention of u_i changes the (noisy-data) solution at the 1e-3 level: it must be OpenCV's.
// ut: rows = singular vectors, w: singular values.
P2P_HD inline void cv_svd3_ut(const double* A, double* ut, double* w) {
    const double eps = 2.220446049250313e-16 * 10, minval = 2.2250738585072014e-308;
    double At[9], W[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < 3; ++k) sd += At[i * 3 + k] * At[i * 3 + k];
        W[i] = sd;
    }
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
        for (int i = 0; i < 2; ++i)
            for (int j = i + 1; j < 3; ++j) {
                double* Ai = At + i * 3;
                double* Aj = At + j * 3;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < 3; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, sn;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    sn = sqrt(delta / gamma);
                    c = p / (gamma * sn * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    sn = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < 3; ++k) {
                    const double t0 = c * Ai[k] + sn * Aj[k];
                    const double t1 = -sn * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
            }
        if (!changed) break;
    }
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < 3; ++k) sd += At[i * 3 + k] * At[i * 3 + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < 2; ++i) {
        int j = i;
        for (int k = i + 1; k < 3; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < 3; ++k) { const double t = At[i * 3 + k]; At[i * 3 + k] = At[j * 3 + k]; At[j * 3 + k] = t; }
        }
    }
    for (int i = 0; i < 3; ++i) {
        w[i] = W[i];
        // (OpenCV regenerates the vector from a fixed-seed RNG when W[i] <= DBL_MIN; such exactly
        //  degenerate point sets -- collinear / identical points -- are left as a zero vector here.)
        const double sc = W[i] > minval ? 1.0 / W[i] : 0.0;
        for (int k = 0; k < 3; ++k) ut[i * 3 + k] = At[i * 3 + k] * sc;
    }
}

// ---- control points (epnp.cpp choose_control_points) from the centroid and the 3x3 scatter
// PW0^T PW0 of the n reference points.
P2P_HD inline void choose_control_points(const double* centroid, const double* scatter, int n, double cws[4][3]) {
    double Ut[9], w[3];
    cv_svd3_ut(scatter, Ut, w);
    for (int j = 0; j < 3; ++j) cws[0][j] = centroid[j];
    for (int i = 1; i < 4; ++i) {
        const double k = sqrt(w[i - 1] / n);
        for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * Ut[3 * (i - 1) + j];
    }
}

// cc_inv of compute_barycentric_coordinates: pseudo-inverse (cvInvert CV_SVD) of [c1-c0 c2-c0 c3-c0].
P2P_HD inline void control_inverse(const double cws[4][3], double* ci) {
    double cc[9], s[3], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    jacobi_svd<3, 3>(cc, s, V);  // cc now holds U
    double thr = (s[0] + s[1] + s[2]) * 2.0 * 2.220446049250313e-16;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int k = 0; k < 3; ++k)
                if (s[k] > thr) acc += V[i * 3 + k] * cc[j * 3 + k] / s[k];
            ci[i * 3 + j] = acc;
        }
}

P2P_HD inline void barycentric(const double* ci, const double cws[4][3], const double* p, double* a) {
    for (int j = 0; j < 3; ++j)
        a[1 + j] = ci[3 * j] * (p[0] - cws[0][0]) + ci[3 * j + 1] * (p[1] - cws[0][1]) + ci[3 * j + 2] * (p[2] - cws[0][2]);
    a[0] = 1.0 - a[1] - a[2] - a[3];
}

// The two rows fill_M writes for one correspondence.
P2P_HD inline void m_rows(const double* a, double u, double v, const Cam& cam, double* m1, double* m2) {
    for (int i = 0; i < 4; ++i) {
        m1[3 * i] = a[i] * cam.fu; m1[3 * i + 1] = 0.0;           m1[3 * i + 2] = a[i] * (cam.uc - u);
        m2[3 * i] = 0.0;           m2[3 * i + 1] = a[i] * cam.fv; m2[3 * i + 2] = a[i] * (cam.vc - v);
    }
}

// ut8 = rows 8..11 of the 12 x 12 eigenvector matrix (4 x 12): the null-space candidates, smallest eigenvalue last
P2P_HD inline void compute_L_6x10(const double* ut8, double* l) {
    const double* v[4] = {ut8 + 12 * 3, ut8 + 12 * 2, ut8 + 12 * 1, ut8};
    double dv[4][6][3];
    for (int i = 0; i < 4; ++i) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; ++j) {
            dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
            dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
            dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
            ++b;
            if (b > 3) { ++a; b = a + 1; }
        }
    }
    for (int i = 0; i < 6; ++i) {
        double* r = l + 10 * i;
        r[0] = dot3(dv[0][i], dv[0][i]);
        r[1] = 2.0 * dot3(dv[0][i], dv[1][i]);
        r[2] = dot3(dv[1][i], dv[1][i]);
        r[3] = 2.0 * dot3(dv[0][i], dv[2][i]);
        r[4] = 2.0 * dot3(dv[1][i], dv[2][i]);
        r[5] = dot3(dv[2][i], dv[2][i]);
        r[6] = 2.0 * dot3(dv[0][i], dv[3][i]);
        r[7] = 2.0 * dot3(dv[1][i], dv[3][i]);
        r[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
        r[9] = dot3(dv[3][i], dv[3][i]);
    }
}

P2P_HD inline void compute_rho(const double cws[4][3], double* rho) {
    rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
    rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
}

// betas10 = [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44]
P2P_HD inline void find_betas_approx_1(const double* l, const double* rho, double* betas) {  // [B11 B12 B13 B14]
    double L[24], b4[4];
    for (int i = 0; i < 6; ++i) { L[i * 4] = l[i * 10]; L[i * 4 + 1] = l[i * 10 + 1]; L[i * 4 + 2] = l[i * 10 + 3]; L[i * 4 + 3] = l[i * 10 + 6]; }
    svd_solve<6, 4>(L, rho, b4);
    if (b4[0] < 0) { betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0]; }
    else { betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0]; }
}
P2P_HD inline void find_betas_approx_2(const double* l, const double* rho, double* betas) {  // [B11 B12 B22]
    double L[18], b3[3];
    for (int i = 0; i < 6; ++i) { L[i * 3] = l[i * 10]; L[i * 3 + 1] = l[i * 10 + 1]; L[i * 3 + 2] = l[i * 10 + 2]; }
    svd_solve<6, 3>(L, rho, b3);
    if (b3[0] < 0) { betas[0] = sqrt(-b3[0]); betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
    else { betas[0] = sqrt(b3[0]); betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0; betas[3] = 0.0;
}
P2P_HD inline void find_betas_approx_3(const double* l, const double* rho, double* betas) {  // [B11 B12 B22 B13 B23]
    double L[30], b5[5];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 5; ++j) L[i * 5 + j] = l[i * 10 + j];
    svd_solve<6, 5>(L, rho, b5);
    if (b5[0] < 0) { betas[0] = sqrt(-b5[0]); betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
    else { betas[0] = sqrt(b5[0]); betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
    betas[3] = 0.0;
}

P2P_HD inline void gauss_newton(const double* l, const double* rho, double* betas) {
    double x[4] = {0, 0, 0, 0};
    for (int it = 0; it < 5; ++it) {
        double A[24], b[6];
        for (int i = 0; i < 6; ++i) {
            const double* r = l + i * 10;
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

// From the 12x12 M^T M: the four eigenvectors of the smallest eigenvalues (rows of ut8) and the three refined
// beta sets (epnp.cpp compute_pose, middle part).
P2P_HD inline void solve_betas(double* mtm, const double cws[4][3], double* ut8, double betas[3][4]) {
    double w[12], l[60], rho[6];
    tridiag_eig_sym<12, 8>(mtm, ut8, w);
    compute_L_6x10(ut8, l);
    compute_rho(cws, rho);
    find_betas_approx_1(l, rho, betas[0]); gauss_newton(l, rho, betas[0]);
    find_betas_approx_2(l, rho, betas[1]); gauss_newton(l, rho, betas[1]);
    find_betas_approx_3(l, rho, betas[2]); gauss_newton(l, rho, betas[2]);
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

// estimate_R_and_t from the centroids and the 3x3 correlation ABt = sum (pc-pc0)(pw-pw0)^T.
P2P_HD inline void rt_from_correlation(const double* abt_in, const double* pc0, const double* pw0, double R[3][3], double* t) {
    double U[9], s[3], V[9];
    for (int i = 0; i < 9; ++i) U[i] = abt_in[i];
    jacobi_svd<3, 3>(U, s, V);
    // complete U to an orthonormal basis when the correlation is rank deficient
    if (!(s[2] > 1e-14 * s[0])) {
        if (!(s[1] > 1e-14 * s[0])) {  // rank <= 1: pick any unit vector orthogonal to u0
            const double ax = fabs(U[0]), ay = fabs(U[3]), az = fabs(U[6]);
            double e[3] = {0, 0, 0};
            e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
            const double d = e[0] * U[0] + e[1] * U[3] + e[2] * U[6];
            double v1[3] = {e[0] - d * U[0], e[1] - d * U[3], e[2] - d * U[6]};
            const double nn = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
            U[1] = v1[0] / nn; U[4] = v1[1] / nn; U[7] = v1[2] / nn;
        }
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
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
P2P_HD inline void rodrigues_to_vec(const double Rin[3][3], double* r) {
    double U[9], s[3], V[9], R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) U[i * 3 + j] = Rin[i][j];
    jacobi_svd<3, 3>(U, s, V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
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
    R[0][0] = c + c1 * x * x;     R[0][1] = c1 * x * y - s * z; R[0][2] = c1 * x * z + s * y;
    R[1][0] = c1 * x * y + s * z; R[1][1] = c + c1 * y * y;     R[1][2] = c1 * y * z - s * x;
    R[2][0] = c1 * x * z - s * y; R[2][1] = c1 * y * z + s * x; R[2][2] = c + c1 * z * z;
}

// Whole EPnP for a small point set held by one thread (the 5-point RANSAC kernel).
// pws: n x 3 object points, us: n x 2 pixel coordinates.  Returns the winning R, t.
template <int MAXN>
P2P_HD inline void solve_small(const double* pws, const double* us, int n, const Cam& cam, double R[3][3], double* t) {
    double cws[4][3], c0[3] = {0, 0, 0}, sc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) c0[j] += pws[3 * i + j];
    for (int j = 0; j < 3; ++j) c0[j] /= n;
    for (int i = 0; i < n; ++i) {
        const double d[3] = {pws[3 * i] - c0[0], pws[3 * i + 1] - c0[1], pws[3 * i + 2] - c0[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sc[a * 3 + b] += d[a] * d[b];
    }
    choose_control_points(c0, sc, n, cws);
    double ci[9], alphas[MAXN * 4];
    control_inverse(cws, ci);
    double mtm[144];
    for (int i = 0; i < 144; ++i) mtm[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        barycentric(ci, cws, pws + 3 * i, alphas + 4 * i);
        double m1[12], m2[12];
        m_rows(alphas + 4 * i, us[2 * i], us[2 * i + 1], cam, m1, m2);
        for (int a = 0; a < 12; ++a)
            for (int b = 0; b < 12; ++b) mtm[a * 12 + b] += m1[a] * m1[b] + m2[a] * m2[b];
    }
    double ut[48], betas[3][4];   // eigenvectors 8..11 only
    solve_betas(mtm, cws, ut, betas);
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
        rt_from_correlation(abt, pc0, pw0, Rk, tk);
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
