// Host build of the __host__ __device__ EPnP / RANSAC helpers (pix2pose_b200/csrc/epnp_core.cuh)
// so their arithmetic can be unit-tested against cv2 without a GPU.  Test infrastructure only.
#include "../../pix2pose_b200/csrc/epnp_core.cuh"

using namespace p2p;

extern "C" {

void host_epnp_small(const double* pws, const double* us, int n, const double* cam4, double* R9, double* t3) {
    epnp::Cam cam = {cam4[0], cam4[1], cam4[2], cam4[3]};
    double R[3][3], mtm[144];
    if (n <= 8) epnp::solve_small<8>(pws, us, n, cam, mtm, R, t3);
    else epnp::solve_small<512>(pws, us, n, cam, mtm, R, t3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R[i][j];
}

void host_ransac_subsets(int count, int iters, int* idx) {
    CvRng rng;
    for (int i = 0; i < iters; ++i) ransac_subset5(rng, count, idx + 5 * i);
}

int host_update_iters(double p, double ep, int mp, int maxit) { return ransac_update_num_iters(p, ep, mp, maxit); }

void host_rodrigues_roundtrip(const double* R9, double* r3, double* R9out) {
    double R[3][3], Ro[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = R9[i * 3 + j];
    double ws[epnp::kWs];
    epnp::rodrigues_to_vec<1>(R, r3, ws);
    epnp::rodrigues_to_mat(r3, Ro);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R9out[i * 3 + j] = Ro[i][j];
}

void host_eig12(const double* A, double* Vt, double* w) {
    double B[144];
    for (int i = 0; i < 144; ++i) B[i] = A[i];
    epnp::tridiag_eig_sym<12>(B, Vt, w);
}

// ---- the bit-exact OpenCV linear-algebra restatements, one export per cv2 function they are checked against
void host_cv_svd12(const double* A, double* w, double* ut, double* vt) {   // cv2.SVDecomp(A) of a 12x12
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) ut[i * 12 + j] = A[j * 12 + i];
    epnp::cv_jacobi_svd<1>(ut, 12, 12, w, vt, 12);
}
void host_cv_svd3(const double* A, double* w, double* ut, double* vt) {
    double ws[epnp::kWs];
    epnp::cv_svd3<1>(A, ws);
    for (int i = 0; i < 9; ++i) { ut[i] = ws[i]; vt[i] = ws[9 + i]; }
    for (int i = 0; i < 3; ++i) w[i] = ws[18 + i];
}
void host_cv_invert3(const double* A, double* x) {                                                   // cv2.invert(A, DECOMP_SVD)
    double ws[epnp::kWs];
    epnp::cv_invert3_svd<1>(A, x, ws);
}
void host_cv_solve6(const double* A, const double* b, int n, double* x) {                           // cv2.solve(A, b, DECOMP_SVD)
    double l[60] = {0}, ws[epnp::kWs];
    const int cols[5] = {0, 1, 2, 3, 4};
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < n; ++j) l[i * 10 + j] = A[i * n + j];
    epnp::cv_solve6_svd<1>(l, cols, n, b, x, ws);
}
// cv2.mulTransposed(M, aTa=True) of the 2n x 12 matrix epnp::fill_M builds from (alphas, us)
void host_mtm(const double* alphas, const double* us, int n, const double* cam4, double* mtm) {
    epnp::Cam cam = {cam4[0], cam4[1], cam4[2], cam4[3]};
    epnp::mtm_exact(alphas, us, n, cam, mtm);
}
// Host emulation of epnp_refit_kernel's large-n path: the 52 structured sums (taken sequentially here, by a block
// reduction on the device), then the same serial tail (epnp::refit_candidates) and candidate choice.
void host_refit_large(const double* pws, const double* us, int n, const double* cam4, double* rvec, double* t3) {
    epnp::Cam cam = {cam4[0], cam4[1], cam4[2], cam4[3]};
    double c0[3] = {0, 0, 0}, sc[9] = {0}, cws[4][3], ci[9], s[52] = {0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) c0[j] += pws[3 * i + j];
    for (int j = 0; j < 3; ++j) c0[j] /= n;
    for (int i = 0; i < n; ++i) {
        const double d[3] = {pws[3 * i] - c0[0], pws[3 * i + 1] - c0[1], pws[3 * i + 2] - c0[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sc[a * 3 + b] += d[a] * d[b];
    }
    double ws[epnp::kWs];
    epnp::choose_control_points<1>(c0, sc, n, cws, ws);
    epnp::control_inverse<1>(cws, ci, ws);
    for (int i = 0; i < n; ++i) {
        double a[4];
        epnp::barycentric(ci, cws, pws + 3 * i, a);
        const double du = cam.uc - us[2 * i], dv = cam.vc - us[2 * i + 1], q = du * du + dv * dv;
        int k = 0;
        for (int x = 0; x < 4; ++x)
            for (int y = x; y < 4; ++y) {
                const double aa = a[x] * a[y];
                s[k] += aa; s[10 + k] += aa * du; s[20 + k] += aa * dv; s[30 + k] += aa * q;
                ++k;
            }
        for (int j = 0; j < 4; ++j)
            for (int c = 0; c < 3; ++c) s[40 + j * 3 + c] += a[j] * (pws[3 * i + c] - c0[c]);
    }
    double Rs[3][3][3], ts[3][3], err[3] = {0, 0, 0};
    epnp::RefitShared sh;
    epnp::refit_prepare(s, cws, ci, cam, pws, sh);
    for (int c = 0; c < 3; ++c) epnp::refit_candidate(c, s, n, c0, sh, Rs[c], ts[c]);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) err[c] += epnp::reproj_dist(Rs[c], ts[c], pws + 3 * i, us[2 * i], us[2 * i + 1], cam);
    int N = 0;
    if (err[1] < err[0]) N = 1;
    if (err[2] < err[N]) N = 2;
    epnp::rodrigues_to_vec<1>(Rs[N], rvec, ws);
    for (int i = 0; i < 3; ++i) t3[i] = ts[N][i];
}
}

// ---- CPU emulation of csrc/pnp_ransac.cu built from the same __host__ __device__ functions
// (test infrastructure: lets the RANSAC semantics be checked against cv2.solvePnPRansac without a GPU).
#include <vector>
static float reproj_err_host(const double* R, const double* t, const float* o, const float* ip, const double* cam) {
    const double X = o[0], Y = o[1], Z = o[2];
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z != 0.0 ? 1.0 / z : 1.0;
    x *= z; y *= z;
    const float u = (float)(x * cam[0] + cam[2]), v = (float)(y * cam[1] + cam[3]);
    volatile float dx = ip[0] - u, dy = ip[1] - v;
    volatile float a = dx * dx, b = dy * dy;
    return a + b;
}

extern "C" int host_ransac(const double* obj64, const double* img64, int n, const double* cam4, float thr, int iters,
                           double conf, double* rvec, double* tvec, unsigned char* mask, int* iters_run) {
    std::vector<float> o(n * 3), ip(n * 2);
    for (int i = 0; i < n * 3; ++i) o[i] = (float)obj64[i];
    for (int i = 0; i < n * 2; ++i) ip[i] = (float)img64[i];
    epnp::Cam cam = {cam4[0], cam4[1], cam4[2], cam4[3]};
    const float thr2 = (float)((double)thr * (double)thr);
    CvRng rng;
    std::vector<double> hyp(iters * 12);
    std::vector<int> counts(iters);
    for (int h = 0; h < iters; ++h) {
        int idx[5];
        ransac_subset5(rng, n, idx);
        double pws[15], us[10];
        for (int j = 0; j < 5; ++j) {
            for (int k = 0; k < 3; ++k) pws[3 * j + k] = o[idx[j] * 3 + k];
            const float xn = (float)(((double)ip[idx[j] * 2] - cam.uc) * (1.0 / cam.fu));
            const float yn = (float)(((double)ip[idx[j] * 2 + 1] - cam.vc) * (1.0 / cam.fv));
            us[2 * j] = (double)xn * cam.fu + cam.uc;
            us[2 * j + 1] = (double)yn * cam.fv + cam.vc;
        }
        double R[3][3], t[3], rv[3], mtm[144];
        epnp::solve_small<5>(pws, us, 5, cam, mtm, R, t);
        epnp::rodrigues_to_vec<1>(R, rv, mtm);
        epnp::rodrigues_to_mat(rv, R);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) hyp[h * 12 + i * 3 + j] = R[i][j];
        for (int i = 0; i < 3; ++i) hyp[h * 12 + 9 + i] = t[i];
        int c = 0;
        for (int i = 0; i < n; ++i) c += reproj_err_host(&hyp[h * 12], &hyp[h * 12 + 9], &o[3 * i], &ip[2 * i], cam4) <= thr2;
        counts[h] = c;
    }
    int max_good = 0, niters = iters > 1 ? iters : 1, best = -1, it = 0;
    for (; it < niters; ++it) {
        if (counts[it] > (max_good > 4 ? max_good : 4)) {
            best = it; max_good = counts[it];
            niters = ransac_update_num_iters(conf, (double)(n - max_good) / n, 5, niters);
        }
    }
    *iters_run = it;
    if (best < 0) return -1;
    std::vector<double> pw, uv;
    for (int i = 0; i < n; ++i) {
        mask[i] = reproj_err_host(&hyp[best * 12], &hyp[best * 12 + 9], &o[3 * i], &ip[2 * i], cam4) <= thr2;
        if (mask[i]) {
            for (int k = 0; k < 3; ++k) pw.push_back(o[3 * i + k]);
            // solvePnP on CV_64F points: undistortPoints to normalised coordinates in double, epnp maps back to pixels
            uv.push_back(((double)ip[2 * i] - cam.uc) * (1.0 / cam.fu) * cam.fu + cam.uc);
            uv.push_back(((double)ip[2 * i + 1] - cam.vc) * (1.0 / cam.fv) * cam.fv + cam.vc);
        }
    }
    double R[3][3], mtm[144];
    epnp::solve_small<20000>(pw.data(), uv.data(), (int)(pw.size() / 3), cam, mtm, R, tvec);   // n-sized scratch on the stack
    epnp::rodrigues_to_vec<1>(R, rvec, mtm);
    return max_good;
}
