// Host build of the __host__ __device__ EPnP / RANSAC helpers (pix2pose_b200/csrc/epnp_core.cuh)
// so their arithmetic can be unit-tested against cv2 without a GPU.  Test infrastructure only.
#include "../../pix2pose_b200/csrc/epnp_core.cuh"

using namespace p2p;

extern "C" {

void host_epnp_small(const double* pws, const double* us, int n, const double* cam4, double* R9, double* t3) {
    epnp::Cam cam = {cam4[0], cam4[1], cam4[2], cam4[3]};
    double R[3][3];
    if (n <= 8) epnp::solve_small<8>(pws, us, n, cam, R, t3);
    else epnp::solve_small<512>(pws, us, n, cam, R, t3);
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
    epnp::rodrigues_to_vec(R, r3);
    epnp::rodrigues_to_mat(r3, Ro);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R9out[i * 3 + j] = Ro[i][j];
}

void host_eig12(const double* A, double* Vt, double* w) {
    double B[144];
    for (int i = 0; i < 144; ++i) B[i] = A[i];
    epnp::jacobi_eig_sym<12>(B, Vt, w);
}
}
