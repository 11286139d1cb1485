// Batched PnP-RANSAC (EPnP minimal solver) on the GPU -- replaces the CPU call
//   cv2.solvePnPRansac(obj, img, camK, None, flags=SOLVEPNP_EPNP, reprojectionError=5, iterationsCount=100)
// of pix2pose_model/recognition.py:216-217 followed by cv2.Rodrigues (:223).
//
// Semantics follow OpenCV's solvePnPRansac / RANSACPointSetRegistrator (third-party, not vendored in
// the reference; see epnp_core.cuh): float32 correspondences, 5-point subsets drawn with OpenCV's
// fixed-seed RNG, squared float32 reprojection error <= thr^2, "accept if goodCount > max(best, 4)",
// adaptive iteration count at the given confidence, final EPnP on all inliers in double.
// Hypotheses are evaluated in parallel, in two waves in iteration order ([0, 32), then the rest for the problems whose
// loop has not ended by then); the sequential accept/terminate rule is replayed in iteration order, which selects the
// same hypothesis OpenCV would have stopped at.
//
// Kernels: (1) one thread per hypothesis: 5-point EPnP (fp64, OpenCV's operation order); (2) one warp per hypothesis:
// score all correspondences, warp-shuffle reduction of the inlier count; (2b) between the two waves: replay of the
// termination rule -> per-problem iteration bound; (3) replay + inlier mask + ascending inlier index list;
// (4) one warp per problem: EPnP refit on the inliers (block sums of the structured 12x12 normal matrix, serial
// eigen / Gauss-Newton tail), small consensus sets in OpenCV's exact serial order.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace p2p {

struct PnpProblem {
    long long offset;  // first correspondence of this problem in the pools
    int n;             // number of correspondences (problems with n < 6 are skipped: status -1)
    int pad;
    double fu, fv, uc, vc;
};

struct PnpResult {
    double rvec[3], tvec[3], R[9];  // R = Rodrigues(rvec), as recognition.py:221-223 builds it
    int n_inliers;                  // len(inliers) or -1 when RANSAC found no model / n < 6
    int best_iter;                  // iteration whose hypothesis won
    int iters_run;                  // iterations OpenCV's adaptive loop would have executed
    int status;                     // 1 ok, 0 no consensus, -1 too few points
    int n_mask;                     // inliers counted while writing the mask (== n_inliers when status is 1)
    int pad;
};

constexpr int kFirstWave = 32;      // hypotheses generated and scored before the first termination check
constexpr int kSmallRefit = 32;     // consensus sets up to this size are refitted in OpenCV's exact serial operation order

class PnpSolver {
  public:
    PnpSolver();
    ~PnpSolver();
    // Device-side batch: pools obj (total x 3 f32), img (total x 2 f32), mask (total u8, written).
    void solve_batch(const PnpProblem* problems_dev, int n_problems, const float* obj_dev, const float* img_dev,
                     uint8_t* mask_dev, PnpResult* results_dev, float reproj_err, int iters, double confidence,
                     cudaStream_t s, int max_n);  // max_n >= the largest problem's n (sizes the scoring grid)
    // Host convenience: one problem, double inputs converted to float32 like OpenCV does.
    void solve_host(const double* obj, const double* img, int n, const double* K9, float reproj_err, int iters,
                    double confidence, PnpResult* out, uint8_t* mask_out);
    // Host convenience, batch: problem i owns the next counts[i] correspondences of the pooled arrays; K9 holds one 3x3 per
    // problem (k_per_problem != 0) or one for all.  device_ms (may be NULL): the four-stage solve timed with CUDA events.
    void solve_host_batch(const double* obj, const double* img, const int* counts, int n_problems, const double* K9,
                          int k_per_problem, float reproj_err, int iters, double confidence, PnpResult* out, uint8_t* mask_out,
                          float* device_ms);
    // Allocates the scratch for a batch up front (so that solve_batch can run inside a stream capture).
    void reserve(int n_problems, int iters) { ensure(n_problems, iters); }
    long long launches = 0;

  private:
    void ensure(int n_problems, int iters);
    DevBuf<double> hyp_;   // [problems][iters][12]
    DevBuf<int> counts_;   // [problems][iters]
    DevBuf<int> best_;     // [problems][2]
    DevBuf<int> small_;    // [problems][kSmallRefit] ascending inlier indices of small consensus sets
    DevBuf<int> limit_;    // [problems] iteration bound (`niters`) of the replayed loop after the first wave of hypotheses
    int cap_problems_ = 0, cap_iters_ = 0;
    cudaStream_t stream_ = nullptr;
    DevBuf<float> h_obj_, h_img_;
    DevBuf<uint8_t> h_mask_;
    DevBuf<PnpProblem> h_prob_;
    DevBuf<PnpResult> h_res_;
};

}  // namespace p2p
