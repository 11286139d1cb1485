// Batched EPnP-RANSAC kernels.  See pnp_ransac.cuh for the contract and epnp_core.cuh for the math.
#include "pnp_ransac.cuh"

#include <algorithm>
#include <vector>

#include "epnp_core.cuh"

namespace p2p {

namespace {

constexpr int kHypThreads = 192;     // x 144 doubles of shared-memory workspace per thread = 216 KB: one block per SM
constexpr int kScoreWarps = 8;
constexpr int kRefitThreads = 32;    // one warp per problem: the kernel is dominated by the serial eigen / Gauss-Newton section of ONE thread per
                                     // problem at 255 registers, so what counts is that all problems are resident at once (8 blocks per SM = one
                                     // wave up to 1184 blocks; at 64 threads 768 problems took two waves: 0.99 -> 0.66 ms).  Fixed, not chosen per
                                     // batch: the order of the block sums must not depend on the batch size (batch invariance).
constexpr int kListCap = 1024;       // consensus sets up to this size are handed to the refit as an index list (no scan over all correspondences)

// cv::projectPoints (no distortion) + PnPRansacCallback::computeError for one correspondence:
// projection in double, stored as float32, squared error accumulated in float32 without FMA.
__device__ __forceinline__ float reproj_err_f32(const double* R, const double* t, const float* o, const float* ip,
                                                double fu, double fv, double uc, double vc) {
    const double X = o[0], Y = o[1], Z = o[2];
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z != 0.0 ? 1.0 / z : 1.0;
    x *= z;
    y *= z;
    const float u = static_cast<float>(x * fu + uc), v = static_cast<float>(y * fv + vc);
    const float dx = __fsub_rn(ip[0], u), dy = __fsub_rn(ip[1], v);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// Inlier test for the scoring kernel: the same decision as reproj_err_f32(...) <= thr2, but evaluated in fp32 first;
// only correspondences whose fp32 error lands within `margin` of the threshold (fp32 error bound ~3e-3 px^2 at
// 5 px) or is not finite are re-evaluated through the exact fp64 path.  Pf = float copy of the projection matrix
// K [R|t] (rows 0-1 scaled by fu / fv, so no separate intrinsics multiply), row 2 = the depth.
__device__ __forceinline__ bool is_inlier_fast(const double* R, const double* t, const float* Pf, const float4 pt, float vrel,
                                               const float* __restrict__ ip_exact, const PnpProblem& pr, float thr2, float margin) {
    // pt = (X, Y, Z, u - uc), vrel = v - vc (the centre is subtracted once per tile, not per hypothesis)
    const float x = fmaf(Pf[0], pt.x, fmaf(Pf[1], pt.y, fmaf(Pf[2], pt.z, Pf[3])));
    const float y = fmaf(Pf[4], pt.x, fmaf(Pf[5], pt.y, fmaf(Pf[6], pt.z, Pf[7])));
    const float z = fmaf(Pf[8], pt.x, fmaf(Pf[9], pt.y, fmaf(Pf[10], pt.z, Pf[11])));
    float iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(z));    // one MUFU: the margin below absorbs its error
    const float dx = fmaf(-x, iz, pt.w), dy = fmaf(-y, iz, vrel);
    const float e = fmaf(dx, dx, dy * dy);
    if (fabsf(e - thr2) > margin && fabsf(z) > 1e-3f) return e <= thr2;   // NaN/inf fall through to the exact path
    const float o[3] = {pt.x, pt.y, pt.z};
    return reproj_err_f32(R, t, o, ip_exact, pr.fu, pr.fv, pr.uc, pr.vc) <= thr2;   // image point as stored (global memory: rare)
}

// ---- (1) hypotheses: one 5-point EPnP per thread, flat over (problem, iteration).  Every thread replays OpenCV's
// RNG up to its own iteration (<= 100 x 5 draws: noise next to the solver).  The 12x12 M^T M / Jacobi workspace of
// each thread lives in shared memory, interleaved across the block ([element][thread]: conflict-free, and it is what
// used to spill into 3.4 KB of local memory per thread).
__global__ void __launch_bounds__(kHypThreads, 1) ransac_hyp_kernel(const PnpProblem* __restrict__ probs, int n_problems,
                                                                 const float* __restrict__ obj, const float* __restrict__ img,
                                                                 double* __restrict__ hyp, int iters, int it0, int it1,
                                                                 const int* __restrict__ limit) {
    extern __shared__ double s_ws[];   // [144][kHypThreads]
    const long long gw = static_cast<long long>(blockIdx.x) * kHypThreads + threadIdx.x;   // flat over (problem, iteration of the wave)
    const int p = static_cast<int>(gw / (it1 - it0)), h = it0 + static_cast<int>(gw % (it1 - it0));
    if (p >= n_problems) return;
    if (limit && h >= limit[p]) return;   // OpenCV's loop ends before this iteration: it never draws this subset
    const long long g = static_cast<long long>(p) * iters + h;
    const PnpProblem pr = probs[p];
    if (pr.n < 6) return;
    int idx[5];
    {
        CvRng rng;
        for (int i = 0; i <= h; ++i) ransac_subset5(rng, pr.n, idx);
    }
    const epnp::Cam cam = {pr.fu, pr.fv, pr.uc, pr.vc};
    const double ifx = 1.0 / pr.fu, ify = 1.0 / pr.fv;
    double pws[15], us[10];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const long long q = pr.offset + idx[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) pws[3 * j + k] = static_cast<double>(obj[q * 3 + k]);
        // cv::undistortPoints on CV_32FC2 (no distortion): normalised in double, stored as float32;
        // epnp::init_points then maps back to pixels in double.
        const float xn = static_cast<float>((static_cast<double>(img[q * 2]) - pr.uc) * ifx);
        const float yn = static_cast<float>((static_cast<double>(img[q * 2 + 1]) - pr.vc) * ify);
        us[2 * j] = static_cast<double>(xn) * pr.fu + pr.uc;
        us[2 * j + 1] = static_cast<double>(yn) * pr.fv + pr.vc;
    }
    double R[3][3], t[3], rv[3];
    epnp::solve_small<5, kHypThreads>(pws, us, 5, cam, s_ws + threadIdx.x, R, t);
    epnp::rodrigues_to_vec<kHypThreads>(R, rv, s_ws + threadIdx.x);   // the RANSAC model is (rvec, tvec) ...
    epnp::rodrigues_to_mat(rv, R);   // ... and projectPoints turns rvec back into a matrix
    double* o = hyp + g * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i * 3 + j] = R[i][j];
    o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
}

// ---- (2) scoring: one warp per hypothesis over the correspondences.  A block stages a tile of kScoreTile
// correspondences of its problem in shared memory ONCE and its 8 warps walk all `iters` hypotheses over it
// (warp w takes hypotheses w, w+8, ...); per-tile inlier counts are warp-shuffle-reduced and added to
// counts[problem][hypothesis].  (Reading the points straight from global memory made this kernel L1/L2-bandwidth
// bound: every point was fetched `iters` times.)
constexpr int kScoreTile = 2048;

__global__ void __launch_bounds__(kScoreWarps * 32) ransac_score_kernel(const PnpProblem* __restrict__ probs,
                                                                       const float* __restrict__ obj,
                                                                       const float* __restrict__ img,
                                                                       const double* __restrict__ hyp, int* __restrict__ counts,
                                                                       int iters, int it0, int it1, const int* __restrict__ limit,
                                                                       float thr2) {
    __shared__ float4 s_pt[kScoreTile];   // (X, Y, Z, u - uc)
    __shared__ float s_v[kScoreTile];     // v - vc
    const PnpProblem pr = probs[blockIdx.y];
    const int base = blockIdx.x * kScoreTile;
    if (pr.n < 6 || base >= pr.n) return;
    if (limit) it1 = min(it1, limit[blockIdx.y]);
    if (it0 >= it1) return;
    const int cnt_pts = min(kScoreTile, pr.n - base);
    const float* go = obj + (pr.offset + base) * 3;
    const float* gi = img + (pr.offset + base) * 2;
    const float ucf = static_cast<float>(pr.uc), vcf = static_cast<float>(pr.vc);
    for (int i = threadIdx.x; i < cnt_pts; i += blockDim.x) {
        s_pt[i] = make_float4(go[3 * i], go[3 * i + 1], go[3 * i + 2], gi[2 * i] - ucf);
        s_v[i] = gi[2 * i + 1] - vcf;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float margin = 0.01f * thr2 + 0.05f;   // fp32 evaluation error at 5 px: ~3e-3 px^2
    for (int h = it0 + warp; h < it1; h += kScoreWarps) {
        const double* m = hyp + (static_cast<long long>(blockIdx.y) * iters + h) * 12;
        double R[9], t[3];
        float Pf[12];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = m[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = m[9 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Pf[i] = static_cast<float>(R[i] * pr.fu);
            Pf[4 + i] = static_cast<float>(R[3 + i] * pr.fv);
            Pf[8 + i] = static_cast<float>(R[6 + i]);
        }
        Pf[3] = static_cast<float>(t[0] * pr.fu); Pf[7] = static_cast<float>(t[1] * pr.fv); Pf[11] = static_cast<float>(t[2]);
        int cnt = 0;
        for (int i = lane; i < cnt_pts; i += 32)
            cnt += is_inlier_fast(R, t, Pf, s_pt[i], s_v[i], gi + 2 * i, pr, thr2, margin) ? 1 : 0;  // NaN -> exact path -> false
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0 && cnt) atomicAdd(&counts[blockIdx.y * iters + h], cnt);
    }
}

// ---- (2b) between waves: has OpenCV's loop already ended?  The accept / adaptive-termination rule of
// RANSACPointSetRegistrator::run replayed over the iterations scored so far; limit[p] = `niters` at the wave's end (it only
// ever decreases): iterations at or beyond it are ones OpenCV never runs, so the later waves skip them.
__global__ void ransac_wave_check_kernel(const PnpProblem* __restrict__ probs, int n_problems, const int* __restrict__ counts,
                                         int* __restrict__ limit, int iters, int wave_end, double confidence) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_problems) return;
    const PnpProblem pr = probs[p];
    if (pr.n < 6) { limit[p] = 0; return; }
    int max_good = 0, niters = iters > 1 ? iters : 1;
    for (int it = 0; it < niters && it < wave_end; ++it) {
        const int c = counts[p * iters + it];
        if (c > (max_good > 4 ? max_good : 4)) {
            max_good = c;
            niters = ransac_update_num_iters(confidence, static_cast<double>(pr.n - c) / pr.n, 5, niters);
        }
    }
    limit[p] = niters;
}

// ---- (3) replay of RANSACPointSetRegistrator::run's accept / adaptive-termination rule + inlier mask
__global__ void __launch_bounds__(256) ransac_select_kernel(const PnpProblem* __restrict__ probs, const float* __restrict__ obj,
                                                            const float* __restrict__ img, const double* __restrict__ hyp,
                                                            const int* __restrict__ counts, int* __restrict__ best,
                                                            uint8_t* __restrict__ mask, int* __restrict__ small_list,
                                                            PnpResult* __restrict__ res,
                                                            int iters, float thr2, double confidence) {
    __shared__ int s_best;
    const PnpProblem pr = probs[blockIdx.x];
    PnpResult* out = res + blockIdx.x;
    if (pr.n < 6) {
        if (threadIdx.x == 0) {
            out->status = -1; out->n_inliers = -1; out->best_iter = -1; out->iters_run = 0; out->n_mask = 0;
            best[2 * blockIdx.x] = -1; best[2 * blockIdx.x + 1] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) {
        int max_good = 0, niters = iters > 1 ? iters : 1, b = -1, it = 0;
        for (; it < niters; ++it) {
            const int c = counts[blockIdx.x * iters + it];
            if (c > (max_good > 4 ? max_good : 4)) {  // goodCount > MAX(maxGoodCount, modelPoints-1)
                b = it;
                max_good = c;
                niters = ransac_update_num_iters(confidence, static_cast<double>(pr.n - c) / pr.n, 5, niters);
            }
        }
        s_best = b;
        best[2 * blockIdx.x] = b;
        best[2 * blockIdx.x + 1] = max_good;
        out->best_iter = b;
        out->iters_run = it;
        out->n_inliers = b >= 0 ? max_good : -1;
        out->status = b >= 0 ? 1 : 0;
    }
    __syncthreads();
    const int b = s_best;
    uint8_t* mk = mask + pr.offset;
    if (b < 0) {
        for (int i = threadIdx.x; i < pr.n; i += blockDim.x) mk[i] = 0;
        if (threadIdx.x == 0) out->n_mask = 0;
        return;
    }
    const double* m = hyp + (static_cast<long long>(blockIdx.x) * iters + b) * 12;
    double R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = m[i];
    t[0] = m[9]; t[1] = m[10]; t[2] = m[11];
    const float* o = obj + pr.offset * 3;
    const float* ip = img + pr.offset * 2;
    // Inlier mask; when the consensus set is small (<= kSmallRefit) also its indices in ascending order, for the
    // exact serial refit (ordered compaction: ballot + running block offset, one 256-wide chunk at a time).
    __shared__ int s_wc[8], s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int* list = small_list + static_cast<long long>(blockIdx.x) * kListCap;
    for (int base = 0; base < pr.n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const bool f = i < pr.n && reproj_err_f32(R, t, o + 3 * i, ip + 2 * i, pr.fu, pr.fv, pr.uc, pr.vc) <= thr2;
        if (i < pr.n) mk[i] = f ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if ((threadIdx.x & 31) == 0) s_wc[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int off = s_cnt;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s_wc[w];
        off += __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
        if (f && off < kListCap) list[off] = i;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 0; w < 8; ++w) s_cnt += s_wc[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out->n_mask = s_cnt;
}

// ---- (4a) refit on a small consensus set (5 is the minimum the accept rule allows): M^T M is (nearly) rank deficient
// again and the result depends on every rounding, so one thread per problem runs OpenCV's exact serial operation order
// on the inliers in index order (solve_small, the routine of the hypothesis kernel; workspace in shared memory).
constexpr int kSmallThreads = 16;   // x 144 doubles of workspace = 18 KB of shared memory per block (of either kind): 8 blocks per SM fit
__device__ __forceinline__ void refit_small_block(int block, const PnpProblem* __restrict__ probs, int n_problems,
                                                  const float* __restrict__ obj, const float* __restrict__ img,
                                                  const int* __restrict__ small_list, PnpResult* __restrict__ res, double* s_ws) {
    if (threadIdx.x >= kSmallThreads) return;
    const int p = block * kSmallThreads + threadIdx.x;
    if (p >= n_problems) return;
    const PnpProblem pr = probs[p];
    PnpResult* out = res + p;
    if (pr.n < 6 || out->status != 1 || out->n_mask > kSmallRefit) return;
    const int m = out->n_mask;
    const float* o = obj + pr.offset * 3;
    const float* ip = img + pr.offset * 2;
    const int* list = small_list + static_cast<long long>(p) * kListCap;
    const epnp::Cam cam = {pr.fu, pr.fv, pr.uc, pr.vc};
    const double ifx = 1.0 / pr.fu, ify = 1.0 / pr.fv;
    double pws[kSmallRefit * 3], us[kSmallRefit * 2], Rk[3][3], tk[3], rv[3], Rf[3][3];
    for (int j = 0; j < m; ++j) {
        const int i = list[j];
        for (int k = 0; k < 3; ++k) pws[3 * j + k] = static_cast<double>(o[3 * i + k]);
        // solvePnP on the CV_64F inliers: undistortPoints to normalised coordinates (double), back to pixels in epnp
        us[2 * j] = (static_cast<double>(ip[2 * i]) - pr.uc) * ifx * pr.fu + pr.uc;
        us[2 * j + 1] = (static_cast<double>(ip[2 * i + 1]) - pr.vc) * ify * pr.fv + pr.vc;
    }
    epnp::solve_small<kSmallRefit, kSmallThreads>(pws, us, m, cam, s_ws + threadIdx.x, Rk, tk);
    epnp::rodrigues_to_vec<kSmallThreads>(Rk, rv, s_ws + threadIdx.x);
    epnp::rodrigues_to_mat(rv, Rf);
    for (int i = 0; i < 3; ++i) {
        out->rvec[i] = rv[i];
        out->tvec[i] = tk[i];
        for (int j = 0; j < 3; ++j) out->R[i * 3 + j] = Rf[i][j];
    }
}

// ---- (4b) EPnP on all inliers (double) of a larger consensus set, one CTA per problem
template <int NV>
__device__ __forceinline__ void block_reduce(double* v, double* s_red, double* s_out) {
    // v[NV] per thread -> s_out[NV] block sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[warp * NV + k] = x;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NV; k += blockDim.x) {
        double x = 0;
        for (int w = 0; w < nw; ++w) x += s_red[w * NV + k];
        s_out[k] = x;
    }
    __syncthreads();
}

// Blocks 0 .. n_problems-1: one problem each (block reductions over its inliers).  The blocks after them run the exact serial
// refits of the small consensus sets (4a), 16 problems per block, CONCURRENTLY with the large ones: both are bound by one
// thread's serial latency, so as separate launches they simply added up (0.6 + 1.0 ms per 768 problems).
__global__ void __launch_bounds__(kRefitThreads, 8) epnp_refit_kernel(const PnpProblem* __restrict__ probs, int n_problems,
                                                                   const float* __restrict__ obj, const float* __restrict__ img,
                                                                   const uint8_t* __restrict__ mask, const int* __restrict__ small_list,
                                                                   PnpResult* __restrict__ res) {
    __shared__ double s_small[144 * kSmallThreads];   // workspace of the small-set blocks; the large ones use its first 2 KB
    if (static_cast<int>(blockIdx.x) >= n_problems) {
        refit_small_block(static_cast<int>(blockIdx.x) - n_problems, probs, n_problems, obj, img, small_list, res, s_small);
        return;
    }
    double* s_red = s_small;                          // [(kRefitThreads / 32) * 52]
    double* s_sum = s_small + (kRefitThreads / 32) * 52;   // [52]
    __shared__ double s_cws[4][3], s_ci[9], s_R[3][3][3], s_t[3][3];
    __shared__ int s_first;
    const PnpProblem pr = probs[blockIdx.x];
    PnpResult* out = res + blockIdx.x;
    if (pr.n < 6 || out->status != 1) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < 3; ++i) { out->rvec[i] = 0; out->tvec[i] = 0; }
            for (int i = 0; i < 9; ++i) out->R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        }
        return;
    }
    if (out->n_mask <= kSmallRefit) return;   // refitted by the small-set blocks
    const float* o = obj + pr.offset * 3;
    const float* ip = img + pr.offset * 2;
    const uint8_t* mk = mask + pr.offset;
    const epnp::Cam cam = {pr.fu, pr.fv, pr.uc, pr.vc};
    // The inliers: the ascending index list the selection kernel left (consensus sets up to kListCap points -- a sparse set in a
    // 16 k-point crop costs its own size per pass, not a scan of the mask), else the mask itself.
    const bool use_list = out->n_mask <= kListCap;
    const int* list = small_list + static_cast<long long>(blockIdx.x) * kListCap;
    const int n_walk = use_list ? out->n_mask : pr.n;
    auto inlier_at = [&](int j) -> int { return use_list ? list[j] : (mk[j] ? j : -1); };
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();

    // pass A: centroid, count, index of the first inlier (solve_for_sign looks at point 0)
    {
        double v[4] = {0, 0, 0, 0};
        int first = 0x7fffffff;
        for (int j = threadIdx.x; j < n_walk; j += blockDim.x) {
            const int i = inlier_at(j);
            if (i >= 0) {
                v[0] += o[3 * i]; v[1] += o[3 * i + 1]; v[2] += o[3 * i + 2]; v[3] += 1.0;
                first = min(first, i);
            }
        }
        atomicMin(&s_first, first);
        block_reduce<4>(v, s_red, s_sum);
    }
    const int m = static_cast<int>(s_sum[3] + 0.5);
    const double ifx = 1.0 / pr.fu, ify = 1.0 / pr.fv;
    const double c0[3] = {s_sum[0] / m, s_sum[1] / m, s_sum[2] / m};
    __syncthreads();
    // pass B: scatter matrix PW0^T PW0
    {
        double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = threadIdx.x; j < n_walk; j += blockDim.x) {
            const int i = inlier_at(j);
            if (i >= 0) {
                const double d[3] = {o[3 * i] - c0[0], o[3 * i + 1] - c0[1], o[3 * i + 2] - c0[2]};
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) v[a * 3 + b] += d[a] * d[b];
            }
        }
        block_reduce<9>(v, s_red, s_sum);
    }
    if (threadIdx.x == 0) {
        double sc[9];
        for (int i = 0; i < 9; ++i) sc[i] = s_sum[i];
        double ws[epnp::kWs];
        epnp::choose_control_points<1>(c0, sc, m, s_cws, ws);
        epnp::control_inverse<1>(s_cws, s_ci, ws);
    }
    __syncthreads();
    // pass C: M^T M in its structured form (10 alpha pairs x {1, uc-u, vc-v, (uc-u)^2+(vc-v)^2}) and
    //         T_j = sum alpha_j (pw - c0) for the R|t correlation
    {
        double v[52];
#pragma unroll
        for (int k = 0; k < 52; ++k) v[k] = 0;
        for (int j = threadIdx.x; j < n_walk; j += blockDim.x) {
            const int i = inlier_at(j);
            if (i >= 0) {
                const double p[3] = {o[3 * i], o[3 * i + 1], o[3 * i + 2]};
                double a[4];
                epnp::barycentric(s_ci, s_cws, p, a);
                const double du = pr.uc - ((static_cast<double>(ip[2 * i]) - pr.uc) * ifx * pr.fu + pr.uc);
                const double dv = pr.vc - ((static_cast<double>(ip[2 * i + 1]) - pr.vc) * ify * pr.fv + pr.vc);
                const double q = du * du + dv * dv;
                int k = 0;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = x; y < 4; ++y) {
                        const double aa = a[x] * a[y];
                        v[k] += aa; v[10 + k] += aa * du; v[20 + k] += aa * dv; v[30 + k] += aa * q;
                        ++k;
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[40 + j * 3 + c] += a[j] * (p[c] - c0[c]);
            }
        }
        block_reduce<52>(v, s_red, s_sum);
    }
    __shared__ epnp::RefitShared s_sh;
    if (threadIdx.x == 0) {
        // sign reference: camera-frame z of the first inlier (epnp.cpp solve_for_sign uses pcs[2])
        const int f = s_first;
        const double pf[3] = {o[3 * f], o[3 * f + 1], o[3 * f + 2]};
        epnp::refit_prepare(s_sum, s_cws, s_ci, cam, pf, s_sh);
    }
    __syncthreads();
    if (threadIdx.x < 3)   // the three beta initialisations are independent: one thread each instead of one after the other
        epnp::refit_candidate(static_cast<int>(threadIdx.x), s_sum, m, c0, s_sh, s_R[threadIdx.x], s_t[threadIdx.x]);
    __syncthreads();
    // pass E: mean reprojection distance of each of the three candidates
    {
        double v[3] = {0, 0, 0};
        for (int j = threadIdx.x; j < n_walk; j += blockDim.x) {
            const int i = inlier_at(j);
            if (i >= 0) {
                const double p[3] = {o[3 * i], o[3 * i + 1], o[3 * i + 2]};
                const double u = (static_cast<double>(ip[2 * i]) - pr.uc) * ifx * pr.fu + pr.uc;
                const double vv = (static_cast<double>(ip[2 * i + 1]) - pr.vc) * ify * pr.fv + pr.vc;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] += epnp::reproj_dist(s_R[c], s_t[c], p, u, vv, cam);
            }
        }
        block_reduce<3>(v, s_red, s_sum);
    }
    if (threadIdx.x == 0) {
        int N = 0;
        if (s_sum[1] < s_sum[0]) N = 1;
        if (s_sum[2] < s_sum[N]) N = 2;
        double rv[3], Rf[3][3], ws[epnp::kWs];
        epnp::rodrigues_to_vec<1>(s_R[N], rv, ws);
        epnp::rodrigues_to_mat(rv, Rf);
        for (int i = 0; i < 3; ++i) {
            out->rvec[i] = rv[i];
            out->tvec[i] = s_t[N][i];
            for (int j = 0; j < 3; ++j) out->R[i * 3 + j] = Rf[i][j];
        }
    }
}

}  // namespace

PnpSolver::PnpSolver() {
    require_device();
    P2P_CUDA(cudaFuncSetAttribute(ransac_hyp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(double) * 144 * kHypThreads)));
    P2P_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
}
PnpSolver::~PnpSolver() {
    if (stream_) cudaStreamDestroy(stream_);
}

void PnpSolver::ensure(int n_problems, int iters) {
    if (n_problems > cap_problems_ || iters > cap_iters_) {
        cap_problems_ = std::max(n_problems, cap_problems_);
        cap_iters_ = std::max(iters, cap_iters_);
        hyp_.alloc(static_cast<size_t>(cap_problems_) * cap_iters_ * 12);
        counts_.alloc(static_cast<size_t>(cap_problems_) * cap_iters_);
        best_.alloc(static_cast<size_t>(cap_problems_) * 2);
        small_.alloc(static_cast<size_t>(cap_problems_) * kListCap);
        limit_.alloc(static_cast<size_t>(cap_problems_));
    }
}

void PnpSolver::solve_batch(const PnpProblem* problems_dev, int n_problems, const float* obj_dev, const float* img_dev,
                            uint8_t* mask_dev, PnpResult* results_dev, float reproj_err, int iters, double confidence,
                            cudaStream_t s, int max_n) {
    if (n_problems <= 0) return;
    P2P_CHECK(iters >= 1 && iters <= 4096, "iterationsCount %d outside [1,4096]", iters);
    P2P_CHECK(confidence > 0 && confidence < 1, "confidence must be in (0,1)");
    ensure(n_problems, iters);
    const float thr2 = static_cast<float>(static_cast<double>(reproj_err) * static_cast<double>(reproj_err));
    static const bool prof = getenv("P2P_PROF_PNP") && atoi(getenv("P2P_PROF_PNP")) != 0;       // per-kernel times to stderr
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    auto mark = [&](int i) {
        if (!prof) return;
        P2P_CUDA(cudaEventCreate(&ev[i]));
        P2P_CUDA(cudaEventRecord(ev[i], s));
    };
    mark(0);
    P2P_CHECK(max_n >= 0, "max_n must be the largest correspondence count of the batch");
    P2P_CUDA(cudaMemsetAsync(counts_.p, 0, sizeof(int) * static_cast<size_t>(n_problems) * iters, s));
    // Waves in iteration order: hypotheses [0, kFirstWave) are generated and scored first; a problem whose replayed loop
    // has ended by then (niters <= kFirstWave: a consensus of 67 % inliers or more) skips the rest, as OpenCV does.
    static const int first_wave = getenv("P2P_RANSAC_WAVE") ? atoi(getenv("P2P_RANSAC_WAVE")) : kFirstWave;
    const int w1 = (first_wave > 0 && iters > first_wave + first_wave / 2) ? first_wave : iters;
    const size_t hyp_smem = sizeof(double) * 144 * kHypThreads;
    const dim3 g(std::max(1, (max_n + kScoreTile - 1) / kScoreTile), n_problems);
    for (int it0 = 0; it0 < iters; it0 = (it0 == 0 ? w1 : iters)) {
        const int it1 = it0 == 0 ? w1 : iters;
        const int* limit = it0 == 0 ? nullptr : limit_.p;
        const long long total = static_cast<long long>(n_problems) * (it1 - it0);
        ransac_hyp_kernel<<<static_cast<unsigned>((total + kHypThreads - 1) / kHypThreads), kHypThreads, hyp_smem, s>>>(
            problems_dev, n_problems, obj_dev, img_dev, hyp_.p, iters, it0, it1, limit);
        P2P_CUDA(cudaGetLastError());
        ransac_score_kernel<<<g, kScoreWarps * 32, 0, s>>>(problems_dev, obj_dev, img_dev, hyp_.p, counts_.p, iters, it0, it1, limit, thr2);
        P2P_CUDA(cudaGetLastError());
        launches += 2;
        if (it1 < iters) {
            ransac_wave_check_kernel<<<(n_problems + 127) / 128, 128, 0, s>>>(problems_dev, n_problems, counts_.p, limit_.p, iters, it1, confidence);
            P2P_CUDA(cudaGetLastError());
            ++launches;
            mark(1);
        }
    }
    if (w1 == iters) mark(1);
    mark(2);
    ransac_select_kernel<<<n_problems, 256, 0, s>>>(problems_dev, obj_dev, img_dev, hyp_.p, counts_.p, best_.p, mask_dev,
                                                    small_.p, results_dev, iters, thr2, confidence);
    P2P_CUDA(cudaGetLastError());
    mark(3);
    epnp_refit_kernel<<<n_problems + (n_problems + kSmallThreads - 1) / kSmallThreads, kRefitThreads, 0, s>>>(
        problems_dev, n_problems, obj_dev, img_dev, mask_dev, small_.p, results_dev);
    P2P_CUDA(cudaGetLastError());
    mark(4);
    if (prof) {
        P2P_CUDA(cudaStreamSynchronize(s));
        float ms[4];
        for (int i = 0; i < 4; ++i) P2P_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
        fprintf(stderr, "pnp[%d problems] hyp+score wave 1 %.3f  wave 2 %.3f  select %.3f  refit %.3f ms\n", n_problems, ms[0], ms[1], ms[2], ms[3]);
        for (auto e : ev) cudaEventDestroy(e);
    }
    launches += 2;
}

void PnpSolver::solve_host(const double* obj, const double* img, int n, const double* K9, float reproj_err, int iters,
                           double confidence, PnpResult* out, uint8_t* mask_out) {
    P2P_CHECK(obj && img && K9 && out, "NULL argument");
    P2P_CHECK(n >= 0, "negative point count");
    std::vector<float> o(static_cast<size_t>(n) * 3), ip(static_cast<size_t>(n) * 2);
    for (size_t i = 0; i < o.size(); ++i) o[i] = static_cast<float>(obj[i]);   // Mat::convertTo(CV_32F)
    for (size_t i = 0; i < ip.size(); ++i) ip[i] = static_cast<float>(img[i]);
    PnpProblem pr;
    pr.offset = 0; pr.n = n; pr.pad = 0;
    pr.fu = K9[0]; pr.fv = K9[4]; pr.uc = K9[2]; pr.vc = K9[5];
    h_obj_.upload(o.data(), o.size(), stream_);
    h_img_.upload(ip.data(), ip.size(), stream_);
    if (h_mask_.n < static_cast<size_t>(n) + 1) h_mask_.alloc(static_cast<size_t>(n) + 1);
    h_prob_.upload(&pr, 1, stream_);
    if (h_res_.n < 1) h_res_.alloc(1);
    solve_batch(h_prob_.p, 1, h_obj_.p, h_img_.p, h_mask_.p, h_res_.p, reproj_err, iters, confidence, stream_, n);
    P2P_CUDA(cudaMemcpyAsync(out, h_res_.p, sizeof(PnpResult), cudaMemcpyDeviceToHost, stream_));
    if (mask_out && n > 0) P2P_CUDA(cudaMemcpyAsync(mask_out, h_mask_.p, n, cudaMemcpyDeviceToHost, stream_));
    P2P_CUDA(cudaStreamSynchronize(stream_));
}

void PnpSolver::solve_host_batch(const double* obj, const double* img, const int* counts, int n_problems, const double* K9,
                                 int k_per_problem, float reproj_err, int iters, double confidence, PnpResult* out,
                                 uint8_t* mask_out, float* device_ms) {
    P2P_CHECK(n_problems >= 0, "negative problem count");
    if (device_ms) *device_ms = 0.f;
    if (n_problems == 0) return;
    P2P_CHECK(obj && img && counts && K9 && out, "NULL argument");
    std::vector<PnpProblem> pr(static_cast<size_t>(n_problems));
    long long total = 0;
    int max_n = 0;
    for (int i = 0; i < n_problems; ++i) {
        P2P_CHECK(counts[i] >= 0, "problem %d: negative point count", i);
        const double* K = K9 + (k_per_problem ? 9 * static_cast<size_t>(i) : 0);
        pr[i].offset = total; pr[i].n = counts[i]; pr[i].pad = 0;
        pr[i].fu = K[0]; pr[i].fv = K[4]; pr[i].uc = K[2]; pr[i].vc = K[5];
        total += counts[i];
        max_n = std::max(max_n, counts[i]);
    }
    std::vector<float> o(static_cast<size_t>(total) * 3), ip(static_cast<size_t>(total) * 2);
    for (size_t i = 0; i < o.size(); ++i) o[i] = static_cast<float>(obj[i]);   // Mat::convertTo(CV_32F)
    for (size_t i = 0; i < ip.size(); ++i) ip[i] = static_cast<float>(img[i]);
    h_obj_.upload(o.data(), o.size(), stream_);
    h_img_.upload(ip.data(), ip.size(), stream_);
    if (h_mask_.n < static_cast<size_t>(total) + 1) h_mask_.alloc(static_cast<size_t>(total) + 1);
    h_prob_.upload(pr.data(), pr.size(), stream_);
    if (h_res_.n < static_cast<size_t>(n_problems)) h_res_.alloc(static_cast<size_t>(n_problems));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (device_ms) {
        ensure(n_problems, iters);   // allocate outside the timed region
        P2P_CUDA(cudaEventCreate(&e0));
        P2P_CUDA(cudaEventCreate(&e1));
        P2P_CUDA(cudaEventRecord(e0, stream_));
    }
    solve_batch(h_prob_.p, n_problems, h_obj_.p, h_img_.p, h_mask_.p, h_res_.p, reproj_err, iters, confidence, stream_, max_n);
    if (device_ms) P2P_CUDA(cudaEventRecord(e1, stream_));
    P2P_CUDA(cudaMemcpyAsync(out, h_res_.p, sizeof(PnpResult) * n_problems, cudaMemcpyDeviceToHost, stream_));
    if (mask_out && total > 0) P2P_CUDA(cudaMemcpyAsync(mask_out, h_mask_.p, static_cast<size_t>(total), cudaMemcpyDeviceToHost, stream_));
    P2P_CUDA(cudaStreamSynchronize(stream_));
    if (device_ms) {
        P2P_CUDA(cudaEventElapsedTime(device_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
}

}  // namespace p2p
