// Device pipeline of est_pose.  See pipeline.cuh; line references are into
// pix2pose_model/recognition.py of the reference.
#include "pipeline.cuh"

#include <algorithm>
#include <vector>

namespace p2p {

namespace {

// ------------------------------------------------------------------------------------------------
// resize helpers: skimage.transform.resize(order=1) semantics as pinned in oracle/resize_oracle.py
struct Lerp { int lo, hi; double w; };

__device__ __forceinline__ Lerp axis_map(int o, int n_in, int n_out) {
    const double src = (o + 0.5) * (static_cast<double>(n_in) / n_out) - 0.5;
    const double f = floor(src);
    Lerp l;
    l.lo = static_cast<int>(f);
    l.hi = static_cast<int>(ceil(src));
    l.w = src - f;
    return l;
}
__device__ __forceinline__ int reflect_idx(int i, int dim) {
    if (dim == 1) return 0;
    const int cmax = dim - 1, period = 2 * cmax;
    i = i < 0 ? -i : i;
    i %= period;
    return i > cmax ? period - i : i;
}
__device__ __forceinline__ double bilerp(double p00, double p01, double p10, double p11, double wr, double wc) {
    const double top = (1 - wc) * p00 + wc * p01;
    const double bot = (1 - wc) * p10 + wc * p11;
    return (1 - wr) * top + wr * bot;
}
// clip=True of skimage warp: clip to the input range, keeping pixels that are exactly cval when cval is outside it
__device__ __forceinline__ double clip_keep(double v, double mn, double mx, double cval) {
    const bool keep = !(mn <= cval && cval <= mx);
    if (keep && v == cval) return v;
    return fmin(fmax(v, mn), mx);
}
// float32 L2 norm over 3 channels exactly as np.linalg.norm(x, axis=2) evaluates it for float32 input
__device__ __forceinline__ float norm3_f32(float a, float b, float c) {
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
}

// recognition.py:28-69 in double / int arithmetic (Python float == double, int() truncates)
__device__ void get_boxes_dev(double box_size, const double* bbox, int v_max, int u_max, bool has_ct, int ct_v, int ct_u,
                              double max_w, int* o) {
    if (!has_ct) {
        ct_v = static_cast<int>((bbox[0] + bbox[2]) / 2);
        ct_u = static_cast<int>((bbox[1] + bbox[3]) / 2);
    }
    const double width = bbox[3] - bbox[1], height = bbox[2] - bbox[0];
    const double w = fmin(max_w, fmax(width * box_size, height * box_size));
    const int half = static_cast<int>(w / 2);
    const int v1o = ct_v - half, v2o = ct_v + half, u1o = ct_u - half, u2o = ct_u + half;
    int v1 = v1o, v2 = v2o, u1 = u1o, u2 = u2o, svn = 0, svx = 0, sun = 0, sux = 0;
    if (v1o < 0) { svn = -v1o; v1 = 0; }
    if (v2o > v_max) { svx = -(v2o - v_max); v2 = v_max; }
    if (u1o < 0) { sun = -u1o; u1 = 0; }
    if (u2o > u_max) { sux = -(u2o - u_max); u2 = u_max; }
    o[0] = v1o; o[1] = v2o; o[2] = u1o; o[3] = u2o; o[4] = v1; o[5] = v2; o[6] = u1; o[7] = u2;
    o[8] = svn; o[9] = svx + (v2o - v1o); o[10] = sun; o[11] = sux + (u2o - u1o);
}

// size guards of recognition.py:78 / :116-118 for a box
__device__ __forceinline__ bool box_too_small(const int* b) {
    return (b[1] - b[0]) < 5 || (b[3] - b[2]) < 5 || (b[5] - b[4]) < 5 || (b[7] - b[6]) < 5 || (b[9] - b[8]) <= 0 ||
           (b[11] - b[10]) <= 0;
}

// ------------------------------------------------------------------------------------------------
// P1 / P3: crop + normalise + zero-pad + bilinear resize to 128x128 ('reflect'); P3 also zeroes the
// background given by the resized stage-1 mask (recognition.py:75-82, :103-121)
// PIX = uint8_t (the BOP drivers) or float (tools/5_evaluation_bop_icp3d.py:369 hands est_pose a float32 image whose
// invalid-depth pixels were scaled by 0.1: non-integer values; numpy promotes `pixel - [128,128,128]` to float64 either way)
template <bool STAGE2, typename PIX>
__global__ void __launch_bounds__(256) crop_resize_kernel(const PIX* __restrict__ frames, int H, int W,
                                                          const DetIn* __restrict__ dets, const DetState* __restrict__ state,
                                                          const uint8_t* __restrict__ bits1, int n_th, float* __restrict__ x) {
    int d, k = 0;
    if (STAGE2) { d = blockIdx.x / n_th; k = blockIdx.x % n_th; } else { d = blockIdx.x; }
    const DetIn& det = dets[d];
    if (det.skip) return;
    const int* b1 = det.box1;
    const int* b = b1;
    long long out_idx = d;
    int tbit = 0, mask_all = 0;
    if (STAGE2) {
        const DetState& st = state[d];
        if (k >= st.n_cand) return;
        b = st.own_box[k];
        tbit = st.cand_th[k];
        mask_all = st.mask_all[k];
        out_idx = st.cand_base + k;
    }
    const int p = blockIdx.y * 256 + threadIdx.x;
    const int oy = p >> 7, ox = p & 127;
    const int side_v = b[1] - b[0], side_u = b[3] - b[2];
    const Lerp ly = axis_map(oy, side_v, 128), lx = axis_map(ox, side_u, 128);
    const int rr[2] = {reflect_idx(ly.lo, side_v), reflect_idx(ly.hi, side_v)};
    const int cc[2] = {reflect_idx(lx.lo, side_u), reflect_idx(lx.hi, side_u)};
    const PIX* fr = frames + static_cast<long long>(det.frame) * H * W * 3;
    const int s1v = b1[1] - b1[0], s1u = b1[3] - b1[2];
    double pix[2][2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int r = rr[a], q = cc[c];
            bool inside = r >= b[8] && r < b[9] && q >= b[10] && q < b[11];
            int v = 0, u = 0;
            if (inside) {
                v = b[4] + r - b[8];
                u = b[6] + q - b[10];
                if (STAGE2) {
                    // bg_full (:105-106): everything outside the stage-1 crop, and inside it the pixels whose
                    // up-sampled mask (:103, constant cval=0, > 0.9) is off
                    bool fg = v >= b1[4] && v < b1[5] && u >= b1[6] && u < b1[7];
                    if (fg) {
                        const int mr = v - b1[4] + b1[8], mc = u - b1[6] + b1[10];
                        const Lerp my = axis_map(mr, 128, s1v), mx = axis_map(mc, 128, s1u);
                        const uint8_t* bm = bits1 + static_cast<long long>(d) * 16384;
                        auto tap = [&](int y, int xx) -> double {
                            if (y < 0 || y >= 128 || xx < 0 || xx >= 128) return 0.0;
                            return (bm[y * 128 + xx] >> tbit) & 1 ? 1.0 : 0.0;
                        };
                        double mv = bilerp(tap(my.lo, mx.lo), tap(my.lo, mx.hi), tap(my.hi, mx.lo), tap(my.hi, mx.hi), my.w, mx.w);
                        if (mask_all) mv = clip_keep(mv, 1.0, 1.0, 0.0);
                        fg = mv > 0.9;
                    }
                    inside = fg;
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                pix[a][c][ch] = inside ? (static_cast<double>(fr[(static_cast<long long>(v) * W + u) * 3 + ch]) - 128.0) / 128.0 : 0.0;
        }
    float* o = x + (out_idx * 16384 + p) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        o[ch] = static_cast<float>(bilerp(pix[0][0][ch], pix[0][1][ch], pix[1][0][ch], pix[1][1][ch], ly.w, lx.w));
}

// ------------------------------------------------------------------------------------------------
// P2: stage-1 masks, counts, bbox / centroid reductions, refined boxes (recognition.py:85-111)
__global__ void __launch_bounds__(256) stage1_post_kernel(const DetIn* __restrict__ dets, DetState* __restrict__ state,
                                                          const float* __restrict__ dec1, const float* __restrict__ prob1,
                                                          uint8_t* __restrict__ bits1, int n_th, int H, int W, double box_size) {
    __shared__ int s_cnt[kMaxTh], s_n, s_minv, s_maxv, s_minu, s_maxu;
    __shared__ long long s_sumv, s_sumu;
    const int d = blockIdx.x;
    const DetIn& det = dets[d];
    DetState& st = state[d];
    if (threadIdx.x == 0) {
        for (int t = 0; t < kMaxTh; ++t) s_cnt[t] = 0;
        s_n = 0; s_minv = 1 << 30; s_maxv = -1; s_minu = 1 << 30; s_maxu = -1; s_sumv = 0; s_sumu = 0;
    }
    __syncthreads();
    if (det.skip) {
        if (threadIdx.x == 0) { st.n_init = 0; st.n_cand = 0; }
        return;
    }
    float thf[kMaxTh];
    for (int t = 0; t < n_th; ++t) thf[t] = static_cast<float>(det.th_o[t]);   // prob is float32: numpy compares in float32
    int cnt[kMaxTh] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n = 0, minv = 1 << 30, maxv = -1, minu = 1 << 30, maxu = -1;
    long long sumv = 0, sumu = 0;
    const float* dc = dec1 + static_cast<long long>(d) * 16384 * 3;
    const float* pb = prob1 + static_cast<long long>(d) * 16384;
    for (int p = threadIdx.x; p < 16384; p += blockDim.x) {
        const bool ng = norm3_f32(dc[p * 3], dc[p * 3 + 1], dc[p * 3 + 2]) > 0.3f;  // :89
        uint8_t bits = ng ? 0x80 : 0;
        if (ng) {
            const int v = p >> 7, u = p & 127;
            ++n; minv = min(minv, v); maxv = max(maxv, v); minu = min(minu, u); maxu = max(maxu, u);
            sumv += v; sumu += u;
            const float pr = pb[p];
            for (int t = 0; t < n_th; ++t)
                if (pr < thf[t]) { bits |= 1u << t; ++cnt[t]; }  // :94-95
        }
        bits1[static_cast<long long>(d) * 16384 + p] = bits;
    }
    for (int t = 0; t < n_th; ++t) atomicAdd(&s_cnt[t], cnt[t]);
    atomicAdd(&s_n, n);
    atomicMin(&s_minv, minv); atomicMax(&s_maxv, maxv); atomicMin(&s_minu, minu); atomicMax(&s_maxu, maxu);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_sumv), static_cast<unsigned long long>(sumv));
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_sumu), static_cast<unsigned long long>(sumu));
    __syncthreads();
    if (threadIdx.x != 0) return;
    st.n_init = s_n;
    const int* b1 = det.box1;
    const double cx_o = (det.bbox[3] + det.bbox[1]) / 2.0, cy_o = (det.bbox[2] + det.bbox[0]) / 2.0;  // :72-73
    const int w_stage_1 = b1[1] - b1[0];
    int appended[kMaxTh][12], n_app = 0, nc = 0;
    for (int t = 0; t < n_th; ++t) {
        if (s_cnt[t] < 10) continue;  // :96
        if (s_n == 0) continue;       // :99
        const double sv = (b1[1] - b1[0]) / 128.0, su = (b1[3] - b1[2]) / 128.0;
        const double bb[4] = {s_minv * sv, s_minu * su, s_maxv * sv, s_maxu * su};                        // :101-102
        const int cx_m = static_cast<int>((static_cast<double>(s_sumu) / s_n - (127 / 2.0)) + cx_o);     // :108
        const int cy_m = static_cast<int>((static_cast<double>(s_sumv) / s_n - (127 / 2.0)) + cy_o);     // :109
        int box2[12];
        // :110 -- get_boxes treats ct[0] == -1 as "no centre given" (:29) and falls back to the bbox centre, also for a
        // computed centre row that happens to be -1 (ROIs at the top image edge)
        get_boxes_dev(box_size, bb, H, W, cy_m != -1, cy_m, cx_m, static_cast<double>(w_stage_1), box2);
        for (int i = 0; i < 12; ++i) appended[n_app][i] = box2[i];                                        // :111
        ++n_app;
        if (box_too_small(box2)) continue;                                                                // :116-119
        st.cand_th[nc] = t;
        st.mask_all[nc] = s_cnt[t] == 16384;
        for (int i = 0; i < 12; ++i) st.own_box[nc][i] = box2[i];
        ++nc;
    }
    for (int k = 0; k < nc; ++k)
        for (int i = 0; i < 12; ++i) st.pair_box[k][i] = appended[k][i];  // quirk Q1: k-th output pairs with k-th appended box
    st.n_cand = nc;
}

// Compact candidate numbering within each model segment + live counts per stage-2 forward chunk.
// One block of 256 threads per segment: thread t owns detections t, t + 256, ... of the segment; block-wide exclusive scan of
// n_cand per round (a single thread walking the 900-byte DetState records paid one global-load latency per detection:
// 144 us for 256 detections).  Candidates of segment s occupy slots seg_start[s] * n_th + 0 .. live_s - 1.
__global__ void __launch_bounds__(256) cand_scan_kernel(DetState* __restrict__ state, CandStats* __restrict__ cands,
                                                        const int* __restrict__ seg_tab, int n_seg, int n_th,
                                                        int* __restrict__ n_active, int chunk) {
    __shared__ int s_w[8];
    __shared__ int s_carry;
    const int seg = blockIdx.x;
    const int d_begin = seg_tab[seg], d_end = seg_tab[seg + 1];
    const int slot0 = d_begin * n_th;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int d0 = d_begin; d0 < d_end; d0 += 256) {
        const int d = d0 + threadIdx.x;
        const int nc = d < d_end ? state[d].n_cand : 0;
        int incl = nc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int base = s_carry;
        for (int w = 0; w < warp; ++w) base += s_w[w];
        base += incl - nc;
        if (d < d_end) {
            state[d].cand_base = slot0 + base;
            for (int k = 0; k < nc; ++k) {
                cands[slot0 + base + k].det = d;
                cands[slot0 + base + k].k = k;
            }
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = base + nc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int total = s_carry;
        const int c0 = seg_tab[n_seg + 1 + seg], c1 = seg_tab[n_seg + 1 + seg + 1];
        for (int j = 0; j < c1 - c0; ++j) n_active[c0 + j] = max(0, min(chunk, total - j * chunk));
    }
}

// ------------------------------------------------------------------------------------------------
// P4: stage-2 post-processing per candidate (recognition.py:132-154, :196-213)
constexpr int kPostThreads = 512;

__global__ void __launch_bounds__(kPostThreads, 2) stage2_post_kernel(const DetIn* __restrict__ dets, const DetState* __restrict__ state,
                                                                   CandStats* __restrict__ cands,
                                                                   const float* __restrict__ dec2, const float* __restrict__ prob2,
                                                                   int n_th, uint8_t* __restrict__ xyz_u8,
                                                                   uint8_t* __restrict__ valid, float* __restrict__ obj,
                                                                   float* __restrict__ img, PnpProblem* __restrict__ problems) {
    const int c = blockIdx.x;
    if (cands[c].det < 0) return;   // slot past its segment's live candidates
    __shared__ float s_pmin, s_pmax, s_imin, s_imax;
    __shared__ int s_ng128, s_base, s_nng;
    __shared__ long long s_sv, s_su;
    CandStats& cs = cands[c];
    const int d = cs.det, k = cs.k;
    const DetIn& det = dets[d];
    const int* b = state[d].pair_box[k];
    const double th_i = det.th_i;
    const float* dc = dec2 + static_cast<long long>(c) * 16384 * 3;
    const float* pb = prob2 + static_cast<long long>(c) * 16384;
    if (threadIdx.x == 0) {
        s_pmin = 3.4e38f; s_pmax = -3.4e38f; s_imin = 3.4e38f; s_imax = -3.4e38f; s_ng128 = 0; s_base = 0; s_nng = 0; s_sv = 0; s_su = 0;
    }
    __syncthreads();
    // ---- phase A: ranges needed by the clip of the three resizes
    {
        float pmin = 3.4e38f, pmax = -3.4e38f, imin = 3.4e38f, imax = -3.4e38f;
        int ng = 0;
        for (int p = threadIdx.x; p < 16384; p += blockDim.x) {
            const float x = dc[p * 3], y = dc[p * 3 + 1], z = dc[p * 3 + 2];
            const bool gray = norm3_f32(x, y, z) < 0.3f;  // :137
            ng += gray ? 0 : 1;
            const float pr = pb[p];
            pmin = fminf(pmin, pr); pmax = fmaxf(pmax, pr);
            const float v[3] = {x, y, z};
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float ip = __fdiv_rn(__fadd_rn(gray ? 0.f : v[ch], 1.f), 2.f);  // :139-143
                ip = fminf(fmaxf(ip, 0.f), 1.f);
                imin = fminf(imin, ip); imax = fmaxf(imax, ip);
            }
        }
        // float atomics via int reinterpretation are awkward for negatives: reduce through shuffles instead
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
            pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
            imin = fminf(imin, __shfl_xor_sync(0xffffffffu, imin, o));
            imax = fmaxf(imax, __shfl_xor_sync(0xffffffffu, imax, o));
            ng += __shfl_xor_sync(0xffffffffu, ng, o);
        }
        __shared__ float s_r[4][kPostThreads / 32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { s_r[0][warp] = pmin; s_r[1][warp] = pmax; s_r[2][warp] = imin; s_r[3][warp] = imax; atomicAdd(&s_ng128, ng); }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 0; w < kPostThreads / 32; ++w) {
                s_pmin = fminf(s_pmin, s_r[0][w]); s_pmax = fmaxf(s_pmax, s_r[1][w]);
                s_imin = fminf(s_imin, s_r[2][w]); s_imax = fmaxf(s_imax, s_r[3][w]);
            }
        }
        __syncthreads();
    }
    const double pmin = s_pmin, pmax = s_pmax, imin = s_imin, imax = s_imax;
    const double gmin = s_ng128 == 16384 ? 1.0 : 0.0, gmax = s_ng128 > 0 ? 1.0 : 0.0;
    // ---- phase B: resize back to the crop, quantise, masks, ordered compaction
    const int side_v = b[1] - b[0], side_u = b[3] - b[2];
    const int h = b[5] - b[4], w = b[7] - b[6];
    const long long pool = det.pool_off + static_cast<long long>(k) * det.cap_px;
    const int npx = (h > 0 && w > 0) ? h * w : 0;
    int nng = 0;
    long long sv = 0, su = 0;
    // Two passes per super-chunk of kMaxChunks x 512 pixels instead of three block barriers per 512-pixel chunk:
    //   B1  every thread resizes / quantises its pixels, writes the uint8 map and the valid mask, and each warp leaves
    //       its per-chunk count of valid pixels in shared memory;
    //   --  one block-wide exclusive scan over the chunk totals;
    //   B2  every thread re-reads its own valid flags / uint8 values (its own writes, no fence needed) and writes the
    //       correspondences at offset (pixels before this super-chunk) + (chunks before) + (warps before) + (lanes before):
    //       the row-major order np.where produces (:205-211).
    constexpr int kMaxChunks = 1024, kWarps = kPostThreads / 32;
    __shared__ uint8_t s_cnt[kMaxChunks * kWarps];
    __shared__ int s_coff[kMaxChunks];
    __shared__ int s_wtot[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sbase = 0; sbase < npx; sbase += kMaxChunks * kPostThreads) {
        const int nch = min(kMaxChunks, (npx - sbase + kPostThreads - 1) / kPostThreads);
        // ---- B1
        for (int c = 0; c < nch; ++c) {
            const int q = sbase + c * kPostThreads + threadIdx.x;
            bool is_valid = false;
            int i = 0, j = 0;
            if (q < npx) {
                i = q / w; j = q - i * w;
                const Lerp ly = axis_map(i + b[8], 128, side_v), lx = axis_map(j + b[10], 128, side_u);
                const int ys[2] = {ly.lo, ly.hi}, xs[2] = {lx.lo, lx.hi};
                double tp[2][2], tg[2][2], tx[2][2][3];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int y = ys[a], x = xs[e];
                        if (y < 0 || y >= 128 || x < 0 || x >= 128) {
                            tp[a][e] = 1.0; tg[a][e] = 0.0; tx[a][e][0] = tx[a][e][1] = tx[a][e][2] = 0.5;  // cval 1 / 0 / 0.5
                        } else {
                            const int p = y * 128 + x;
                            const float vx = dc[p * 3], vy = dc[p * 3 + 1], vz = dc[p * 3 + 2];
                            const bool gray = norm3_f32(vx, vy, vz) < 0.3f;
                            tp[a][e] = pb[p];
                            tg[a][e] = gray ? 0.0 : 1.0;
                            const float v3[3] = {vx, vy, vz};
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                float ip = __fdiv_rn(__fadd_rn(gray ? 0.f : v3[ch], 1.f), 2.f);
                                tx[a][e][ch] = fminf(fmaxf(ip, 0.f), 1.f);
                            }
                        }
                    }
                const double prob_ori = clip_keep(bilerp(tp[0][0], tp[0][1], tp[1][0], tp[1][1], ly.w, lx.w), pmin, pmax, 1.0);  // :134
                const bool ng = clip_keep(bilerp(tg[0][0], tg[0][1], tg[1][0], tg[1][1], ly.w, lx.w), gmin, gmax, 0.0) > 0.9;   // :146
                uint8_t u8[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const double v = clip_keep(bilerp(tx[0][0][ch], tx[0][1][ch], tx[1][0][ch], tx[1][1][ch], ly.w, lx.w), imin, imax, 0.5) * 255;  // :144
                    u8[ch] = static_cast<uint8_t>(static_cast<int>(v));  // :154 float -> uint8 truncation
                    xyz_u8[(pool + q) * 3 + ch] = u8[ch];
                }
                if (ng) { ++nng; sv += i + b[4]; su += j + b[6]; }
                is_valid = ng && prob_ori < th_i;  // :203-204
                valid[pool + q] = is_valid ? 1 : 0;
            }
            const unsigned ball = __ballot_sync(0xffffffffu, is_valid);
            if (lane == 0) s_cnt[c * kWarps + warp] = static_cast<uint8_t>(__popc(ball));
        }
        __syncthreads();
        // ---- exclusive scan of the chunk totals (thread t owns chunks 2t, 2t + 1)
        {
            int t0 = 0, t1 = 0;
            const int c0 = 2 * threadIdx.x, c1 = c0 + 1;
            if (c0 < nch)
                for (int wv = 0; wv < kWarps; ++wv) t0 += s_cnt[c0 * kWarps + wv];
            if (c1 < nch)
                for (int wv = 0; wv < kWarps; ++wv) t1 += s_cnt[c1 * kWarps + wv];
            const int pair = t0 + t1;
            int incl = pair;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_wtot[warp] = incl;
            __syncthreads();
            int before = 0;
            for (int wv = 0; wv < warp; ++wv) before += s_wtot[wv];
            const int excl = before + incl - pair;
            if (c0 < nch) s_coff[c0] = excl;
            if (c1 < nch) s_coff[c1] = excl + t0;
            __syncthreads();
        }
        // ---- B2
        for (int c = 0; c < nch; ++c) {
            const int q = sbase + c * kPostThreads + threadIdx.x;
            const bool is_valid = q < npx && valid[pool + q] != 0;
            const unsigned ball = __ballot_sync(0xffffffffu, is_valid);
            if (is_valid) {
                int before = s_base + s_coff[c];
                for (int wv = 0; wv < warp; ++wv) before += s_cnt[c * kWarps + wv];
                const long long dst = pool + before + __popc(ball & ((1u << lane) - 1));
                const int i = q / w, j = q - i * w;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    obj[dst * 3 + ch] = static_cast<float>((static_cast<double>(xyz_u8[(pool + q) * 3 + ch]) / 255 * 2 - 1) * det.scale[ch] + det.ct[ch]);  // :198-202
                img[dst * 2] = static_cast<float>(j + b[6]);      // u + u1  (:209-211)
                img[dst * 2 + 1] = static_cast<float>(i + b[4]);  // v + v1
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int wv = 0; wv < kWarps; ++wv) tot += s_wtot[wv];
            s_base += tot;
        }
        __syncthreads();
    }
    atomicAdd(&s_nng, nng);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_sv), static_cast<unsigned long long>(sv));
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_su), static_cast<unsigned long long>(su));
    __syncthreads();
    if (threadIdx.x == 0) {
        cs.n_non_gray = s_nng;
        cs.skipped = s_nng < 10 ? 1 : 0;  // :149-150
        cs.n_pts = cs.skipped ? 0 : s_base;
        cs.sum_v = static_cast<double>(s_sv);
        cs.sum_u = static_cast<double>(s_su);
        PnpProblem pr;
        pr.offset = pool; pr.n = cs.n_pts; pr.pad = 0;
        pr.fu = det.fu; pr.fv = det.fv; pr.uc = det.uc; pr.vc = det.vc;
        problems[c] = pr;
    }
}

// ------------------------------------------------------------------------------------------------
// P5: candidate selection (recognition.py:158-178, :189-193)
__global__ void select_kernel(const DetIn* __restrict__ dets, const DetState* __restrict__ state, const CandStats* __restrict__ cands,
                              const PnpResult* __restrict__ pnp, PoseRecord* __restrict__ recs, int n_det) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_det) return;
    const DetIn& det = dets[d];
    const DetState& st = state[d];
    PoseRecord r;
    for (int i = 0; i < 9; ++i) r.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    r.t[0] = r.t[1] = r.t[2] = 0;
    r.frac_inlier = -1; r.status = det.skip ? -2 : 0; r.n_inliers = -1; r.best_cand = -1; r.n_cand = st.n_cand;
    r.n_init = st.n_init; r.mask_all_true = 0; r.cand_base = st.cand_base; r.pad = 0;
    for (int i = 0; i < 4; ++i) r.bbox_t[i] = det.box1[4 + i];
    for (int i = 0; i < 12; ++i) r.best_box[i] = det.box1[i];
    int max_inlier = -1;
    double min_dist = 9999999;
    for (int k = 0; k < st.n_cand; ++k) {
        const int c = st.cand_base + k;
        const int* b = st.pair_box[k];
        for (int i = 0; i < 4; ++i) r.bbox_t[i] = b[4 + i];  // quirk Q2: loop variables survive the loop
        const CandStats& cs = cands[c];
        if (cs.skipped) continue;
        const PnpResult& pr = pnp[c];
        const bool ok = pr.status == 1;
        const int n_inl = ok ? pr.n_inliers : -1;
        const double tz = ok ? pr.tvec[2] : 0.0;
        double dist;
        if (tz == 0) {
            dist = 99999;
        } else {
            const double ctv = cs.sum_v / cs.n_non_gray, ctu = cs.sum_u / cs.n_non_gray;
            const double pu = det.fu * pr.tvec[0] / tz + det.uc, pv = det.fv * pr.tvec[1] / tz + det.vc;
            dist = ((pv - ctv) * (pv - ctv) + (pu - ctu) * (pu - ctu)) / (n_inl + 1e-6);
        }
        if (dist < min_dist) {
            min_dist = dist;
            max_inlier = n_inl;
            r.best_cand = k;
            for (int i = 0; i < 9; ++i) r.R[i] = ok ? pr.R[i] : ((i % 4 == 0) ? 1.0 : 0.0);
            for (int i = 0; i < 3; ++i) r.t[i] = ok ? pr.tvec[i] : 0.0;
            for (int i = 0; i < 12; ++i) r.best_box[i] = b[i];
            r.mask_all_true = pr.status == 0 ? 1 : 0;
        }
    }
    r.n_inliers = max_inlier;
    if (max_inlier != -1) {
        r.status = 1;
        r.frac_inlier = static_cast<double>(max_inlier) / st.n_init;
    }
    recs[d] = r;
}

// P6: mask IoU of the estimated pose's mask against the detector's instance mask (tools/5_evaluation_bop_basic.py:307-316):
// mask_pred is the full-frame bool array est_pose returns (zeros, the winner's valid mask pasted at best_box; all-true inside
// the box when PnP returned inliers = None, recognition.py:177, :219).  One CTA per detection counts |A and B| and |A or B|.
__global__ void __launch_bounds__(256) mask_iou_kernel(const DetIn* __restrict__ dets, const PoseRecord* __restrict__ recs,
                                                       const uint8_t* __restrict__ valid, const uint8_t* __restrict__ det_masks,
                                                       int H, int W, long long* __restrict__ out) {
    const int d = blockIdx.x;
    const PoseRecord& r = recs[d];
    const DetIn& di = dets[d];
    const bool has = r.status == 1 && r.best_cand >= 0;
    const int v1 = r.best_box[4], v2 = r.best_box[5], u1 = r.best_box[6], u2 = r.best_box[7];
    const int w = u2 - u1;
    const long long pool = di.pool_off + static_cast<long long>(has ? r.best_cand : 0) * di.cap_px;
    const uint8_t* m = det_masks + static_cast<long long>(d) * H * W;
    int inter = 0, uni = 0;
    for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
        const int y = p / W, x = p - y * W;
        const bool a = m[p] != 0;
        bool b = false;
        if (has && y >= v1 && y < v2 && x >= u1 && x < u2) b = r.mask_all_true || valid[pool + static_cast<long long>(y - v1) * w + (x - u1)] != 0;
        inter += (a && b) ? 1 : 0;
        uni += (a || b) ? 1 : 0;
    }
    __shared__ int s_i[8], s_u[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inter += __shfl_xor_sync(0xffffffffu, inter, o);
        uni += __shfl_xor_sync(0xffffffffu, uni, o);
    }
    if ((threadIdx.x & 31) == 0) { s_i[threadIdx.x >> 5] = inter; s_u[threadIdx.x >> 5] = uni; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long ti = 0, tu = 0;
        for (int k = 0; k < 8; ++k) { ti += s_i[k]; tu += s_u[k]; }
        out[2 * d] = ti;
        out[2 * d + 1] = tu;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
Pipeline::Pipeline(Engine* eng, int max_dets_, int n_th_) : engine(eng), max_dets(max_dets_), n_th(n_th_) {
    P2P_CHECK(n_th >= 1 && n_th <= 7, "number of outlier thresholds must be in [1,7], got %d", n_th);
    P2P_CHECK(max_dets >= 1, "max_dets must be positive");
    const size_t D = max_dets, C = static_cast<size_t>(max_dets) * n_th;
    dets_.alloc(D); state_.alloc(D); cands_.alloc(C); recs_.alloc(D);
    x1_.alloc(D * 16384 * 3); dec1_.alloc(D * 16384 * 3); prob1_.alloc(D * 16384);
    x2_.alloc(C * 16384 * 3); dec2_.alloc(C * 16384 * 3); prob2_.alloc(C * 16384);
    bits1_.alloc(D * 16384);
    problems_.alloc(C); pnp_res_.alloc(C);
    P2P_CUDA(cudaMallocHost(&pinned_dets_, sizeof(DetIn) * D));
    P2P_CUDA(cudaMallocHost(&pinned_recs_, sizeof(PoseRecord) * D));
    if (const char* e = getenv("P2P_GRAPH")) use_graph = atoi(e) != 0;
}
Pipeline::~Pipeline() {
    for (auto& g : graphs_)
        for (auto& e : g.second.exec)
            if (e) cudaGraphExecDestroy(e);
    if (pinned_dets_) cudaFreeHost(pinned_dets_);
    if (pinned_recs_) cudaFreeHost(pinned_recs_);
    for (auto e : fwd_ev_) cudaEventDestroy(e);
    if (done_ev_) cudaEventDestroy(done_ev_);
    if (copy_stream_) {
        cudaStreamDestroy(copy_stream_);
        for (auto& e : frames_ready_) cudaEventDestroy(e);
        for (auto& e : frames_free_) cudaEventDestroy(e);
    }
}

void Pipeline::ensure_pool(long long px) {
    if (px <= pool_px_) return;
    pool_px_ = px + px / 4;
    xyz_u8_.alloc(pool_px_ * 3); valid_.alloc(pool_px_); pnp_mask_.alloc(pool_px_);
    obj_.alloc(pool_px_ * 3); img_.alloc(pool_px_ * 2);
    ++pool_gen_;   // captured graphs hold the old pointers
}

// The kernels of one run, in stream order.  Nothing here depends on values the host does not already have: live counts
// stay on the device (n_active_, CandStats::det), so the sequence can be captured into a CUDA graph.
// phase 0: stage-1 crops | 1: stage-1 forwards | 2: stage-1 post, candidate scan, stage-2 crops | 3: stage-2 forwards |
// 4: stage-2 post, EPnP-RANSAC, selection.  (Five graphs per configuration rather than one, so that run() can put real
// timing events around the generator forwards: events recorded by graph nodes cannot be used with cudaEventElapsedTime.)
void Pipeline::enqueue(int phase, const Model* const* models, const int* seg_counts, int n_seg, const void* frames_dev, bool frames_f32,
                       int H, int W, int n, float reproj_err, int iters, double confidence, int max_cap, cudaStream_t s) {
    const int cap = engine->cap;
    const int C = n * n_th;
    const uint8_t* fr_u8 = static_cast<const uint8_t*>(frames_dev);
    const float* fr_f32 = static_cast<const float*>(frames_dev);
    (void)fr_u8; (void)fr_f32;
    if (phase == 0) {
        if (frames_f32) crop_resize_kernel<false, float><<<dim3(n, 64), 256, 0, s>>>(fr_f32, H, W, dets_.p, nullptr, nullptr, n_th, x1_.p);
        else crop_resize_kernel<false, uint8_t><<<dim3(n, 64), 256, 0, s>>>(fr_u8, H, W, dets_.p, nullptr, nullptr, n_th, x1_.p);
        P2P_CUDA(cudaGetLastError());
        launches += 1;
    } else if (phase == 1) {
        for (int sg = 0, d0 = 0; sg < n_seg; d0 += seg_counts[sg], ++sg)
            for (int b = 0; b < seg_counts[sg]; b += cap) {
                const int nb = std::min(cap, seg_counts[sg] - b);
                const size_t at = static_cast<size_t>(d0 + b);
                engine->forward(*models[sg], x1_.p + at * 16384 * 3, nb, dec1_.p + at * 16384 * 3, prob1_.p + at * 16384, nullptr, s);
            }
    } else if (phase == 2) {
        if (!ov_dec_[0].empty()) {   // parity hook (never inside a captured graph)
            P2P_CUDA(cudaMemcpyAsync(dec1_.p, ov_dec_[0].data(), std::min(ov_dec_[0].size(), dec1_.n) * sizeof(float), cudaMemcpyHostToDevice, s));
            P2P_CUDA(cudaMemcpyAsync(prob1_.p, ov_prob_[0].data(), std::min(ov_prob_[0].size(), prob1_.n) * sizeof(float), cudaMemcpyHostToDevice, s));
            P2P_CUDA(cudaStreamSynchronize(s));
            ov_dec_[0].clear(); ov_prob_[0].clear();
        }
        stage1_post_kernel<<<n, 256, 0, s>>>(dets_.p, state_.p, dec1_.p, prob1_.p, bits1_.p, n_th, H, W, box_size);
        P2P_CUDA(cudaGetLastError());
        P2P_CUDA(cudaMemsetAsync(cands_.p, 0xff, sizeof(CandStats) * C, s));    // det = -1: slot not (yet) a live candidate
        cand_scan_kernel<<<n_seg, 256, 0, s>>>(state_.p, cands_.p, seg_tab_.p, n_seg, n_th, n_active_.p, cap);
        P2P_CUDA(cudaGetLastError());
        if (frames_f32) crop_resize_kernel<true, float><<<dim3(C, 64), 256, 0, s>>>(fr_f32, H, W, dets_.p, state_.p, bits1_.p, n_th, x2_.p);
        else crop_resize_kernel<true, uint8_t><<<dim3(C, 64), 256, 0, s>>>(fr_u8, H, W, dets_.p, state_.p, bits1_.p, n_th, x2_.p);
        P2P_CUDA(cudaGetLastError());
        launches += 3;
    } else if (phase == 3) {
        for (int sg = 0, d0 = 0; sg < n_seg; d0 += seg_counts[sg], ++sg) {
            const int Cs = seg_counts[sg] * n_th;
            const int c0 = host_seg_[n_seg + 1 + sg];
            for (int j = 0; j * cap < Cs; ++j) {
                const int nb = std::min(cap, Cs - j * cap);
                const size_t at = static_cast<size_t>(d0) * n_th + static_cast<size_t>(j) * cap;
                engine->forward(*models[sg], x2_.p + at * 16384 * 3, nb, dec2_.p + at * 16384 * 3, prob2_.p + at * 16384,
                                n_active_.p + c0 + j, s);
            }
        }
    } else {
        if (!ov_dec_[1].empty()) {
            P2P_CUDA(cudaMemcpyAsync(dec2_.p, ov_dec_[1].data(), std::min(ov_dec_[1].size(), dec2_.n) * sizeof(float), cudaMemcpyHostToDevice, s));
            P2P_CUDA(cudaMemcpyAsync(prob2_.p, ov_prob_[1].data(), std::min(ov_prob_[1].size(), prob2_.n) * sizeof(float), cudaMemcpyHostToDevice, s));
            P2P_CUDA(cudaStreamSynchronize(s));
            ov_dec_[1].clear(); ov_prob_[1].clear();
        }
        P2P_CUDA(cudaMemsetAsync(problems_.p, 0, sizeof(PnpProblem) * C, s));  // dead slots: n = 0 -> skipped by the PnP kernels
        stage2_post_kernel<<<C, kPostThreads, 0, s>>>(dets_.p, state_.p, cands_.p, dec2_.p, prob2_.p, n_th, xyz_u8_.p, valid_.p, obj_.p,
                                                      img_.p, problems_.p);
        P2P_CUDA(cudaGetLastError());
        pnp.solve_batch(problems_.p, C, obj_.p, img_.p, pnp_mask_.p, pnp_res_.p, reproj_err, iters, confidence, s, max_cap);
        select_kernel<<<(n + 127) / 128, 128, 0, s>>>(dets_.p, state_.p, cands_.p, pnp_res_.p, recs_.p, n);
        P2P_CUDA(cudaGetLastError());
        launches += 2;
    }
}

void Pipeline::run(const Model* const* models, const int* seg_counts, int n_seg, const void* frames_dev, bool frames_f32, int F, int H,
                   int W, const DetIn* dets, int n, const double* th_o, double th_i, float reproj_err, int iters, double confidence,
                   PoseRecord* out) {
    P2P_CHECK(n >= 0 && n <= max_dets, "run: %d detections, pipeline built for %d", n, max_dets);
    P2P_CHECK(pending_n_ < 0, "run: the previous asynchronous run of this pipeline has not been collected (p2p_pipeline_wait)");
    if (n == 0) {
        if (async_mode) {   // an empty batch still has to be collectable
            if (!done_ev_) P2P_CUDA(cudaEventCreateWithFlags(&done_ev_, cudaEventDisableTiming));
            P2P_CUDA(cudaEventRecord(done_ev_, engine->stream));
            pending_n_ = 0;
        }
        return;
    }
    P2P_CHECK(frames_dev && dets && (out || async_mode) && models && seg_counts && n_seg >= 1, "NULL argument");
    cudaStream_t s = engine->stream;
    host_dets_.assign(dets, dets + n);
    host_seg_.assign(2 * (n_seg + 1), 0);
    {
        int tot = 0, chunks = 0;
        for (int sg = 0; sg < n_seg; ++sg) {
            P2P_CHECK(models[sg] && seg_counts[sg] >= 1, "segment %d: empty or without a model", sg);
            host_seg_[sg] = tot;
            host_seg_[n_seg + 1 + sg] = chunks;
            for (int d = tot; d < tot + seg_counts[sg] && d < n; ++d) host_dets_[d].seg = sg;
            tot += seg_counts[sg];
            chunks += (seg_counts[sg] * n_th + engine->cap - 1) / engine->cap;
        }
        P2P_CHECK(tot == n, "segment counts sum to %d, %d detections given", tot, n);
        host_seg_[n_seg] = tot;
        host_seg_[2 * n_seg + 1] = chunks;
        if (n_active_.n < static_cast<size_t>(chunks + 1)) n_active_.alloc(chunks + 1);
    }
    long long px = 0;
    int max_cap = 0;
    for (int d = 0; d < n; ++d) {
        DetIn& di = host_dets_[d];
        P2P_CHECK(di.frame >= 0 && di.frame < F, "detection %d: frame %d outside [0,%d)", d, di.frame, F);
        const long long side = std::max(0, di.box1[1] - di.box1[0]);
        di.cap_px = static_cast<int>(std::min<long long>(side * side, 1 << 30));
        di.pool_off = px;
        max_cap = std::max(max_cap, di.cap_px);
        px += static_cast<long long>(di.cap_px) * n_th;
        if (th_o) {
            for (int t = 0; t < kMaxTh; ++t) di.th_o[t] = t < n_th ? th_o[t] : 0.0;
            di.th_i = th_i;
        }
    }
    ensure_pool(px + 1);
    pnp.reserve(n * n_th, iters);
    int fslot = -1;
    for (int i = 0; i < 2; ++i)
        if (frames_[i].p && frames_dev == frames_[i].p) fslot = i;
    if (fslot >= 0) P2P_CUDA(cudaStreamWaitEvent(s, frames_ready_[fslot], 0));
    memcpy(pinned_dets_, host_dets_.data(), sizeof(DetIn) * n);
    P2P_CUDA(cudaMemcpyAsync(dets_.p, pinned_dets_, sizeof(DetIn) * n, cudaMemcpyHostToDevice, s));
    seg_tab_.upload(host_seg_.data(), host_seg_.size(), s);

    const bool overrides = !ov_dec_[0].empty() || !ov_dec_[1].empty();
    static const bool prof_pnp = getenv("P2P_PROF_PNP") && atoi(getenv("P2P_PROF_PNP")) != 0;
    const bool graphs = use_graph && !overrides && !prof_pnp && !engine->prof;
    GraphEntry* ge = nullptr;
    if (graphs) {
        // everything a kernel argument or a grid dimension depends on
        std::vector<long long> key = {n, n_seg, H, W, frames_f32 ? 1 : 0, reinterpret_cast<long long>(frames_dev), iters,
                                      static_cast<long long>(reproj_err * 4096.0), static_cast<long long>(confidence * 1e9),
                                      static_cast<long long>(box_size * 1e6), pool_gen_, (max_cap + 2047) / 2048};
        for (int sg = 0; sg < n_seg; ++sg) {
            key.push_back(models[sg]->id);   // not the address: a new model may be allocated where a destroyed one lived
            key.push_back(seg_counts[sg]);
        }
        auto it = graphs_.find(key);
        if (it == graphs_.end() && seen_keys_.insert(key).second) {
            // first run of this configuration: launch kernel by kernel.  Capturing and instantiating five graphs costs about
            // as much as two plain runs' worth of launches, which only pays off when the configuration comes back (throughput
            // loops do; the per-image evaluation loop, where the detection count changes from image to image, mostly does not)
            if (seen_keys_.size() > 4096) seen_keys_.clear();
        } else if (it == graphs_.end()) {
            if (graphs_.size() >= 32) {   // bounded cache: drop everything rather than track recency
                for (auto& g : graphs_)
                    for (auto& e : g.second.exec) cudaGraphExecDestroy(e);
                graphs_.clear();
            }
            GraphEntry fresh;
            const long long l0 = launches + pnp.launches + engine->launches;
            for (int ph = 0; ph < 5; ++ph) {
                cudaGraph_t graph = nullptr;
                P2P_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
                try {
                    enqueue(ph, models, seg_counts, n_seg, frames_dev, frames_f32, H, W, n, reproj_err, iters, confidence, max_cap, s);
                } catch (...) {
                    cudaStreamEndCapture(s, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    for (int q = 0; q < ph; ++q) cudaGraphExecDestroy(fresh.exec[q]);
                    throw;
                }
                P2P_CUDA(cudaStreamEndCapture(s, &graph));
                cudaError_t err = cudaGraphInstantiate(&fresh.exec[ph], graph, 0);
                cudaGraphDestroy(graph);
                if (err != cudaSuccess) {
                    for (int q = 0; q < ph; ++q) cudaGraphExecDestroy(fresh.exec[q]);
                    P2P_CUDA(err);
                }
            }
            fresh.launches = launches + pnp.launches + engine->launches - l0;
            it = graphs_.emplace(key, fresh).first;
        } else {
            launches += it->second.launches;   // the counters only saw these kernels while they were captured
        }
        if (it != graphs_.end()) ge = &it->second;
    }
    if (fwd_ev_.empty()) {
        fwd_ev_.resize(4);
        for (auto& e : fwd_ev_) P2P_CUDA(cudaEventCreate(&e));
    }
    for (int ph = 0; ph < 5; ++ph) {
        // real events around the generator forwards (phases 1 and 3): forward_ms() = bench.py's live roofline numerator
        if (ph == 1) P2P_CUDA(cudaEventRecord(fwd_ev_[0], s));
        if (ph == 3) P2P_CUDA(cudaEventRecord(fwd_ev_[2], s));
        if (ge) P2P_CUDA(cudaGraphLaunch(ge->exec[ph], s));
        else enqueue(ph, models, seg_counts, n_seg, frames_dev, frames_f32, H, W, n, reproj_err, iters, confidence, max_cap, s);
        if (ph == 1) P2P_CUDA(cudaEventRecord(fwd_ev_[1], s));
        if (ph == 3) P2P_CUDA(cudaEventRecord(fwd_ev_[3], s));
    }
    if (fslot >= 0) P2P_CUDA(cudaEventRecord(frames_free_[fslot], s));
    last_n_fwd_ev_ = 4;
    P2P_CUDA(cudaMemcpyAsync(pinned_recs_, recs_.p, sizeof(PoseRecord) * n, cudaMemcpyDeviceToHost, s));
    if (!done_ev_) P2P_CUDA(cudaEventCreateWithFlags(&done_ev_, cudaEventDisableTiming));
    P2P_CUDA(cudaEventRecord(done_ev_, s));
    pending_n_ = n;
    if (!async_mode) wait(out);
}

void Pipeline::wait(PoseRecord* out) {
    // Collects the records of the run launched last (async_mode: run() returns as soon as the work is queued, so the host
    // can prepare and queue the next batch -- on another pipeline instance -- while this one computes).
    P2P_CHECK(pending_n_ >= 0 && done_ev_, "wait: no run in flight");
    P2P_CHECK(out, "NULL argument");
    P2P_CUDA(cudaEventSynchronize(done_ev_));
    memcpy(out, pinned_recs_, sizeof(PoseRecord) * pending_n_);
    pending_n_ = -1;
}

void Pipeline::debug_select(const double* Rt, const int* n_inliers, const int* status, int n_cands, int n, PoseRecord* out) {
    // parity hook: candidate selection (recognition.py:158-178, :189-193) re-run on the LAST run's candidates with PnP
    // results supplied by the caller (e.g. cv2's own per-candidate R|t), everything else as the device computed it
    P2P_CHECK(Rt && n_inliers && status && out, "NULL argument");
    P2P_CHECK(n >= 1 && n <= static_cast<int>(host_dets_.size()) && n_cands >= 0 && n_cands <= n * n_th, "debug_select: bad counts");
    cudaStream_t s = engine->stream;
    std::vector<PnpResult> h(static_cast<size_t>(n) * n_th);
    P2P_CUDA(cudaMemcpyAsync(h.data(), pnp_res_.p, sizeof(PnpResult) * h.size(), cudaMemcpyDeviceToHost, s));
    P2P_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < n_cands; ++c) {
        PnpResult& r = h[c];
        for (int i = 0; i < 9; ++i) r.R[i] = Rt[c * 12 + i];
        for (int i = 0; i < 3; ++i) r.tvec[i] = Rt[c * 12 + 9 + i];
        r.n_inliers = n_inliers[c];
        r.status = status[c];
    }
    DevBuf<PnpResult> tmp(h.size());
    P2P_CUDA(cudaMemcpyAsync(tmp.p, h.data(), sizeof(PnpResult) * h.size(), cudaMemcpyHostToDevice, s));
    select_kernel<<<(n + 127) / 128, 128, 0, s>>>(dets_.p, state_.p, cands_.p, tmp.p, recs_.p, n);
    P2P_CUDA(cudaGetLastError());
    P2P_CUDA(cudaMemcpyAsync(out, recs_.p, sizeof(PoseRecord) * n, cudaMemcpyDeviceToHost, s));
    P2P_CUDA(cudaStreamSynchronize(s));
}

double Pipeline::forward_ms() {
    // events 0-1 bracket the stage-1 forwards, 2-3 the stage-2 forwards of the last run
    P2P_CHECK(fwd_ev_.size() >= 4 && last_n_fwd_ev_ == 4, "forward_ms: no run yet");
    double tot = 0;
    for (int i = 0; i + 1 < 4; i += 2) {
        float ms = 0;
        P2P_CUDA(cudaEventElapsedTime(&ms, fwd_ev_[i], fwd_ev_[i + 1]));
        tot += ms;
    }
    return tot;
}

void Pipeline::fetch_crop(int d, const PoseRecord& rec, uint8_t* xyz_out, uint8_t* mask_out) {
    P2P_CHECK(d >= 0 && d < static_cast<int>(host_dets_.size()), "fetch_crop: bad detection index");
    P2P_CHECK(rec.best_cand >= 0, "fetch_crop: detection has no winning candidate");
    const DetIn& di = host_dets_[d];
    const long long pool = di.pool_off + static_cast<long long>(rec.best_cand) * di.cap_px;
    const int h = rec.best_box[5] - rec.best_box[4], w = rec.best_box[7] - rec.best_box[6];
    const size_t npx = static_cast<size_t>(std::max(h, 0)) * std::max(w, 0);
    if (xyz_out) P2P_CUDA(cudaMemcpy(xyz_out, xyz_u8_.p + pool * 3, npx * 3, cudaMemcpyDeviceToHost));
    if (mask_out) P2P_CUDA(cudaMemcpy(mask_out, valid_.p + pool, npx, cudaMemcpyDeviceToHost));
}

void Pipeline::mask_iou(const uint8_t* masks_host, int n, int H, int W, long long* inter_union_out) {
    P2P_CHECK(masks_host && inter_union_out, "NULL argument");
    P2P_CHECK(n >= 0 && n <= static_cast<int>(host_dets_.size()), "mask_iou: %d masks for %zu detections of the last run", n, host_dets_.size());
    if (n == 0) return;
    cudaStream_t s = engine->stream;
    const size_t bytes = static_cast<size_t>(n) * H * W;
    if (det_masks_.n < bytes) det_masks_.alloc(bytes);
    if (iou_.n < static_cast<size_t>(2 * n)) iou_.alloc(2 * n);
    P2P_CUDA(cudaMemcpyAsync(det_masks_.p, masks_host, bytes, cudaMemcpyHostToDevice, s));
    mask_iou_kernel<<<n, 256, 0, s>>>(dets_.p, recs_.p, valid_.p, det_masks_.p, H, W, iou_.p);
    P2P_CUDA(cudaGetLastError());
    ++launches;
    P2P_CUDA(cudaMemcpyAsync(inter_union_out, iou_.p, sizeof(long long) * 2 * n, cudaMemcpyDeviceToHost, s));
    P2P_CUDA(cudaStreamSynchronize(s));
}

void Pipeline::fetch_decode(int stage, int index, float* out) {
    P2P_CHECK(stage == 1 || stage == 2, "stage must be 1 or 2");
    fetch_buffer(stage, index, out);
}

void Pipeline::fetch_buffer(int what, int index, float* out) {
    P2P_CHECK(what >= 1 && what <= 6 && out, "fetch_buffer: bad selector %d", what);
    const bool s2 = what == 2 || what == 4 || what == 6;
    const size_t lim = s2 ? static_cast<size_t>(max_dets) * n_th : static_cast<size_t>(max_dets);
    P2P_CHECK(index >= 0 && static_cast<size_t>(index) < lim, "fetch_buffer: bad index %d", index);
    const float* base[7] = {nullptr, dec1_.p, dec2_.p, x1_.p, x2_.p, prob1_.p, prob2_.p};
    const size_t cnt = what >= 5 ? 16384 : 16384 * 3;
    P2P_CUDA(cudaMemcpy(out, base[what] + static_cast<size_t>(index) * cnt, cnt * sizeof(float), cudaMemcpyDeviceToHost));
}

void Pipeline::set_override(int stage, const float* dec, const float* prob, int n) {
    P2P_CHECK((stage == 1 || stage == 2) && dec && prob && n >= 0, "set_override: bad argument");
    ov_dec_[stage - 1].assign(dec, dec + static_cast<size_t>(n) * 16384 * 3);
    ov_prob_[stage - 1].assign(prob, prob + static_cast<size_t>(n) * 16384);
}

const void* Pipeline::upload_frames(const void* frames_host, bool f32, int F, int H, int W) {
    // Two rotating device buffers filled on a copy stream: the frames of the next batch can travel while the previous
    // batch is still computing; run() makes the compute stream wait for the copy it consumes.
    const size_t bytes = static_cast<size_t>(F) * H * W * 3 * (f32 ? sizeof(float) : 1);
    if (!copy_stream_) {
        P2P_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        for (auto& e : frames_ready_) P2P_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : frames_free_) P2P_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int slot = frames_next_;
    frames_next_ ^= 1;
    if (frames_[slot].n < bytes) {
        P2P_CUDA(cudaStreamSynchronize(engine->stream));   // a run may still be reading the old allocation
        frames_[slot].alloc(bytes);
        ++pool_gen_;
    }
    P2P_CUDA(cudaStreamWaitEvent(copy_stream_, frames_free_[slot], 0));   // last run that read this buffer has finished
    P2P_CUDA(cudaMemcpyAsync(frames_[slot].p, frames_host, bytes, cudaMemcpyHostToDevice, copy_stream_));
    P2P_CUDA(cudaEventRecord(frames_ready_[slot], copy_stream_));
    return frames_[slot].p;
}

}  // namespace p2p
