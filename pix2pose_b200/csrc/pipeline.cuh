// Batched device pipeline of pix2pose.est_pose (pix2pose_model/recognition.py:70-193): everything
// between the two network forwards and after the second one runs on the GPU with no host
// round-trip: crop / normalise / bilinear resize (:75-82, :113-121), stage-1 masks, counts, bbox and
// centroid reductions and the refined boxes (:85-111), stage-2 resize-back, uint8 quantisation, mask
// and the row-major correspondence compaction (:132-154, :196-213), then PnP-RANSAC
// (pnp_ransac.cuh) and the candidate selection (:158-178, :189-193).
#pragma once
#include <stdint.h>

#include <map>
#include <set>
#include <vector>

#include "common.cuh"
#include "engine.cuh"
#include "pnp_ransac.cuh"

namespace p2p {

constexpr int kMaxTh = 8;  // outlier thresholds per detection (reference configs use 1, 3 or 4)

// One detection = one est_pose(rgb, bbox) call.  Filled by the host.
struct DetIn {
    int frame;        // index into the frame batch
    int skip;         // 1: stage-1 size guard tripped (recognition.py:78-79) -> status -2
    int bbox[4];      // roi [v0,u0,v1,u1] as passed to est_pose
    int box1[12];     // get_boxes(bbox, H, W) (recognition.py:71)
    double fu, fv, uc, vc;      // self.camK at call time
    double scale[3], ct[3];     // obj_param (recognition.py:17-18)
    long long pool_off;         // first pixel slot of this detection's candidate pools
    int cap_px;                 // pixels reserved per candidate (side1^2)
    int seg;                    // index of the model (object) segment this detection belongs to (internal)
    double th_o[kMaxTh];        // self.th_o of the detection's object (cfg_tless_paper.json gives every object its own)
    double th_i;                // self.th_i
};

// Stage-1 statistics and the stage-2 candidates of one detection (device-written).
struct DetState {
    int n_init;                 // np.sum(non_gray), recognition.py:90
    int n_cand;                 // len(input_refined)
    int cand_th[kMaxTh];        // threshold index whose crop feeds candidate k
    int own_box[kMaxTh][12];    // refined box the crop of candidate k was cut with
    int pair_box[kMaxTh][12];   // box_refined[k]: the box stage 2 pairs with network output k (quirk Q1)
    int mask_all[kMaxTh];       // 1 when every one of the 128x128 pixels passed (resize clip corner case)
    int cand_base;              // first compact candidate index of this detection
    int pad;
};

struct CandStats {              // per compact candidate, written by stage-2 post
    int det, k;
    int n_non_gray;             // np.sum(non_gray) after resize-back (:148)
    int skipped;                // 1: n_non_gray < 10 (:149-150)
    int n_pts;                  // correspondences (valid_mask pixels)
    int pad;
    double sum_v, sum_u;        // frame coordinates of the non_gray pixels (centroid, :159-162)
};

struct PoseRecord {             // est_pose result of one detection
    double R[9], t[3];
    double frac_inlier;         // max_inlier / n_init_mask, or -1
    int status;                 // 1 ok; 0 no valid pose (reference returns -1 sentinels); -2 size guard
    int n_inliers;
    int best_cand;              // k of the winning candidate, -1 if none
    int n_cand;
    int bbox_t[4];              // [v1,v2,u1,u2] the reference returns (quirk Q2)
    int best_box[12];           // paired box of the winner (where img_pred_f / valid_mask live)
    int n_init;
    int mask_all_true;          // 1: PnP returned inliers=None -> valid_mask = -1 -> all-true mask (:219, :177)
    int cand_base;              // compact index of this detection's first candidate (for fetch_decode)
    int pad;
};

class Pipeline {
  public:
    Pipeline(Engine* engine, int max_dets, int n_th);
    ~Pipeline();
    // frames_dev: (F,H,W,3) uint8 (or float32 when frames_f32) on the device; dets: host array of n (<= max_dets).
    // Results are copied to `out` (host).  th_o has n_th entries.
    // Detections of several objects in one run: `dets` holds n_seg contiguous segments of seg_counts[s] detections, segment s
    // is evaluated with models[s] (only the generator forwards are per segment; crops, post-processing, PnP and selection
    // run over the whole batch).  th_o == nullptr: thresholds are taken from each DetIn; otherwise th_o / th_i apply to all.
    void run(const Model* const* models, const int* seg_counts, int n_seg, const void* frames_dev, bool frames_f32, int F, int H,
             int W, const DetIn* dets, int n, const double* th_o, double th_i, float reproj_err, int iters, double confidence,
             PoseRecord* out);
    // async_mode: run() returns once the work is queued (`out` unused); wait() blocks until that run's records are on the host.
    bool async_mode = false;
    void wait(PoseRecord* out);
    // After run(): copy the winner's uint8 XYZ crop (h,w,3) and valid mask (h,w) of detection d
    // (h = v2-v1, w = u2-u1 of best_box) into host buffers sized for cap_px pixels.
    void fetch_crop(int d, const PoseRecord& rec, uint8_t* xyz_out, uint8_t* mask_out);
    // After run(): |mask_pred and m_d| and |mask_pred or m_d| for the first n detections against detector masks (n,H,W) uint8
    // (host), mask_pred being the full-frame mask est_pose returns; out = n x {intersection, union}.
    void mask_iou(const uint8_t* masks_host, int n, int H, int W, long long* inter_union_out);
    // Raw network output (128,128,3) of stage 1 (index = detection) or stage 2 (index = compact candidate).
    void fetch_decode(int stage, int index, float* out);
    // Debug / parity: raw float buffers. what: 1 dec1, 2 dec2, 3 x1, 4 x2 ((128,128,3) each), 5 prob1, 6 prob2 ((128,128)).
    void fetch_buffer(int what, int index, float* out);
    // Test hook: replace the network outputs of `stage` (1|2) in the NEXT run by these host arrays
    // (n crops of decode (128,128,3) and prob (128,128)); used by planted-pose parity tests only.
    void set_override(int stage, const float* dec, const float* prob, int n);
    // Test hook: re-run the candidate selection of the last run with caller-supplied PnP results (n_cands compact candidates:
    // R row-major 9 + t 3 per candidate, inlier count, status 1 / 0) -> records of the first n detections.
    void debug_select(const double* Rt, const int* n_inliers, const int* status, int n_cands, int n, PoseRecord* out);
    // Host frames (F,H,W,3) uint8 / float32 -> device copy owned by the pipeline.
    const void* upload_frames(const void* frames_host, bool f32, int F, int H, int W);
    // Device milliseconds the generator forwards of the last run took (stage 1 + stage 2; event nodes inside the run).
    double forward_ms();
    long long launches = 0;
    double box_size = 1.5;   // recognition.py:19 (refined boxes, :110)
    Engine* engine;
    PnpSolver pnp;
    int max_dets, n_th;

  private:
    void ensure_pool(long long px);
    DevBuf<DetIn> dets_;
    DevBuf<DetState> state_;
    DevBuf<CandStats> cands_;
    DevBuf<PoseRecord> recs_;
    DevBuf<float> x1_, dec1_, prob1_, x2_, dec2_, prob2_;
    DevBuf<uint8_t> bits1_;       // (D,128,128): bit t = non_gray & prob<th_t ; bit 7 = non_gray
    DevBuf<int> n_active_;        // live candidates per (segment, stage-2 chunk)
    DevBuf<int> seg_tab_;         // [n_seg + 1] first detection of each segment, then [n_seg + 1] first n_active_ slot
    std::vector<int> host_seg_;
    DevBuf<uint8_t> xyz_u8_, valid_, pnp_mask_;
    DevBuf<float> obj_, img_;
    DevBuf<PnpProblem> problems_;
    DevBuf<PnpResult> pnp_res_;
    DevBuf<uint8_t> frames_[2], det_masks_;
    int frames_next_ = 0;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t frames_ready_[2] = {nullptr, nullptr}, frames_free_[2] = {nullptr, nullptr};
    DevBuf<long long> iou_;
    std::vector<float> ov_dec_[2], ov_prob_[2];
    long long pool_px_ = 0;
    std::vector<DetIn> host_dets_;
    // The kernel sequence of a run has no host-dependent control flow (live counts stay on the device), so it is captured
    // into CUDA graphs (one per phase, see enqueue) per configuration and replayed: ~150 launches cost five graph launches.
    struct GraphEntry { cudaGraphExec_t exec[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; long long launches = 0; };
    std::map<std::vector<long long>, GraphEntry> graphs_;
    std::set<std::vector<long long>> seen_keys_;   // configurations run once so far (captured on their second appearance)
    long long pool_gen_ = 0;
    std::vector<cudaEvent_t> fwd_ev_;
    int n_fwd_ev_ = 0, last_n_fwd_ev_ = 0;
    cudaEvent_t done_ev_ = nullptr;
    int pending_n_ = -1;
    DetIn* pinned_dets_ = nullptr;
    PoseRecord* pinned_recs_ = nullptr;
    void enqueue(int phase, const Model* const* models, const int* seg_counts, int n_seg, const void* frames_dev, bool frames_f32, int H,
                 int W, int n, float reproj_err, int iters, double confidence, int max_cap, cudaStream_t s);

  public:
    bool use_graph = true;        // P2P_GRAPH=0 launches kernel by kernel
};

}  // namespace p2p
