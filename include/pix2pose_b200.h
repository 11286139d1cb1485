/* pix2pose_b200 -- C ABI of the B200-native Pix2Pose per-detection inference hot path.
 *
 * The reference (kirumang/Pix2Pose) is pure Python and has no FFI; the boundary it exposes is the
 * class `pix2pose` of pix2pose_model/recognition.py and the Keras `generator_train` duck type
 * (`load_weights`, `predict`).  Each entry point below names the reference call it replaces.
 * pix2pose_b200/{recognition,ae_model}.py bind these through ctypes.
 *
 * Conventions: plain pointers and sizes, caller owns every buffer it passes; handles own their
 * device memory; every function returns 0 (P2P_OK) or a negative status and never throws --
 * `p2p_last_error()` returns the message of the last failure on the calling thread.  There is no
 * CPU fallback: without an sm_100 device, creation fails with P2P_ERR_NO_DEVICE.
 */
#ifndef PIX2POSE_B200_H
#define PIX2POSE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define P2P_API __attribute__((visibility("default")))
#else
#define P2P_API
#endif

#define P2P_OK 0
#define P2P_ERR_INVALID (-1)
#define P2P_ERR_CUDA (-2)
#define P2P_ERR_NO_DEVICE (-3)
#define P2P_ERR_INTERNAL (-4)

#define P2P_PREC_FP16X3 0 /* fp16 hi/lo operand pairs, 3 MMAs per k-step: matches fp32 to ~1e-5 (default) */
#define P2P_PREC_FP16 1   /* plain fp16 operands: ~3x faster contractions, ~1e-2 max abs error      */

typedef struct p2p_engine p2p_engine_t; /* per device: activation workspace + plan for one backbone   */
typedef struct p2p_model p2p_model_t;   /* per object: packed weights (reference keeps one Keras model
                                           per object id, tools/5_evaluation_bop_basic.py:206-225)   */

P2P_API const char* p2p_last_error(void);
P2P_API const char* p2p_version(void);

/* Selects the CUDA device for the calling thread's subsequent p2p_* calls (one process per GPU:
 * pass LOCAL_RANK).  The library links its own CUDA runtime, so torch.cuda.set_device does not reach it. */
P2P_API int p2p_set_device(int device);

/* Number of fp32 values in a weight blob for `backbone` ("resnet50" | "paper"): all tensors of
 * pix2pose_b200/weights.py:param_names(backbone), Keras layouts, concatenated in that order
 * (= the order Model.load_weights walks an inference*.hdf5, recognition.py:23-26). 0 on error. */
P2P_API size_t p2p_param_count(const char* backbone);
/* Algorithmic FLOPs (2*MAC) of one 128x128 forward; 0 on error. */
P2P_API double p2p_flops_per_crop(const char* backbone);

/* ae.aemodel_unet_resnet50(p) / ae.aemodel_unet_prob(p) (ae_model.py:175, :70): builds the graph
 * for batches of up to `capacity` crops on the current CUDA device. */
P2P_API int p2p_engine_create(const char* backbone, int capacity, int precision, p2p_engine_t** out);
P2P_API void p2p_engine_destroy(p2p_engine_t* e);
P2P_API int p2p_engine_capacity(const p2p_engine_t* e);
P2P_API long long p2p_engine_launch_count(const p2p_engine_t* e);

/* generator_train.load_weights(weight_fn) (recognition.py:23,26) after the file has been parsed
 * on the host: `blob` holds p2p_param_count() floats. BN is folded, kernels are repacked. */
P2P_API int p2p_model_create(p2p_engine_t* e, const float* blob, size_t n_floats, p2p_model_t** out);
P2P_API void p2p_model_destroy(p2p_model_t* m);

/* generator_train.predict(x) (recognition.py:84,129): x (n,128,128,3) fp32 NHWC host memory ->
 * decode (n,128,128,3) tanh, prob (n,128,128,1) sigmoid, host memory. Any n >= 0 (chunked). */
P2P_API int p2p_predict(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, float* decode, float* prob);
/* Same with device pointers on `stream` (a cudaStream_t, may be NULL), n <= capacity, asynchronous. */
P2P_API int p2p_predict_device(p2p_engine_t* e, const p2p_model_t* m, const float* x_dev, int n, float* decode_dev,
                       float* prob_dev, void* stream);

/* cv2.solvePnPRansac(obj, img, camK, None, flags=cv2.SOLVEPNP_EPNP, reprojectionError=reproj_err,
 * iterationsCount=iters, confidence=confidence) followed by cv2.Rodrigues(rvec) (recognition.py:216-223).
 * obj_pts (n,3) / img_pts (n,2) double host arrays (converted to float32 as OpenCV does), K row-major 3x3.
 * Outputs: rvec, tvec (EPnP refit on the inliers), R = Rodrigues(rvec), *n_inliers = len(inliers) or -1 when
 * OpenCV would return inliers=None (no hypothesis with more than 4 inliers) or n < 6 (the reference never
 * calls with n < 6, recognition.py:214), inlier_mask (n bytes, may be NULL), *iters_run (may be NULL) =
 * iterations the adaptive RANSAC loop executes. */
P2P_API int p2p_pnp_ransac(const double* obj_pts, const double* img_pts, int n, const double* K, float reproj_err, int iters,
                   double confidence, double* rvec, double* tvec, double* R, int* n_inliers, uint8_t* inlier_mask,
                   int* iters_run);

/* The same for a batch of independent problems in one device run -- the three stage-2 candidates of every detection
 * (recognition.py:196-217) are such a batch.  Problem i owns the next counts[i] rows of the pooled obj_pts / img_pts arrays;
 * K = one row-major 3x3 per problem (k_per_problem != 0) or one for all.  results[i] is what p2p_pnp_ransac returns for
 * problem i (status 1 = pose, 0 = OpenCV would return inliers=None, -1 = fewer than 6 points); inlier_mask (sum of counts
 * bytes, may be NULL); *device_ms (may be NULL) = device time of the solve (CUDA events, copies excluded). */
typedef struct p2p_pnp_result {
    double rvec[3], tvec[3], R[9]; /* R = Rodrigues(rvec) */
    int32_t n_inliers;             /* len(inliers), -1 without a model */
    int32_t best_iter;             /* iteration whose hypothesis won */
    int32_t iters_run;             /* iterations OpenCV's adaptive loop executes */
    int32_t status;
    int32_t n_mask, pad;
} p2p_pnp_result_t;
P2P_API int p2p_pnp_ransac_batch(const double* obj_pts, const double* img_pts, const int* counts, int n_problems,
                         const double* K, int k_per_problem, float reproj_err, int iters, double confidence,
                         p2p_pnp_result_t* results, uint8_t* inlier_mask, float* device_ms);

/* ---- batched est_pose (recognition.py:70-193) ------------------------------------------------
 * One p2p_det_t per est_pose(rgb, bbox) call.  The host fills everything except pool_off / cap_px
 * (pix2pose_b200/recognition.py does the get_boxes arithmetic exactly as the reference). */
typedef struct p2p_det {
    int frame;           /* index into the frame batch                                              */
    int skip;            /* 1: stage-1 size guard (recognition.py:78-79) tripped                    */
    int bbox[4];         /* roi [v0,u0,v1,u1]                                                       */
    int box1[12];        /* get_boxes(bbox, H, W) -> v1_ori,v2_ori,u1_ori,u2_ori,v1,v2,u1,u2,vv1,vv2,uu1,uu2 */
    double fu, fv, uc, vc; /* self.camK                                                             */
    double scale[3], ct[3]; /* obj_param                                                            */
    long long pool_off;  /* internal */
    int cap_px, seg;     /* internal */
    double th_o[8];      /* self.th_o (recognition.py:15) of this detection's object, first n_thresholds entries; read by
                          * p2p_pipeline_run_multi only -- the single-object calls take them as arguments           */
    double th_i;         /* self.th_i (recognition.py:16), likewise                                                  */
} p2p_det_t;

typedef struct p2p_pose {
    double R[9], t[3];   /* rot_pred, tra_pred (model units, mm)                                     */
    double frac_inlier;  /* max_inlier / n_init_mask, or -1                                          */
    int status;          /* 1 ok; 0 the reference returns the -1 sentinels; -2 stage-1 size guard    */
    int n_inliers, best_cand, n_cand;
    int bbox_t[4];       /* [v1,v2,u1,u2] as returned by the reference                               */
    int best_box[12];    /* box of the winning candidate (where img_pred / valid_mask live)          */
    int n_init;          /* n_init_mask                                                              */
    int mask_all_true;   /* valid_mask == -1 case (recognition.py:219)                               */
    int cand_base, pad;
} p2p_pose_t;

typedef struct p2p_pipeline p2p_pipeline_t;
P2P_API int p2p_pipeline_create(p2p_engine_t* e, int max_dets, int n_thresholds, p2p_pipeline_t** out);
P2P_API void p2p_pipeline_destroy(p2p_pipeline_t* p);
/* frames: (F,H,W,3) uint8 host memory (copied to the device inside the call); dets[n]; th_outlier
 * [n_thresholds]; PnP parameters as hard-coded by recognition.py:216-217 are (5, 100, 0.99). */
P2P_API int p2p_pipeline_run(p2p_pipeline_t* p, const p2p_model_t* m, const uint8_t* frames, int F, int H, int W,
                     const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                     int iters, double confidence, p2p_pose_t* out);
/* Same for a float32 image batch (F,H,W,3): tools/5_evaluation_bop_icp3d.py:369 hands est_pose a float32 copy of the
 * frame whose invalid-depth pixels were scaled by 0.1 (non-integer values; numpy evaluates recognition.py:76-77, :114-115
 * in float64 on them).  Crops are cut from the float values, nothing is truncated to uint8. */
P2P_API int p2p_pipeline_run_f32(p2p_pipeline_t* p, const p2p_model_t* m, const float* frames, int F, int H, int W,
                     const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                     int iters, double confidence, p2p_pose_t* out);
/* Same with frames already on the device. */
P2P_API int p2p_pipeline_run_device(p2p_pipeline_t* p, const p2p_model_t* m, const uint8_t* frames_dev, int F, int H, int W,
                            const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                            int iters, double confidence, p2p_pose_t* out);
/* Detections of SEVERAL objects in one device run -- the per-object dispatch of tools/5_evaluation_bop_basic.py:289-304
 * (`obj_pix2pose[model_ids_list.index(obj_id)].est_pose(image, roi)` per ROI, one resident model per object, :206-225):
 * dets[] holds n_seg contiguous groups of seg_counts[s] detections, group s is evaluated with models[s]; obj_param,
 * camK and the thresholds travel in each p2p_det_t.  frames_dev: device pointer from p2p_pipeline_upload_frames
 * (uint8) -- or float32 frames when frames_f32 != 0. */
P2P_API int p2p_pipeline_run_multi(p2p_pipeline_t* p, const p2p_model_t* const* models, const int* seg_counts, int n_seg,
                           const void* frames_dev, int frames_f32, int F, int H, int W, const p2p_det_t* dets, int n,
                           float reproj_err, int iters, double confidence, p2p_pose_t* out);
/* After a run: winner's uint8 XYZ crop (h,w,3) and valid mask (h,w), h = best_box[5]-best_box[4], w = [7]-[6]. */
P2P_API int p2p_pipeline_fetch_crop(p2p_pipeline_t* p, int det, const p2p_pose_t* rec, uint8_t* xyz, uint8_t* mask);
/* After a run: mask-IoU ingredients of tools/5_evaluation_bop_basic.py:307-316 on the device.  masks: (n,H,W) uint8 detector
 * instance masks of the first n detections of the last run (host); out: n x {|mask_pred & m|, |mask_pred | m|} where mask_pred
 * is the full-frame mask est_pose returns (recognition.py:170-178). */
P2P_API int p2p_pipeline_mask_iou(p2p_pipeline_t* p, const uint8_t* masks, int n, int H, int W, long long* out);
/* Raw (128,128,3) network output: stage 1 (index = detection) or stage 2 (index = cand_base + k). */
P2P_API int p2p_pipeline_fetch_decode(p2p_pipeline_t* p, int stage, int index, float* out);
/* Debug / parity: raw float buffers of the last run. what: 1 dec1, 2 dec2, 3 x1, 4 x2 ((128,128,3)), 5 prob1, 6 prob2 ((128,128)). */
P2P_API int p2p_pipeline_fetch_buffer(p2p_pipeline_t* p, int what, int index, float* out);
/* Test hook: the NEXT run replaces the network outputs of `stage` (1|2) by these host arrays (planted-pose parity tests). */
P2P_API int p2p_pipeline_debug_override(p2p_pipeline_t* p, int stage, const float* decode, const float* prob, int n);
/* box_size of pix2pose.__init__ (recognition.py:10, :19; used by get_boxes :33-34 for the stage-1 box and for the
 * refined boxes :110).  Applies to the following runs; default 1.5. */
P2P_API int p2p_pipeline_set_box_size(p2p_pipeline_t* p, double box_size);
/* Asynchronous use (throughput loops): after p2p_pipeline_set_async(p, 1) the p2p_pipeline_run* calls return as soon as the
 * batch is queued on the device (their `out` argument is ignored); p2p_pipeline_wait blocks until the records of that run
 * are on the host and copies them to `out`.  One run in flight per pipeline: alternate between two pipelines to prepare and
 * queue batch k+1 while batch k computes (pix2pose_b200.recognition.pix2pose.est_pose_batch_async). */
P2P_API int p2p_pipeline_set_async(p2p_pipeline_t* p, int on);
P2P_API int p2p_pipeline_wait(p2p_pipeline_t* p, p2p_pose_t* out);
/* Test hook: candidate selection (recognition.py:158-178, :189-193) of the last run repeated with caller-supplied PnP results
 * for its n_cands compact candidates (Rt: 12 doubles each = R row-major, t; n_inliers; status 1 = pose, 0 = inliers None). */
P2P_API int p2p_pipeline_debug_select(p2p_pipeline_t* p, const double* Rt, const int* n_inliers, const int* status, int n_cands,
                              int n, p2p_pose_t* out);
P2P_API long long p2p_pipeline_launch_count(const p2p_pipeline_t* p);
/* Measurement: device milliseconds the generator forwards (recognition.py:84 and :129) of the LAST run took, from CUDA
 * event nodes recorded around them inside the run itself (they are part of the captured graph). */
P2P_API int p2p_pipeline_forward_ms(p2p_pipeline_t* p, double* ms);

/* Measurement helper: uploads x (n <= capacity crops, host), then times `iters` device-resident
 * forwards after `warmup` untimed ones with CUDA events on the engine's stream; *ms_per_iter is
 * the mean.  Inputs stay in HBM: this is the kernel-only number (bench.py `value`). */
P2P_API int p2p_time_forward(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, int warmup, int iters,
                     float* ms_per_iter);
/* Measurement: CUDA events on the engine's stream (the stream every kernel of the engine / pipeline is
 * launched on). slot in [0,8). p2p_engine_event_elapsed synchronises on slot_b. */
P2P_API int p2p_engine_event_record(p2p_engine_t* e, int slot);
P2P_API int p2p_engine_event_elapsed(p2p_engine_t* e, int slot_a, int slot_b, float* ms);
/* Measurement across whole pipeline runs: between _begin and _end every generator launch of this engine is bracketed by
 * CUDA events on its stream (runs are then issued kernel by kernel instead of as a captured graph); _end returns the
 * summed milliseconds of the tcgen05 convolution launches (ms[0]) and of the other generator kernels (ms[1]) and their
 * launch counts.  bench.py derives roofline.achieved from it on the real step, not on an isolated forward. */
P2P_API int p2p_engine_prof_begin(p2p_engine_t* e);
P2P_API int p2p_engine_prof_end(p2p_engine_t* e, double* ms, int* counts);
/* Measurement: one forward of n <= capacity crops with every launch bracketed by CUDA events;
 * ms[0] = summed duration of the tcgen05 conv launches, ms[1] = other kernels, counts[0..1] likewise. */
P2P_API int p2p_engine_profile_forward(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, double* ms, int* counts);
/* Copies frames to one of two rotating pipeline-owned device buffers (for p2p_pipeline_run_device / _run_multi); *dev
 * receives the pointer.  The copy is issued on a copy stream and the call returns at once: the frames of the next batch
 * travel while the previous batch computes, and the run that consumes *dev waits for the copy on the device.  With pinned
 * host memory (p2p_host_alloc) the caller must leave `frames` untouched until that run has returned. */
P2P_API int p2p_pipeline_upload_frames(p2p_pipeline_t* p, const uint8_t* frames, int F, int H, int W, void** dev);
/* Page-locked host buffers for the end-to-end path (bench.py `e2e`). */
P2P_API void* p2p_host_alloc(size_t bytes);
P2P_API void p2p_host_free(void* p);

/* Debug/parity: copy intermediate activation `name` (e.g. "f1", "act3d", "d2_uni") of the last
 * forward as (n,H,W,C) fp32; *h,*w,*c receive the shape. `out` may be NULL to query the shape. */
P2P_API int p2p_engine_read_tensor(p2p_engine_t* e, const char* name, int n, float* out, int* h, int* w, int* c);

/* ---- depth -> ICP inputs (pix2pose_util/common_util.py:13-90; callers tools/5_evaluation_bop_icp3d.py:78-79,
 * ros_kinetic/ros_pix2pose.py:180-181, 296-297).  Host buffers, float64, bbox = [v0,u0,v1,u1] or NULL for the whole image.
 * p2p_depth_xyz      = getXYZ(depth, fx, fy, cx, cy, bbox)                           -> out (h,w,3)
 * p2p_depth_normals  = get_normal(...) after its cv2.inpaint step: Gaussian smoothing with `sigma` (2 in the reference;
 *                      0 = refine=False), np.gradient(.., 2, edge_order=2) on the bbox crop, unit normals -> out (h,w,3) */
P2P_API int p2p_depth_xyz(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double* out);
P2P_API int p2p_depth_normals(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox,
                              double sigma, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PIX2POSE_B200_H */
