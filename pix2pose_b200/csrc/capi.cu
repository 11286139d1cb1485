// C ABI (include/pix2pose_b200.h) over the engine.  No exception crosses this boundary.
#include "../../include/pix2pose_b200.h"

#include "engine.cuh"
#include "pnp_ransac.cuh"
#include "pipeline.cuh"

using namespace p2p;

namespace p2p {  // csrc/depth.cu
void depth_xyz(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double* out);
void depth_normals(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double sigma, double* out);
}

struct p2p_engine { std::unique_ptr<Engine> e; };
struct p2p_model { std::unique_ptr<Model> m; };
struct p2p_pipeline { std::unique_ptr<Pipeline> p; };
static_assert(sizeof(p2p_det_t) == sizeof(DetIn), "p2p_det_t must mirror DetIn");
static_assert(sizeof(p2p_pose_t) == sizeof(PoseRecord), "p2p_pose_t must mirror PoseRecord");

extern "C" {

const char* p2p_last_error(void) { return get_last_error(); }
const char* p2p_version(void) { return "pix2pose_b200 0.1 (sm_100a)"; }

int p2p_set_device(int device) {
    return guarded([&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw Error(P2P_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
        P2P_CHECK(device >= 0 && device < count, "device %d outside [0,%d)", device, count);
        P2P_CUDA(cudaSetDevice(device));
    });
}

size_t p2p_param_count(const char* backbone) {
    size_t n = 0;
    guarded([&] { n = param_count(parse_backbone(backbone)); });
    return n;
}

double p2p_flops_per_crop(const char* backbone) {
    double f = 0;
    guarded([&] { f = build_plan(parse_backbone(backbone)).flops_per_crop(); });
    return f;
}

int p2p_engine_create(const char* backbone, int capacity, int precision, p2p_engine_t** out) {
    return guarded([&] {
        P2P_CHECK(out != nullptr, "out is NULL");
        *out = nullptr;
        std::unique_ptr<p2p_engine> h(new p2p_engine);
        h->e.reset(new Engine(parse_backbone(backbone), capacity, precision));
        *out = h.release();
    });
}
void p2p_engine_destroy(p2p_engine_t* e) { delete e; }
int p2p_engine_capacity(const p2p_engine_t* e) { return e ? e->e->cap : 0; }
long long p2p_engine_launch_count(const p2p_engine_t* e) { return e ? e->e->launches : 0; }

int p2p_model_create(p2p_engine_t* e, const float* blob, size_t n_floats, p2p_model_t** out) {
    return guarded([&] {
        P2P_CHECK(e && blob && out, "NULL argument");
        *out = nullptr;
        std::unique_ptr<p2p_model> h(new p2p_model);
        h->m.reset(new Model(e->e.get(), blob, n_floats));
        *out = h.release();
    });
}
void p2p_model_destroy(p2p_model_t* m) { delete m; }

int p2p_predict(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, float* decode, float* prob) {
    return guarded([&] {
        P2P_CHECK(e && m && (n == 0 || (x && decode && prob)), "NULL argument");
        e->e->predict_host(*m->m, x, n, decode, prob);
    });
}

int p2p_predict_device(p2p_engine_t* e, const p2p_model_t* m, const float* x_dev, int n, float* decode_dev,
                       float* prob_dev, void* stream) {
    return guarded([&] {
        P2P_CHECK(e && m && x_dev && decode_dev && prob_dev, "NULL argument");
        e->e->forward(*m->m, x_dev, n, decode_dev, prob_dev, nullptr, static_cast<cudaStream_t>(stream));
    });
}

static PnpSolver& default_pnp() {
    static std::unique_ptr<PnpSolver> s;
    if (!s) s.reset(new PnpSolver());
    return *s;
}

int p2p_pnp_ransac(const double* obj_pts, const double* img_pts, int n, const double* K, float reproj_err, int iters,
                   double confidence, double* rvec, double* tvec, double* R, int* n_inliers, uint8_t* inlier_mask,
                   int* iters_run) {
    return guarded([&] {
        P2P_CHECK(obj_pts && img_pts && K && rvec && tvec && R && n_inliers, "NULL argument");
        PnpResult r;
        default_pnp().solve_host(obj_pts, img_pts, n, K, reproj_err, iters, confidence, &r, inlier_mask);
        for (int i = 0; i < 3; ++i) { rvec[i] = r.rvec[i]; tvec[i] = r.tvec[i]; }
        for (int i = 0; i < 9; ++i) R[i] = r.R[i];
        *n_inliers = r.n_inliers;
        if (iters_run) *iters_run = r.iters_run;
    });
}

static_assert(sizeof(p2p_pnp_result_t) == sizeof(PnpResult), "p2p_pnp_result_t mirrors PnpResult");

int p2p_pnp_ransac_batch(const double* obj_pts, const double* img_pts, const int* counts, int n_problems, const double* K,
                         int k_per_problem, float reproj_err, int iters, double confidence, p2p_pnp_result_t* results,
                         uint8_t* inlier_mask, float* device_ms) {
    return guarded([&] {
        default_pnp().solve_host_batch(obj_pts, img_pts, counts, n_problems, K, k_per_problem, reproj_err, iters, confidence,
                                       reinterpret_cast<PnpResult*>(results), inlier_mask, device_ms);
    });
}

int p2p_pipeline_create(p2p_engine_t* e, int max_dets, int n_thresholds, p2p_pipeline_t** out) {
    return guarded([&] {
        P2P_CHECK(e && out, "NULL argument");
        *out = nullptr;
        std::unique_ptr<p2p_pipeline> h(new p2p_pipeline);
        h->p.reset(new Pipeline(e->e.get(), max_dets, n_thresholds));
        *out = h.release();
    });
}
void p2p_pipeline_destroy(p2p_pipeline_t* p) { delete p; }

int p2p_pipeline_run_device(p2p_pipeline_t* p, const p2p_model_t* m, const uint8_t* frames_dev, int F, int H, int W,
                            const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                            int iters, double confidence, p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p && m, "NULL argument");
        P2P_CHECK(th_outlier, "NULL argument");
        const Model* mm = m->m.get();
        p->p->run(&mm, &n, 1, frames_dev, false, F, H, W, reinterpret_cast<const DetIn*>(dets), n, th_outlier, th_inlier, reproj_err,
                  iters, confidence, reinterpret_cast<PoseRecord*>(out));
    });
}

int p2p_pipeline_run_f32(p2p_pipeline_t* p, const p2p_model_t* m, const float* frames, int F, int H, int W,
                         const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                         int iters, double confidence, p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p && m && frames, "NULL argument");
        P2P_CHECK(F >= 1 && H >= 1 && W >= 1, "bad frame batch shape");
        P2P_CHECK(th_outlier, "NULL argument");
        const void* fd = p->p->upload_frames(frames, true, F, H, W);
        const Model* mm = m->m.get();
        p->p->run(&mm, &n, 1, fd, true, F, H, W, reinterpret_cast<const DetIn*>(dets), n, th_outlier, th_inlier, reproj_err, iters,
                  confidence, reinterpret_cast<PoseRecord*>(out));
    });
}

int p2p_pipeline_run(p2p_pipeline_t* p, const p2p_model_t* m, const uint8_t* frames, int F, int H, int W,
                     const p2p_det_t* dets, int n, const double* th_outlier, double th_inlier, float reproj_err,
                     int iters, double confidence, p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p && m && frames, "NULL argument");
        P2P_CHECK(F >= 1 && H >= 1 && W >= 1, "bad frame batch shape");
        P2P_CHECK(th_outlier, "NULL argument");
        const void* fd = p->p->upload_frames(frames, false, F, H, W);
        const Model* mm = m->m.get();
        p->p->run(&mm, &n, 1, fd, false, F, H, W, reinterpret_cast<const DetIn*>(dets), n, th_outlier, th_inlier, reproj_err, iters,
                  confidence, reinterpret_cast<PoseRecord*>(out));
    });
}

int p2p_pipeline_run_multi(p2p_pipeline_t* p, const p2p_model_t* const* models, const int* seg_counts, int n_seg,
                           const void* frames_dev, int frames_f32, int F, int H, int W, const p2p_det_t* dets, int n,
                           float reproj_err, int iters, double confidence, p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p && models && seg_counts && n_seg >= 1 && n_seg <= 4096, "bad argument");
        std::vector<const Model*> mm(n_seg);
        for (int i = 0; i < n_seg; ++i) {
            P2P_CHECK(models[i], "segment %d has no model", i);
            mm[i] = models[i]->m.get();
        }
        p->p->run(mm.data(), seg_counts, n_seg, frames_dev, frames_f32 != 0, F, H, W, reinterpret_cast<const DetIn*>(dets), n, nullptr,
                  0.0, reproj_err, iters, confidence, reinterpret_cast<PoseRecord*>(out));
    });
}

int p2p_pipeline_fetch_crop(p2p_pipeline_t* p, int det, const p2p_pose_t* rec, uint8_t* xyz, uint8_t* mask) {
    return guarded([&] {
        P2P_CHECK(p && rec, "NULL argument");
        p->p->fetch_crop(det, *reinterpret_cast<const PoseRecord*>(rec), xyz, mask);
    });
}
int p2p_pipeline_fetch_decode(p2p_pipeline_t* p, int stage, int index, float* out) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->fetch_decode(stage, index, out);
    });
}
int p2p_pipeline_fetch_buffer(p2p_pipeline_t* p, int what, int index, float* out) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->fetch_buffer(what, index, out);
    });
}
int p2p_pipeline_debug_override(p2p_pipeline_t* p, int stage, const float* decode, const float* prob, int n) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->set_override(stage, decode, prob, n);
    });
}
int p2p_depth_xyz(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double* out) {
    return guarded([&] { depth_xyz(depth, H, W, fx, fy, cx, cy, bbox, out); });
}
int p2p_depth_normals(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double sigma,
                      double* out) {
    return guarded([&] { depth_normals(depth, H, W, fx, fy, cx, cy, bbox, sigma, out); });
}
int p2p_pipeline_mask_iou(p2p_pipeline_t* p, const uint8_t* masks, int n, int H, int W, long long* out) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->mask_iou(masks, n, H, W, out);
    });
}
int p2p_pipeline_set_box_size(p2p_pipeline_t* p, double box_size) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        P2P_CHECK(box_size > 0.0 && box_size < 100.0, "box_size %g outside (0, 100)", box_size);
        p->p->box_size = box_size;
    });
}
int p2p_pipeline_set_async(p2p_pipeline_t* p, int on) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->async_mode = on != 0;
    });
}
int p2p_pipeline_wait(p2p_pipeline_t* p, p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->wait(reinterpret_cast<PoseRecord*>(out));
    });
}
int p2p_pipeline_debug_select(p2p_pipeline_t* p, const double* Rt, const int* n_inliers, const int* status, int n_cands, int n,
                              p2p_pose_t* out) {
    return guarded([&] {
        P2P_CHECK(p, "NULL argument");
        p->p->debug_select(Rt, n_inliers, status, n_cands, n, reinterpret_cast<PoseRecord*>(out));
    });
}
int p2p_pipeline_forward_ms(p2p_pipeline_t* p, double* ms) {
    return guarded([&] {
        P2P_CHECK(p && ms, "NULL argument");
        *ms = p->p->forward_ms();
    });
}
long long p2p_pipeline_launch_count(const p2p_pipeline_t* p) { return p ? p->p->launches + p->p->pnp.launches : 0; }

int p2p_time_forward(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, int warmup, int iters,
                     float* ms_per_iter) {
    return guarded([&] {
        P2P_CHECK(e && m && x && ms_per_iter && iters > 0 && warmup >= 0, "bad argument");
        Engine& E = *e->e;
        P2P_CHECK(n >= 1 && n <= E.cap, "n=%d outside [1,%d]", n, E.cap);
        P2P_CUDA(cudaMemcpyAsync(E.x.p, x, static_cast<size_t>(n) * 128 * 128 * 3 * sizeof(float), cudaMemcpyHostToDevice, E.stream));
        for (int i = 0; i < warmup; ++i) E.forward(*m->m, E.x.p, n, E.dec.p, E.prob.p, nullptr, E.stream);
        cudaEvent_t a, b;
        P2P_CUDA(cudaEventCreate(&a));
        P2P_CUDA(cudaEventCreate(&b));
        P2P_CUDA(cudaStreamSynchronize(E.stream));
        P2P_CUDA(cudaEventRecord(a, E.stream));
        for (int i = 0; i < iters; ++i) E.forward(*m->m, E.x.p, n, E.dec.p, E.prob.p, nullptr, E.stream);
        P2P_CUDA(cudaEventRecord(b, E.stream));
        P2P_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        P2P_CUDA(cudaEventElapsedTime(&ms, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *ms_per_iter = ms / iters;
    });
}

int p2p_engine_event_record(p2p_engine_t* e, int slot) {
    return guarded([&] {
        P2P_CHECK(e && slot >= 0 && slot < 8, "bad argument");
        Engine& E = *e->e;
        if (!E.ev[slot]) P2P_CUDA(cudaEventCreate(&E.ev[slot]));
        P2P_CUDA(cudaEventRecord(E.ev[slot], E.stream));
    });
}
int p2p_engine_event_elapsed(p2p_engine_t* e, int slot_a, int slot_b, float* ms) {
    return guarded([&] {
        P2P_CHECK(e && ms && slot_a >= 0 && slot_a < 8 && slot_b >= 0 && slot_b < 8, "bad argument");
        Engine& E = *e->e;
        P2P_CHECK(E.ev[slot_a] && E.ev[slot_b], "event slot not recorded");
        P2P_CUDA(cudaEventSynchronize(E.ev[slot_b]));
        P2P_CUDA(cudaEventElapsedTime(ms, E.ev[slot_a], E.ev[slot_b]));
    });
}
int p2p_engine_profile_forward(p2p_engine_t* e, const p2p_model_t* m, const float* x, int n, double* ms, int* counts) {
    return guarded([&] {
        P2P_CHECK(e && m && x && ms && counts, "NULL argument");
        Engine& E = *e->e;
        P2P_CHECK(n >= 1 && n <= E.cap, "n=%d outside [1,%d]", n, E.cap);
        P2P_CUDA(cudaMemcpyAsync(E.x.p, x, static_cast<size_t>(n) * 128 * 128 * 3 * sizeof(float), cudaMemcpyHostToDevice, E.stream));
        for (int i = 0; i < 2; ++i) E.forward(*m->m, E.x.p, n, E.dec.p, E.prob.p, nullptr, E.stream);
        double prof[4] = {0, 0, 0, 0};
        E.prof = prof;
        try {
            E.forward(*m->m, E.x.p, n, E.dec.p, E.prob.p, nullptr, E.stream);
        } catch (...) {
            E.prof = nullptr;
            throw;
        }
        E.prof = nullptr;
        ms[0] = prof[0]; ms[1] = prof[1];
        counts[0] = static_cast<int>(prof[2]); counts[1] = static_cast<int>(prof[3]);
    });
}
int p2p_engine_prof_begin(p2p_engine_t* e) {
    return guarded([&] {
        P2P_CHECK(e, "NULL argument");
        Engine& E = *e->e;
        for (double& v : E.prof_acc) v = 0;
        E.prof = E.prof_acc;
    });
}
int p2p_engine_prof_end(p2p_engine_t* e, double* ms, int* counts) {
    return guarded([&] {
        P2P_CHECK(e && ms && counts, "NULL argument");
        Engine& E = *e->e;
        E.prof = nullptr;
        ms[0] = E.prof_acc[0]; ms[1] = E.prof_acc[1];
        counts[0] = static_cast<int>(E.prof_acc[2]); counts[1] = static_cast<int>(E.prof_acc[3]);
    });
}
int p2p_pipeline_upload_frames(p2p_pipeline_t* p, const uint8_t* frames, int F, int H, int W, void** dev) {
    return guarded([&] {
        P2P_CHECK(p && frames && dev, "NULL argument");
        *dev = const_cast<void*>(p->p->upload_frames(frames, false, F, H, W));   // asynchronous (copy stream); the run waits for it
    });
}

void* p2p_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void p2p_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int p2p_engine_read_tensor(p2p_engine_t* e, const char* name, int n, float* out, int* h, int* w, int* c) {
    return guarded([&] {
        P2P_CHECK(e && name, "NULL argument");
        const int id = e->e->plan.tensor_id(name);
        const TensorSpec& t = e->e->plan.tensors[id];
        if (h) *h = t.H;
        if (w) *w = t.W;
        if (c) *c = t.C;
        if (out) e->e->read_tensor(name, n, out);
    });
}

}  // extern "C"
