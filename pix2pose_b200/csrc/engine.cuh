// Generator execution engine: topology ("plan"), activation workspace, TMA tensor maps,
// weight packing, forward.  Host side of csrc/conv_tc.cuh.
//
// Reference graph: pix2pose_model/ae_model.py:175-240 (resnet50) and :70-150 (paper);
// resnet50_mod.py:40-118, 200-213.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"

namespace p2p {

enum Backbone : int { BB_RESNET50 = 0, BB_PAPER = 1 };
enum Precision : int { PREC_FP16X3 = 0, PREC_FP16 = 1 };

enum LayerKind : int { L_CONV = 0, L_CONVT = 1, L_DENSE = 2, L_BN = 3 };
struct LayerDef {
    std::string name;
    int kind;
    int shape[4];  // Keras layout; BN: shape[0] = C
    size_t count() const;   // floats of all tensors of the layer (kernel + bias, or 4*C)
};
std::vector<LayerDef> layer_table(int backbone);
size_t param_count(int backbone);
int parse_backbone(const char* s);

enum ConvKind : int { K_CONV = 0, K_CONV_S2 = 1, K_CONVT = 2, K_DENSE = 3, K_PATCH = 4, K_CONVT_FUSED = 5, K_STEM = 6 };

struct SrcSpec { int tensor, c_begin, c_count; int view = 0; };  // view 1: every second pixel of the tensor (stride-2 1x1 source)
struct WPart { std::string layer, bn; };

struct KWeight { int kh, kw, cin_begin, nvalid; int part = -1; };  // where a k-iteration's 64 K-rows come from (part >= 0: K-concatenated conv)

struct ConvSpec {
    std::string name;
    int kind = K_CONV, ksize = 1;
    std::vector<SrcSpec> srcs;
    std::vector<WPart> parts;
    int out_tensor = -1, c_off = 0, act = ACT_NONE, res_tensor = -1;
    // derived by Plan::finalize
    int H = 0, W = 0, tw = 0, th = 0, nb = 0, phases = 1;
    int Cin = 0, Cout = 0, Cout_pad = 0, BN = 0;
    int kstart[5] = {0, 0, 0, 0, 0};
    int splitk = 1, splitk_chunk = 0;  // K slices (Dense layers whose M x N grid cannot fill the GPU)
    int oy_off[4] = {0, 0, 0, 0}, ox_off[4] = {0, 0, 0, 0}, sy = 1, sx = 1;
    std::vector<int4> kit;
    std::vector<KWeight> kw;
    struct MapReq { int tensor, view, climit; };  // view: 0 normal, 1..4 parity (py*2+px+1), 5 dense
    std::vector<MapReq> maps;
    bool kcat = false;                 // parts are concatenated along K (one part per source; BN scales folded into the weights):
                                       // out = act(sum_p BN_p(conv_p(src_p))) -- the ResNet shortcut conv fused into branch2c
    bool slab = false;                 // runs on conv_tc_slab_kernel (k x k stride-1 conv, Cout 128, W % 8 == 0, H % 16 == 0; or the heads)
    int slab_ks = 0;                   // tap grid of the slab formulation (= ksize; 3 for the phase-fused heads: 3x3 input neighbourhood)
    std::vector<int4> slabs;           // slab kernel: one entry per (source, 64-channel chunk)
};

struct TensorSpec { std::string name; int H, W, C; };

enum StepKind : int { S_IM2COL = 0, S_CONV = 1, S_MAXPOOL = 2, S_STEMCVT = 3 };
struct Step { int kind; int a, b, c, d; };  // CONV: a = conv index; IM2COL: a=out tensor,b=ks,c=pad ; MAXPOOL: a=in,b=out ; STEMCVT: a=out tensor,b=left pad (pixels)

struct Plan {
    int backbone;
    std::vector<TensorSpec> tensors;
    std::vector<ConvSpec> convs;
    std::vector<Step> steps;
    int tensor_id(const std::string& n) const;
    double flops_per_crop() const;  // 2 * MAC of the contractions (algorithmic, unpadded)
};
Plan build_plan(int backbone);

struct ModelConv {
    DevBuf<__half> packed;  // [KI][NP][Cout_pad][64]
    DevBuf<float> scale, shift;
    CUtensorMap mapB;
    CUtensorMap mapBmc[2];  // BN = 128 convs: one-plane boxes of 64 / 32 rows for the 2- / 4-CTA weight multicast (conv_tc_slab.cuh)
    CUtensorMap mapBh;   // same weights, box of BN / 2 rows: each CTA of a pair loads one half (conv_tc_pair.cuh)
};

class Engine;

class Model {
  public:
    Model(Engine* eng, const float* blob, size_t n_floats);
    std::vector<ModelConv> convs;
    Engine* engine;
    long long id;   // unique per process (keys of the captured graphs)
};

class Engine {
  public:
    Engine(int backbone, int capacity, int precision);
    ~Engine();
    int backbone, cap, np;
    Plan plan;
    cudaStream_t stream = nullptr;
    struct Tensor { DevBuf<__half> buf; long long plane = 0; };
    std::vector<Tensor> tensors;
    DevBuf<float> x, dec, prob;  // (cap,128,128,3), (cap,128,128,3), (cap,128,128)
    DevBuf<float> partial;       // split-K scratch [slices][cap][Cout_pad]
    struct ConvRt { DevBuf<int4> kit, slabs; CUtensorMap mapA[4]; CUtensorMap mapSlab[2]; CUtensorMap mapOut[4]; bool has_out = false; CUtensorMap mapRes; bool has_res = false; };
    std::vector<ConvRt> conv_rt;
    int num_sms = 148;
    bool prof_layers = false; // print per-launch times in profile mode (P2P_PROF_LAYERS)
    int epi_nk = 8;           // layers with at most this many k-iterations per tile trade operand stages for output staging tiles (P2P_EPI_NK)
    int single_acc_steps = 0;   // accumulation chains up to this many k16 steps use ONE TMEM accumulator (P2P_SINGLE_ACC_STEPS).  0 = never: below N = 256 an MMA costs ~75 cycles whatever its width (the 4 KB A-operand read), so the widened two-MMA form (N = 2 BN, then BN) beats three MMAs of width BN also on short chains (3x3 64-ch convs 78 -> 70 us per 256 crops)
    bool res_tma = true;      // residual tiles by TMA into shared memory (P2P_RES_TMA=0 = per-thread loads)
    bool tma_store = true;    // TMA-store epilogue in the persistent kernel (default; P2P_TMA_STORE=0 = direct 16-byte stores)
    int slab_cluster = 1;     // CTAs sharing every weight stage by TMA multicast in the slab kernel (P2P_SLAB_CLUSTER=1|2|4; measured:
                              // 2 gains 1 %, 4 loses 75 % to cluster placement -- the layer is bound by shared-memory bandwidth, not by L2)
    bool slab = true;         // slab-reuse kernel for the Cout-128 k x k convs (P2P_SLAB=0 disables)
    bool pair = true;         // CTA-pair kernel (cta_group::2, M = 256) for the wide decoder convs (P2P_PAIR=0 disables)

    // x_dev -> dec_dev / prob_dev for n <= cap crops; n_active (device int) optionally limits work further.
    void forward(const Model& m, const float* x_dev, int n, float* dec_dev, float* prob_dev, const int* n_active,
                 cudaStream_t s);
    // host in / host out, chunks of `cap`.
    void predict_host(const Model& m, const float* x, int n, float* dec, float* prob);
    void read_tensor(const std::string& name, int n, float* out);  // debug: activation (n,H,W,C) as fp32
    long long launches = 0;  // kernels launched so far (for bench's gpu_launches)
    // Measurement: when `prof` is set, forward() brackets every launch with CUDA events on its stream and
    // adds the elapsed times (ms) to prof[0] (tcgen05 conv kernels) / prof[1] (other kernels); counts in prof[2..3].
    double* prof = nullptr;
    double prof_acc[4] = {0, 0, 0, 0};   // target of p2p_engine_prof_begin / _end (profiling across whole pipeline runs)
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // user timing slots
};

}  // namespace p2p
