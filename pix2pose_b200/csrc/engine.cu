// Host side of the generator: layer tables, plan (topology -> k-iteration tables), TMA maps,
// weight packing (BN folding, fp16 hi/lo split), forward.  See engine.cuh.
#include "engine.cuh"

#include <math.h>

#include <algorithm>
#include <mutex>

#include "conv_tc_pair.cuh"
#include "conv_tc_slab.cuh"
#include "net_kernels.cuh"

namespace p2p {

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const char* get_last_error() { return g_last_error.c_str(); }

void require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(P2P_ERR_NO_DEVICE,
                    fmt("pix2pose_b200 needs a CUDA device (sm_100); none visible (%s). There is no CPU fallback.",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)));
    int dev = 0;
    P2P_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    P2P_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        throw Error(P2P_ERR_NO_DEVICE, fmt("device %d is sm_%d%d; the kernels are built for sm_100a only", dev,
                                           prop.major, prop.minor));
}

// ------------------------------------------------------------------------------------------------
// layer tables (mirrors pix2pose_b200/weights.py:layer_table; Keras topological order)
size_t LayerDef::count() const {
    if (kind == L_BN) return 4 * static_cast<size_t>(shape[0]);
    size_t k = 1;
    const int nd = kind == L_DENSE ? 2 : 4;
    for (int i = 0; i < nd; ++i) k *= shape[i];
    const int nb = kind == L_CONVT ? shape[2] : (kind == L_DENSE ? shape[1] : shape[3]);
    return k + nb;
}

static LayerDef mk(const std::string& n, int kind, int a, int b = 0, int c = 0, int d = 0) {
    LayerDef l;
    l.name = n;
    l.kind = kind;
    l.shape[0] = a; l.shape[1] = b; l.shape[2] = c; l.shape[3] = d;
    return l;
}

static void resnet_block(std::vector<LayerDef>& L, int stage, char block, int cin, int f1, int f2, int f3, bool sc) {
    const std::string base = fmt("res%d%c_branch", stage, block), bnb = fmt("bn%d%c_branch", stage, block);
    L.push_back(mk(base + "2a", L_CONV, 1, 1, cin, f1)); L.push_back(mk(bnb + "2a", L_BN, f1));
    L.push_back(mk(base + "2b", L_CONV, 3, 3, f1, f2));  L.push_back(mk(bnb + "2b", L_BN, f2));
    L.push_back(mk(base + "2c", L_CONV, 1, 1, f2, f3));  L.push_back(mk(bnb + "2c", L_BN, f3));
    if (sc) { L.push_back(mk(base + "1", L_CONV, 1, 1, cin, f3)); L.push_back(mk(bnb + "1", L_BN, f3)); }
}

static void decoder_layers(std::vector<LayerDef>& L, int s3, int s2, int s1) {
    L.push_back(mk("dense_1", L_DENSE, 8 * 8 * 512, 256));
    L.push_back(mk("dense_2", L_DENSE, 256, 8 * 8 * 256));
    L.push_back(mk("convT1", L_CONVT, 5, 5, 256, 256));      L.push_back(mk("bn_convT1", L_BN, 256));
    L.push_back(mk("deconv1", L_CONV, 5, 5, 256 + s3, 256)); L.push_back(mk("bn_deconv1", L_BN, 256));
    L.push_back(mk("convT2", L_CONVT, 5, 5, 128, 256));      L.push_back(mk("bn_convT2", L_BN, 128));
    L.push_back(mk("deconv2", L_CONV, 5, 5, 128 + s2, 256)); L.push_back(mk("bn_deconv2", L_BN, 256));
    L.push_back(mk("convT3", L_CONVT, 5, 5, 64, 256));       L.push_back(mk("bn_convT3", L_BN, 64));
    L.push_back(mk("deconv3", L_CONV, 5, 5, 64 + s1, 128));  L.push_back(mk("bn_deconv3", L_BN, 128));
    L.push_back(mk("convT_xyz", L_CONVT, 5, 5, 3, 128));
    L.push_back(mk("convT_prob", L_CONVT, 5, 5, 1, 128));
}

std::vector<LayerDef> layer_table(int backbone) {
    std::vector<LayerDef> L;
    if (backbone == BB_RESNET50) {
        L.push_back(mk("conv1", L_CONV, 7, 7, 3, 64));
        L.push_back(mk("bn_conv1", L_BN, 64));
        resnet_block(L, 2, 'a', 64, 64, 64, 256, true);
        resnet_block(L, 2, 'b', 256, 64, 64, 256, false);
        resnet_block(L, 2, 'c', 256, 64, 64, 256, false);
        resnet_block(L, 3, 'a', 256, 128, 128, 512, true);
        resnet_block(L, 3, 'b', 512, 128, 128, 512, false);
        resnet_block(L, 3, 'c', 512, 128, 128, 512, false);
        resnet_block(L, 3, 'd', 512, 128, 128, 512, false);
        L.push_back(mk("conv4_1", L_CONV, 5, 5, 512, 256)); L.push_back(mk("bn_conv4_1", L_BN, 256));
        L.push_back(mk("conv4_2", L_CONV, 5, 5, 512, 256)); L.push_back(mk("bn_conv4_2", L_BN, 256));
        decoder_layers(L, 128, 128, 32);
    } else {
        const int lv[4][3] = {{1, 3, 64}, {2, 128, 128}, {3, 256, 128}, {4, 256, 256}};
        for (auto& l : lv)
            for (int br = 1; br <= 2; ++br) {
                const std::string n = fmt("conv%d_%d", l[0], br);
                L.push_back(mk(n, L_CONV, 5, 5, l[1], l[2]));
                L.push_back(mk("bn_" + n, L_BN, l[2]));
            }
        decoder_layers(L, 128, 128, 64);
    }
    return L;
}

size_t param_count(int backbone) {
    size_t n = 0;
    for (auto& l : layer_table(backbone)) n += l.count();
    return n;
}

int parse_backbone(const char* s) {
    if (s && !strcmp(s, "resnet50")) return BB_RESNET50;
    if (s && !strcmp(s, "paper")) return BB_PAPER;
    throw Error(P2P_ERR_INVALID, fmt("backbone must be 'resnet50' or 'paper', got '%s'", s ? s : "(null)"));
}

// ------------------------------------------------------------------------------------------------
// plan
int Plan::tensor_id(const std::string& n) const {
    for (size_t i = 0; i < tensors.size(); ++i)
        if (tensors[i].name == n) return static_cast<int>(i);
    throw Error(P2P_ERR_INTERNAL, "unknown tensor " + n);
}

namespace {

struct PlanBuilder {
    Plan& P;
    explicit PlanBuilder(Plan& p) : P(p) {}
    int T(const std::string& name, int H, int W, int C) {
        P.tensors.push_back({name, H, W, C});
        return static_cast<int>(P.tensors.size()) - 1;
    }
    int id(const std::string& n) { return P.tensor_id(n); }

    ConvSpec& conv(const std::string& name, int kind, int ks, std::vector<SrcSpec> srcs, std::vector<WPart> parts,
                   int out, int act, int res = -1, int c_off = 0) {
        ConvSpec c;
        c.name = name; c.kind = kind; c.ksize = ks; c.srcs = std::move(srcs); c.parts = std::move(parts);
        c.out_tensor = out; c.act = act; c.res_tensor = res; c.c_off = c_off;
        P.convs.push_back(c);
        P.steps.push_back({S_CONV, static_cast<int>(P.convs.size()) - 1, 0, 0, 0});
        return P.convs.back();
    }
    SrcSpec all(int t) { return {t, 0, P.tensors[t].C}; }

    // resnet50_mod.py:40-118
    int bottleneck(int in, int stage, char blk, int f1, int f3, int stride, bool shortcut) {
        const std::string base = fmt("res%d%c_branch", stage, blk), bnb = fmt("bn%d%c_branch", stage, blk);
        const int Hin = P.tensors[in].H, Ho = Hin / stride;
        const std::string tag = fmt("%d%c", stage, blk);
        const int ta = T("t" + tag + "_a", Ho, Ho, f1), tb = T("t" + tag + "_b", Ho, Ho, f1);
        const int out = T("act" + tag, Ho, Ho, f3);
        conv(base + "2a", stride == 2 ? K_CONV_S2 : K_CONV, 1, {all(in)}, {{base + "2a", bnb + "2a"}}, ta, ACT_RELU);
        conv(base + "2b", K_CONV, 3, {all(ta)}, {{base + "2b", bnb + "2b"}}, tb, ACT_RELU);
        static const bool fuse_sc = !getenv("P2P_FUSE_SHORTCUT") || atoi(getenv("P2P_FUSE_SHORTCUT")) != 0;
        if (shortcut && fuse_sc) {
            // conv_block (resnet50_mod.py:76-118): relu(BN(conv2c(tb)) + BN(conv1(x))) as ONE contraction over K = [tb | x]
            // (x sampled at every second pixel in the stride-2 block); saves writing and re-reading the shortcut tensor
            SrcSpec sx = all(in);
            sx.view = stride == 2 ? 1 : 0;
            ConvSpec& c = conv(base + "2c", K_CONV, 1, {all(tb), sx}, {{base + "2c", bnb + "2c"}, {base + "1", bnb + "1"}}, out, ACT_RELU);
            c.kcat = true;
            return out;
        }
        int res = in;
        if (shortcut) {
            res = T("sc" + tag, Ho, Ho, f3);
            conv(base + "1", stride == 2 ? K_CONV_S2 : K_CONV, 1, {all(in)}, {{base + "1", bnb + "1"}}, res, ACT_NONE);
        }
        conv(base + "2c", K_CONV, 1, {all(tb)}, {{base + "2c", bnb + "2c"}}, out, ACT_RELU, res);
        return out;
    }

    void decoder(int f4, SrcSpec s3, SrcSpec s2, SrcSpec s1) {
        const int enc = T("enc", 1, 1, 256), d0 = T("d0", 8, 8, 256);
        conv("dense_1", K_DENSE, 1, {all(f4)}, {{"dense_1", ""}}, enc, ACT_NONE);
        conv("dense_2", K_DENSE, 1, {all(enc)}, {{"dense_2", ""}}, d0, ACT_NONE);
        const int d1 = T("d1", 16, 16, 256), d1u = T("d1_uni", 16, 16, 256);
        conv("convT1", K_CONVT, 5, {all(d0)}, {{"convT1", "bn_convT1"}}, d1, ACT_LRELU);
        conv("deconv1", K_CONV, 5, {all(d1), s3}, {{"deconv1", "bn_deconv1"}}, d1u, ACT_LRELU);
        const int d2 = T("d2", 32, 32, 128), d2u = T("d2_uni", 32, 32, 256);
        conv("convT2", K_CONVT, 5, {all(d1u)}, {{"convT2", "bn_convT2"}}, d2, ACT_LRELU);
        conv("deconv2", K_CONV, 5, {all(d2), s2}, {{"deconv2", "bn_deconv2"}}, d2u, ACT_LRELU);
        const int d3 = T("d3", 64, 64, 64), d3u = T("d3_uni", 64, 64, 128);
        conv("convT3", K_CONVT_FUSED, 5, {all(d2u)}, {{"convT3", "bn_convT3"}}, d3, ACT_LRELU);  // Cout 64: N = 4 x 64 fills the tile
        conv("deconv3", K_CONV, 5, {all(d3), s1}, {{"deconv3", "bn_deconv3"}}, d3u, ACT_LRELU);
        conv("heads", K_CONVT_FUSED, 5, {all(d3u)}, {{"convT_xyz", ""}, {"convT_prob", ""}}, -1, ACT_HEADS);
    }
};

int find_layer(const std::vector<LayerDef>& L, const std::string& n) {
    for (size_t i = 0; i < L.size(); ++i)
        if (L[i].name == n) return static_cast<int>(i);
    throw Error(P2P_ERR_INTERNAL, "unknown layer " + n);
}

int layer_cout(const LayerDef& l) { return l.kind == L_CONVT ? l.shape[2] : (l.kind == L_DENSE ? l.shape[1] : l.shape[3]); }

void finalize_conv(const Plan& P, const std::vector<LayerDef>& L, ConvSpec& c) {
    // channel bookkeeping
    c.Cin = 0;
    for (auto& s : c.srcs) c.Cin += s.c_count;
    c.Cout = 0;
    for (auto& w : c.parts) c.Cout += layer_cout(L[find_layer(L, w.layer)]);
    if (c.kcat) c.Cout = layer_cout(L[find_layer(L, c.parts[0].layer)]);  // parts add up along K, not along N
    if (c.Cout < 64) { c.BN = 16; c.Cout_pad = 16; }
    else if (c.Cout % 128 == 0) { c.BN = 128; c.Cout_pad = c.Cout; }
    else { c.BN = 64; c.Cout_pad = (c.Cout + 63) / 64 * 64; }

    const TensorSpec& t0 = P.tensors[c.srcs[0].tensor];
    auto chunks = [](int cnt) { return (cnt + 63) / 64; };
    c.kit.clear(); c.kw.clear(); c.maps.clear();
    c.phases = 1; c.sy = c.sx = 1;
    for (int i = 0; i < 4; ++i) c.oy_off[i] = c.ox_off[i] = 0;

    if (c.kind == K_CONV || c.kind == K_PATCH) {
        c.H = t0.H; c.W = t0.W;
        const int pad = (c.ksize - 1) / 2;
        int concat_off = 0;
        for (size_t si = 0; si < c.srcs.size(); ++si) {
            const SrcSpec& s = c.srcs[si];
            c.maps.push_back({s.tensor, s.view, s.c_begin + s.c_count});
            const int ks = c.kind == K_PATCH ? 1 : c.ksize;
            for (int kh = 0; kh < ks; ++kh)
                for (int kw = 0; kw < ks; ++kw)
                    for (int ch = 0; ch < chunks(s.c_count); ++ch) {
                        c.kit.push_back(make_int4(static_cast<int>(si), c.kind == K_PATCH ? 0 : kh - pad,
                                                  c.kind == K_PATCH ? 0 : kw - pad, s.c_begin + ch * 64));
                        c.kw.push_back({kh, kw, (c.kcat ? 0 : concat_off) + ch * 64, std::min(64, s.c_count - ch * 64), c.kcat ? static_cast<int>(si) : -1});
                    }
            concat_off += s.c_count;
        }
        c.kstart[1] = static_cast<int>(c.kit.size());
    } else if (c.kind == K_STEM) {
        // Direct stride-2 stem conv on the 4-channel padded input (net_kernels.cuh stem_convert_kernel).  One k-iteration
        // per kernel row kh; its K slice is the window (kw, c4) = ksize * 4 fp16 values starting at padded pixel 2*ox, which
        // the overlapping tensor-map views 7 / 8 (even / odd input rows) deliver as the "channel" dimension.  Input row
        // 2*oy + kh - pad = 2*(oy + dy) + parity; rows outside the image are zero-filled by TMA.
        c.H = 64; c.W = 64;
        const SrcSpec& s = c.srcs[0];
        const int pad = c.ksize == 7 ? 3 : 1;  // ZeroPadding2D(3)+valid (resnet50_mod.py:200) / TF 'same' k5 s2: 1 before
        c.maps.push_back({s.tensor, 7, 64});
        c.maps.push_back({s.tensor, 8, 64});
        for (int kh = 0; kh < c.ksize; ++kh) {
            const int d = kh - pad, par = ((d % 2) + 2) % 2, dy = (d - par) / 2;
            c.kit.push_back(make_int4(par, dy, 0, 0));
            c.kw.push_back({kh, -1, 0, c.ksize * 4});
        }
        c.kstart[1] = static_cast<int>(c.kit.size());
    } else if (c.kind == K_CONV_S2) {
        c.H = t0.H / 2; c.W = t0.W / 2;
        const SrcSpec& s = c.srcs[0];
        if (c.ksize == 1) {
            c.maps.push_back({s.tensor, 1, s.c_begin + s.c_count});
            for (int ch = 0; ch < chunks(s.c_count); ++ch) {
                c.kit.push_back(make_int4(0, 0, 0, s.c_begin + ch * 64));
                c.kw.push_back({0, 0, ch * 64, std::min(64, s.c_count - ch * 64)});
            }
        } else {
            // TF 'same', k=5, s=2, even size: pad 1 before / 2 after -> input index = 2*o + k - 1
            for (int v = 1; v <= 4; ++v) c.maps.push_back({s.tensor, v, s.c_begin + s.c_count});
            for (int kh = 0; kh < 5; ++kh)
                for (int kw = 0; kw < 5; ++kw) {
                    const int py = (kh + 1) & 1, px = (kw + 1) & 1;
                    const int dy = (kh - 1 - py) / 2, dx = (kw - 1 - px) / 2;  // exact: (k-1-p) is even
                    for (int ch = 0; ch < chunks(s.c_count); ++ch) {
                        c.kit.push_back(make_int4(py * 2 + px, dy, dx, s.c_begin + ch * 64));
                        c.kw.push_back({kh, kw, ch * 64, std::min(64, s.c_count - ch * 64)});
                    }
                }
        }
        c.kstart[1] = static_cast<int>(c.kit.size());
    } else if (c.kind == K_CONVT) {
        // out[o] = sum_{i,k: 2i+k = o+1} x[i] w[k]  (full scatter cropped [1:2N+1])
        //   o = 2m   : (k=1,i=m) (k=3,i=m-1)          o = 2m+1 : (k=0,i=m+1) (k=2,i=m) (k=4,i=m-1)
        c.H = t0.H; c.W = t0.W; c.phases = 4; c.sy = c.sx = 2;
        const SrcSpec& s = c.srcs[0];
        c.maps.push_back({s.tensor, 0, s.c_begin + s.c_count});
        const int tk[2][3] = {{1, 3, -1}, {0, 2, 4}}, td[2][3] = {{0, -1, 0}, {1, 0, -1}}, tn[2] = {2, 3};
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                const int z = a * 2 + b;
                c.oy_off[z] = a; c.ox_off[z] = b;
                for (int iy = 0; iy < tn[a]; ++iy)
                    for (int ix = 0; ix < tn[b]; ++ix)
                        for (int ch = 0; ch < chunks(s.c_count); ++ch) {
                            c.kit.push_back(make_int4(0, td[a][iy], td[b][ix], s.c_begin + ch * 64));
                            c.kw.push_back({tk[a][iy], tk[b][ix], ch * 64, std::min(64, s.c_count - ch * 64)});
                        }
                c.kstart[z + 1] = static_cast<int>(c.kit.size());
            }
    } else if (c.kind == K_CONVT_FUSED) {
        // All four output phases in one GEMM: N = 4 phases x Cout (tiny heads), K = 3x3 input neighbourhood.
        // Phase a (output row parity) uses kernel row kh(a,dy): a=0: dy=0->1, dy=-1->3 ; a=1: dy=+1->0, dy=0->2, dy=-1->4.
        c.H = t0.H; c.W = t0.W; c.sy = c.sx = 2;
        const SrcSpec& s = c.srcs[0];
        c.maps.push_back({s.tensor, 0, s.c_begin + s.c_count});
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
                for (int ch = 0; ch < chunks(s.c_count); ++ch) {
                    c.kit.push_back(make_int4(0, dy, dx, s.c_begin + ch * 64));
                    c.kw.push_back({dy, dx, ch * 64, std::min(64, s.c_count - ch * 64)});
                }
        c.kstart[1] = static_cast<int>(c.kit.size());
        c.Cout *= 4;   // accumulator columns
    } else if (c.kind == K_DENSE) {
        c.H = 1; c.W = 1;
        const SrcSpec& s = c.srcs[0];
        const int K = t0.H * t0.W * t0.C;
        c.Cin = K;
        c.maps.push_back({s.tensor, 5, K});
        for (int ch = 0; ch < K / 64; ++ch) {
            c.kit.push_back(make_int4(0, 0, 0, ch * 64));
            c.kw.push_back({0, 0, ch * 64, 64});
        }
        c.kstart[1] = static_cast<int>(c.kit.size());
    }
    if (c.kind == K_CONVT_FUSED) {
        if (c.act == ACT_HEADS) { c.BN = 16; c.Cout_pad = 16; }
        else { c.BN = 128; c.Cout_pad = (c.Cout + 127) / 128 * 128; }
    }
    for (size_t i = 0; i < c.kit.size(); ++i) c.kit[i].x |= ((c.kw[i].nvalid + 15) / 16) << 8;  // k16 steps actually needed
    if (c.kind == K_DENSE && c.kit.size() >= 64) {  // dense_1: K = 32768 -> 32 slices of 16 k-iterations
        c.splitk_chunk = 16;
        c.splitk = static_cast<int>((c.kit.size() + c.splitk_chunk - 1) / c.splitk_chunk);
    }
    {
        // N = 256 tiles for wide layers with a long K loop: the tensor-pipe-bound convs are power / clock limited and an
        // N = 256 MMA reads 25 % less shared memory per FLOP (measured: deconv2 1.77 -> 1.55 ms per 256 crops).  Short-K
        // layers stay at 128: they are epilogue-bound and BN = 256 leaves no shared memory for staging tiles.
        const char* e = getenv("P2P_BN256");
        const int min_kit = e ? atoi(e) : 100;   // P2P_BN256=0 disables, otherwise = minimum k-iterations per tile
        static const bool fused256 = !getenv("P2P_BN256_FUSED") || atoi(getenv("P2P_BN256_FUSED")) != 0;  // convT3: 659 -> 592 us per 256 crops
        if (min_kit > 0 && c.BN == 128 && c.Cout % 256 == 0 && c.splitk <= 1 && c.phases == 1 &&
            (c.kind != K_CONVT_FUSED || (fused256 && c.act != ACT_HEADS)) &&
            static_cast<int>(c.kit.size()) >= (c.kind == K_CONVT_FUSED ? 1 : min_kit)) {
            c.BN = 256; c.Cout_pad = c.Cout;
        }
    }
    if (c.phases == 1)
        for (int z = 1; z < 5; ++z) c.kstart[z] = c.kstart[1];
    c.tw = std::min(c.W, 16);
    c.th = std::min(c.H, 128 / c.tw);
    c.nb = 128 / (c.tw * c.th);
    // slab-reuse kernel (conv_tc_slab.cuh): 8 x 16 pixel tiles, one entry per (source, 64-channel chunk)
    c.slab = false;
    c.slabs.clear();
    static const bool slab_on = !getenv("P2P_SLAB") || atoi(getenv("P2P_SLAB")) != 0;
    static const bool slab64 = !getenv("P2P_SLAB64") || atoi(getenv("P2P_SLAB64")) != 0;   // 3x3 64->64 convs of the stage-2 bottlenecks
    if (slab_on && c.kind == K_CONV && c.ksize >= 3 && c.ksize <= 5 && (c.BN == 128 || (c.BN == 64 && slab64 && c.Cout == 64)) && c.W % 8 == 0 &&
        c.H % 16 == 0 && c.srcs.size() <= 2 && c.res_tensor < 0 && c.splitk <= 1) {
        c.slab = true;
        c.slab_ks = c.ksize;
    }
    // The heads (phase-fused transposed conv, N = 16) are a 3 x 3 stride-1 conv over d3_uni from the GEMM's point of view, so
    // they CAN run on the slab kernel (activations fetched 3 x instead of 9 x, weights resident in shared memory).  Measured:
    // no gain (365 us generic, 376 us slab per 256 crops) -- the layer is bound by its 144 MMAs per 128-pixel tile at ~87
    // cycles each (an M = 128 MMA reads its 4 KB A operand from shared memory at ~64 B/clk whatever N is), not by operand
    // traffic or load latency.  Kept as an experiment: P2P_SLAB_HEADS=1.
    static const bool slab_heads = getenv("P2P_SLAB_HEADS") && atoi(getenv("P2P_SLAB_HEADS")) != 0;
    if (slab_on && slab_heads && c.kind == K_CONVT_FUSED && c.act == ACT_HEADS && c.W % 8 == 0 && c.H % 16 == 0 && c.srcs.size() == 1) {
        c.slab = true;
        c.slab_ks = 3;
    }
    if (c.slab) {
        c.tw = 8; c.th = 16; c.nb = 1;
        int src_base = 0;
        for (size_t si = 0; si < c.srcs.size(); ++si) {
            const SrcSpec& sp = c.srcs[si];
            const int nch = chunks(sp.c_count);
            for (int ch = 0; ch < nch; ++ch) {
                const int nvalid = std::min(64, sp.c_count - ch * 64);
                c.slabs.push_back(make_int4(static_cast<int>(si) | (((nvalid + 15) / 16) << 8), sp.c_begin + ch * 64, src_base + ch, nch));
            }
            src_base += c.slab_ks * c.slab_ks * nch;
        }
    }
}

}  // namespace

// padded row width of the stem input in pixels: the last window starts at pixel 2 * 63 and is 16 pixels wide
constexpr int kStemWp = 144;
// The stem's K slice per kernel row is ksize * 4 <= 32 fp16 values: with 64-byte windows (SWIZZLE_64B A tiles) TMA moves half
// the bytes of the 128-byte ones (the stem is bound by exactly that traffic: the overlapping windows are re-fetched per pixel).
static bool stem_sw64() {
    const char* e = getenv("P2P_STEM_SW64");
    return !e || atoi(e) != 0;
}
static bool stem_direct() {
    const char* e = getenv("P2P_STEM_DIRECT");  // 0 = im2col + 1x1 GEMM (the first implementation)
    return !e || atoi(e) != 0;
}

Plan build_plan(int backbone) {
    Plan P;
    P.backbone = backbone;
    PlanBuilder B(P);
    if (backbone == BB_RESNET50) {
        const int f1 = B.T("f1", 64, 64, 64);
        if (stem_direct()) {
            const int xpad = B.T("xpad", 128, kStemWp, 4);
            P.steps.push_back({S_STEMCVT, xpad, 3, 0, 0});
            B.conv("conv1", K_STEM, 7, {B.all(xpad)}, {{"conv1", "bn_conv1"}}, f1, ACT_RELU);
        } else {
            const int patches = B.T("patches", 64, 64, 192);
            P.steps.push_back({S_IM2COL, patches, 7, 3, 192});
            B.conv("conv1", K_PATCH, 7, {B.all(patches)}, {{"conv1", "bn_conv1"}}, f1, ACT_RELU);
        }
        const int pool = B.T("pool1", 32, 32, 64);
        P.steps.push_back({S_MAXPOOL, f1, pool, 0, 0});
        int x = B.bottleneck(pool, 2, 'a', 64, 256, 1, true);
        x = B.bottleneck(x, 2, 'b', 64, 256, 1, false);
        const int f2 = B.bottleneck(x, 2, 'c', 64, 256, 1, false);
        x = B.bottleneck(f2, 3, 'a', 128, 512, 2, true);
        x = B.bottleneck(x, 3, 'b', 128, 512, 1, false);
        x = B.bottleneck(x, 3, 'c', 128, 512, 1, false);
        const int f3 = B.bottleneck(x, 3, 'd', 128, 512, 1, false);
        const int f4 = B.T("f4", 8, 8, 512);
        B.conv("conv4", K_CONV_S2, 5, {B.all(f3)}, {{"conv4_1", "bn_conv4_1"}, {"conv4_2", "bn_conv4_2"}}, f4, ACT_LRELU);
        B.decoder(f4, {f3, 0, 128}, {f2, 0, 128}, {f1, 0, 32});  // ae_model.py:186-188 channel slices
    } else {
        const int f1 = B.T("f1", 64, 64, 128), f2 = B.T("f2", 32, 32, 256), f3 = B.T("f3", 16, 16, 256);
        const int f4 = B.T("f4", 8, 8, 512);
        if (stem_direct()) {
            const int xpad = B.T("xpad", 128, kStemWp, 4);
            P.steps.push_back({S_STEMCVT, xpad, 1, 0, 0});
            B.conv("conv1", K_STEM, 5, {B.all(xpad)}, {{"conv1_1", "bn_conv1_1"}, {"conv1_2", "bn_conv1_2"}}, f1, ACT_LRELU);
        } else {
            const int patches = B.T("patches", 64, 64, 128);
            P.steps.push_back({S_IM2COL, patches, 5, 1, 128});
            B.conv("conv1", K_PATCH, 5, {B.all(patches)}, {{"conv1_1", "bn_conv1_1"}, {"conv1_2", "bn_conv1_2"}}, f1, ACT_LRELU);
        }
        B.conv("conv2", K_CONV_S2, 5, {B.all(f1)}, {{"conv2_1", "bn_conv2_1"}, {"conv2_2", "bn_conv2_2"}}, f2, ACT_LRELU);
        B.conv("conv3", K_CONV_S2, 5, {B.all(f2)}, {{"conv3_1", "bn_conv3_1"}, {"conv3_2", "bn_conv3_2"}}, f3, ACT_LRELU);
        B.conv("conv4", K_CONV_S2, 5, {B.all(f3)}, {{"conv4_1", "bn_conv4_1"}, {"conv4_2", "bn_conv4_2"}}, f4, ACT_LRELU);
        B.decoder(f4, {f3, 128, 128}, {f2, 128, 128}, {f1, 64, 64});  // second twin of each level (ae_model.py:115,125,135)
    }
    const std::vector<LayerDef> L = layer_table(backbone);
    for (auto& c : P.convs) finalize_conv(P, L, c);
    return P;
}

double Plan::flops_per_crop() const {
    const std::vector<LayerDef> L = layer_table(backbone);
    double mac = 0;
    for (auto& c : convs) {
        for (auto& w : c.parts) {
            const LayerDef& l = L[find_layer(L, w.layer)];
            double per_out;
            double outs;
            if (l.kind == L_DENSE) { per_out = l.shape[0]; outs = l.shape[1]; }
            else if (l.kind == L_CONVT) { per_out = 25.0 / 4.0 * l.shape[3]; outs = 4.0 * c.H * c.W * l.shape[2]; }
            else { per_out = static_cast<double>(l.shape[0]) * l.shape[1] * l.shape[2]; outs = static_cast<double>(c.H) * c.W * l.shape[3]; }
            mac += per_out * outs;
        }
    }
    return 2.0 * mac;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) throw Error(P2P_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    return fn;
}

void encode(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, bool sw64 = false) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides_bytes, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(P2P_ERR_CUDA, fmt("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu %llu %llu strides %llu %llu %llu %llu box %u %u %u %u %u base %p)",
                                      static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                                      (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)dims[4],
                                      (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
                                      (unsigned long long)strides_bytes[2], (unsigned long long)strides_bytes[3], box[0], box[1], box[2],
                                      box[3], box[4], base));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// engine
Engine::Engine(int bb, int capacity, int precision) : backbone(bb), cap(capacity), np(precision == PREC_FP16 ? 1 : 2) {
    require_device();
    P2P_CHECK(capacity >= 1 && capacity <= 4096, "capacity must be in [1,4096], got %d", capacity);
    P2P_CHECK(precision == PREC_FP16X3 || precision == PREC_FP16, "unknown precision %d", precision);
    plan = build_plan(bb);
    int dev = 0;
    P2P_CUDA(cudaGetDevice(&dev));
    P2P_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    P2P_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (const char* e = getenv("P2P_PAIR")) pair = atoi(e) != 0;
    if (const char* e = getenv("P2P_SLAB")) slab = atoi(e) != 0;
    if (const char* e = getenv("P2P_SLAB_CLUSTER")) slab_cluster = atoi(e);
    if (const char* e = getenv("P2P_TMA_STORE")) tma_store = atoi(e) != 0;
    if (const char* e = getenv("P2P_SINGLE_ACC_STEPS")) single_acc_steps = atoi(e);
    if (const char* e = getenv("P2P_EPI_NK")) epi_nk = atoi(e);
    if (const char* e = getenv("P2P_RES_TMA")) res_tma = atoi(e) != 0;
    if (const char* e = getenv("P2P_PROF_LAYERS")) prof_layers = atoi(e) != 0;

    tensors.resize(plan.tensors.size());
    for (size_t i = 0; i < plan.tensors.size(); ++i) {
        const TensorSpec& t = plan.tensors[i];
        const size_t plane = static_cast<size_t>(cap) * t.H * t.W * t.C;
        tensors[i].buf.alloc(plane * np);
        tensors[i].plane = np == 2 ? static_cast<long long>(plane) : 0;
        P2P_CUDA(cudaMemset(tensors[i].buf.p, 0, plane * np * sizeof(__half)));
    }
    x.alloc(static_cast<size_t>(cap) * 128 * 128 * 3);
    dec.alloc(static_cast<size_t>(cap) * 128 * 128 * 3);
    prob.alloc(static_cast<size_t>(cap) * 128 * 128);

    {
        size_t need = 0;
        for (auto& c : plan.convs)
            if (c.splitk > 1) need = std::max(need, static_cast<size_t>(c.splitk) * cap * c.Cout_pad);
        if (need) partial.alloc(need);
    }
    conv_rt.resize(plan.convs.size());
    for (size_t ci = 0; ci < plan.convs.size(); ++ci) {
        const ConvSpec& c = plan.convs[ci];
        ConvRt& rt = conv_rt[ci];
        rt.kit.upload(c.kit.data(), c.kit.size());
        memset(rt.mapA, 0, sizeof(rt.mapA));
        for (size_t mi = 0; mi < c.maps.size(); ++mi) {
            const ConvSpec::MapReq& rq = c.maps[mi];
            const TensorSpec& t = plan.tensors[rq.tensor];
            __half* base = tensors[rq.tensor].buf.p;
            const cuuint64_t C = t.C, W = t.W, H = t.H;
            const cuuint64_t plane_b = static_cast<cuuint64_t>(cap) * H * W * C * 2;
            cuuint64_t dims[5], str[4];
            cuuint32_t box[5] = {64, (cuuint32_t)c.tw, (cuuint32_t)c.th, (cuuint32_t)c.nb, (cuuint32_t)np};
            bool sw64 = false;
            if (rq.view == 0) {
                dims[0] = rq.climit; dims[1] = W; dims[2] = H; dims[3] = cap; dims[4] = np;
                str[0] = C * 2; str[1] = W * C * 2; str[2] = H * W * C * 2; str[3] = plane_b;
            } else if (rq.view == 7 || rq.view == 8) {
                // stem input (N, 128, Wp, 4): dim0 = 64 consecutive fp16 values (16 pixels), dim1 = output column (window start
                // advances 2 pixels = 16 bytes: overlapping windows), dim2 = input rows of one parity
                base += static_cast<size_t>(rq.view - 7) * W * C;
                sw64 = stem_sw64();
                dims[0] = box[0] = sw64 ? 32 : 64;
                dims[1] = 64; dims[2] = H / 2; dims[3] = cap; dims[4] = np;
                str[0] = 16; str[1] = 2 * W * C * 2; str[2] = H * W * C * 2; str[3] = plane_b;
            } else if (rq.view >= 1 && rq.view <= 4) {
                const int py = (rq.view - 1) >> 1, px = (rq.view - 1) & 1;
                base += (static_cast<size_t>(py) * W + px) * C;
                dims[0] = rq.climit; dims[1] = W / 2; dims[2] = H / 2; dims[3] = cap; dims[4] = np;
                str[0] = 2 * C * 2; str[1] = 2 * W * C * 2; str[2] = H * W * C * 2; str[3] = plane_b;
            } else {
                const cuuint64_t K = H * W * C;
                dims[0] = K; dims[1] = 1; dims[2] = 1; dims[3] = cap; dims[4] = np;
                str[0] = K * 2; str[1] = K * 2; str[2] = K * 2; str[3] = plane_b;
            }
            encode(&rt.mapA[mi], base, 5, dims, str, box, sw64);
        }
        for (size_t mi = c.maps.size(); mi < 4; ++mi) rt.mapA[mi] = rt.mapA[0];
        memset(rt.mapOut, 0, sizeof(rt.mapOut));
        rt.has_out = false;
        if (c.act != ACT_HEADS && c.splitk <= 1 && c.out_tensor >= 0 && c.BN >= 64) {
            const TensorSpec& to = plan.tensors[c.out_tensor];
            __half* obase = tensors[c.out_tensor].buf.p;
            const cuuint64_t OC = to.C, OW = to.W, OH = to.H;
            const cuuint64_t oplane_b = static_cast<cuuint64_t>(cap) * OH * OW * OC * 2;
            cuuint32_t box[5] = {64, (cuuint32_t)c.tw, (cuuint32_t)c.th, (cuuint32_t)c.nb, (cuuint32_t)np};
            if (c.kind == K_DENSE) {
                const cuuint64_t K = OH * OW * OC;
                cuuint64_t dims[5] = {K, 1, 1, (cuuint64_t)cap, (cuuint64_t)np};
                cuuint64_t str[4] = {K * 2, K * 2, K * 2, oplane_b};
                encode(&rt.mapOut[0], obase, 5, dims, str, box);
                for (int i = 1; i < 4; ++i) rt.mapOut[i] = rt.mapOut[0];
            } else if (c.sy == 1) {
                cuuint64_t dims[5] = {OC, OW, OH, (cuuint64_t)cap, (cuuint64_t)np};
                cuuint64_t str[4] = {OC * 2, OW * OC * 2, OH * OW * OC * 2, oplane_b};
                encode(&rt.mapOut[0], obase, 5, dims, str, box);
                for (int i = 1; i < 4; ++i) rt.mapOut[i] = rt.mapOut[0];
            } else {
                for (int ph = 0; ph < 4; ++ph) {  // output pixel (2y+a, 2x+b): strided view of the output tensor
                    const int a = ph >> 1, b = ph & 1;
                    cuuint64_t dims[5] = {OC, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)cap, (cuuint64_t)np};
                    cuuint64_t str[4] = {2 * OC * 2, 2 * OW * OC * 2, OH * OW * OC * 2, oplane_b};
                    encode(&rt.mapOut[ph], obase + (static_cast<size_t>(a) * OW + b) * OC, 5, dims, str, box);
                }
            }
            rt.has_out = true;
        }
        memset(&rt.mapRes, 0, sizeof(rt.mapRes));
        rt.has_res = false;
        if (rt.has_out && c.res_tensor >= 0 && c.sy == 1 && c.kind != K_DENSE && c.kind != K_CONVT_FUSED && c.Cout % c.BN == 0) {
            // residual tensor viewed exactly like the output tensor: the producer warp fetches the (64 ch, tw, th, nb, plane)
            // box of every output slice into shared memory
            const TensorSpec& tr = plan.tensors[c.res_tensor];
            const TensorSpec& to = plan.tensors[c.out_tensor];
            if (tr.H == to.H && tr.W == to.W && tr.C >= c.Cout) {
                const cuuint64_t RC = tr.C, RW = tr.W, RH = tr.H;
                cuuint32_t box[5] = {64, (cuuint32_t)c.tw, (cuuint32_t)c.th, (cuuint32_t)c.nb, (cuuint32_t)np};
                cuuint64_t dims[5] = {RC, RW, RH, (cuuint64_t)cap, (cuuint64_t)np};
                cuuint64_t str[4] = {RC * 2, RW * RC * 2, RH * RW * RC * 2, static_cast<cuuint64_t>(cap) * RH * RW * RC * 2};
                encode(&rt.mapRes, tensors[c.res_tensor].buf.p, 5, dims, str, box);
                rt.has_res = true;
            }
        }
        memset(rt.mapSlab, 0, sizeof(rt.mapSlab));
        if (c.slab) {
            rt.slabs.upload(c.slabs.data(), c.slabs.size());
            const int pad = (c.slab_ks - 1) / 2;
            for (size_t si = 0; si < c.srcs.size(); ++si) {
                const SrcSpec& sp = c.srcs[si];
                const TensorSpec& t = plan.tensors[sp.tensor];
                const cuuint64_t C = t.C, W = t.W, H = t.H;
                cuuint64_t dims[5] = {(cuuint64_t)(sp.c_begin + sp.c_count), W, H, (cuuint64_t)cap, (cuuint64_t)np};
                cuuint64_t str[4] = {C * 2, W * C * 2, H * W * C * 2, static_cast<cuuint64_t>(cap) * H * W * C * 2};
                cuuint32_t box[5] = {64, 8, (cuuint32_t)(16 + 2 * pad), 1, (cuuint32_t)np};
                encode(&rt.mapSlab[si], tensors[sp.tensor].buf.p, 5, dims, str, box);
            }
            if (c.srcs.size() < 2) rt.mapSlab[1] = rt.mapSlab[0];
        }
    }
    P2P_CUDA(cudaDeviceSynchronize());
}

Engine::~Engine() {
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------------
// model = packed weights for one object
static long long next_model_id() {
    static long long n = 0;
    return ++n;
}

Model::Model(Engine* eng, const float* blob, size_t n_floats) : engine(eng), id(next_model_id()) {
    const std::vector<LayerDef> L = layer_table(eng->backbone);
    const size_t want = param_count(eng->backbone);
    P2P_CHECK(n_floats == want, "weight blob has %zu floats, backbone needs %zu", n_floats, want);
    std::vector<const float*> ptr(L.size());
    {
        const float* q = blob;
        for (size_t i = 0; i < L.size(); ++i) { ptr[i] = q; q += L[i].count(); }
    }
    const int np = eng->np;
    convs.resize(eng->plan.convs.size());
    for (size_t ci = 0; ci < eng->plan.convs.size(); ++ci) {
        const ConvSpec& c = eng->plan.convs[ci];
        ModelConv& mc = convs[ci];
        struct PartW { const float* k; const float* b; const float* bn; int cout; const LayerDef* l; std::vector<float> fold; };
        std::vector<PartW> pw;
        float wmax = 0.f;
        for (auto& w : c.parts) {
            const int li = find_layer(L, w.layer);
            PartW q;
            q.l = &L[li];
            q.k = ptr[li];
            q.cout = layer_cout(L[li]);
            q.b = ptr[li] + (L[li].count() - q.cout);
            q.bn = w.bn.empty() ? nullptr : ptr[find_layer(L, w.bn)];
            const size_t nk = L[li].count() - q.cout;
            if (c.kcat) {
                // K-concatenated parts share one epilogue, so each part's BN scale is folded into its weights
                // (conv kernels are (kh,kw,Cin,Cout): the output channel is the fastest index)
                P2P_CHECK(q.bn && q.l->kind == L_CONV, "K-concatenated parts must be Conv2D + BatchNormalization");
                q.fold.resize(q.cout);
                for (int co = 0; co < q.cout; ++co) q.fold[co] = q.bn[co] / sqrtf(q.bn[3 * q.cout + co] + 1e-3f);
                for (size_t i = 0; i < nk; ++i) wmax = std::max(wmax, fabsf(q.k[i] * q.fold[i % q.cout]));
            } else {
                for (size_t i = 0; i < nk; ++i) wmax = std::max(wmax, fabsf(q.k[i]));
            }
            pw.push_back(q);
        }
        // power-of-two pre-scale so |w| <= 128: keeps the fp16 lo parts out of the subnormal range
        const float wscale = wmax > 0.f ? exp2f(floorf(log2f(128.f / wmax))) : 1.f;
        const size_t KI = c.kit.size();
        std::vector<__half> packed(KI * np * c.Cout_pad * 64, __float2half(0.f));
        if (c.kind == K_CONVT_FUSED) {
            const int khmap[2][3] = {{3, 1, -1}, {4, 2, 0}};  // [phase parity][d + 1] -> kernel index (or -1)
            int tot = 0;
            for (auto& q : pw) tot += q.cout;
            P2P_CHECK(tot * 4 <= c.Cout_pad, "fused transposed conv: too many output channels");
            for (size_t it = 0; it < KI; ++it) {
                const KWeight& kw = c.kw[it];
                for (int ph = 0; ph < 4; ++ph) {
                    const int kh = khmap[ph >> 1][kw.kh + 1], kwi = khmap[ph & 1][kw.kw + 1];
                    if (kh < 0 || kwi < 0) continue;
                    int ch_base = 0;
                    for (auto& q : pw) {
                        const LayerDef& l = *q.l;
                        for (int co = 0; co < q.cout; ++co) {
                            const int n = ph * tot + ch_base + co;
                            for (int j = 0; j < kw.nvalid; ++j) {
                                const int cin = kw.cin_begin + j;
                                const float v = q.k[((static_cast<size_t>(kh) * 5 + kwi) * l.shape[2] + co) * l.shape[3] + cin] * wscale;
                                const __half h = __float2half_rn(v);
                                const size_t o = ((it * np + 0) * c.Cout_pad + n) * 64 + j;
                                packed[o] = h;
                                if (np == 2) packed[o + static_cast<size_t>(c.Cout_pad) * 64] = __float2half_rn(v - __half2float(h));
                            }
                        }
                        ch_base += q.cout;
                    }
                }
            }
        }
        for (size_t it = 0; c.kind != K_CONVT_FUSED && it < KI; ++it) {
            const KWeight& kw = c.kw[it];
            int n_base = 0;
            for (size_t pi = 0; pi < pw.size(); ++pi) {
                const PartW& q = pw[pi];
                if (c.kcat && static_cast<int>(pi) != kw.part) continue;   // this K slice belongs to one part only
                const LayerDef& l = *q.l;
                for (int co = 0; co < q.cout; ++co) {
                    const int n = n_base + co;
                    for (int j = 0; j < kw.nvalid; ++j) {
                        const int cin = kw.cin_begin + j;
                        float v;
                        if (c.kind == K_DENSE) {
                            v = q.k[static_cast<size_t>(cin) * l.shape[1] + co];
                        } else if (c.kind == K_PATCH) {
                            if (cin >= l.shape[0] * l.shape[1] * l.shape[2]) continue;
                            v = q.k[static_cast<size_t>(cin) * l.shape[3] + co];
                        } else if (c.kind == K_STEM) {
                            const int kx = j / 4, ci = j % 4;  // K slice of a kernel row: (kw, channel padded to 4)
                            if (ci == 3) continue;
                            v = q.k[((static_cast<size_t>(kw.kh) * l.shape[1] + kx) * l.shape[2] + ci) * l.shape[3] + co];
                        } else if (l.kind == L_CONVT) {
                            v = q.k[((static_cast<size_t>(kw.kh) * 5 + kw.kw) * l.shape[2] + co) * l.shape[3] + cin];
                        } else {
                            v = q.k[((static_cast<size_t>(kw.kh) * l.shape[1] + kw.kw) * l.shape[2] + cin) * l.shape[3] + co];
                        }
                        if (c.kcat) v *= q.fold[co];
                        v *= wscale;
                        const __half h = __float2half_rn(v);
                        const size_t o = ((it * np + 0) * c.Cout_pad + n) * 64 + j;
                        packed[o] = h;
                        if (np == 2) packed[o + static_cast<size_t>(c.Cout_pad) * 64] = __float2half_rn(v - __half2float(h));
                    }
                }
                if (!c.kcat) n_base += q.cout;
            }
        }
        std::vector<float> sc(c.Cout_pad, 0.f), sh(c.Cout_pad, 0.f);
        int n_base = 0;
        if (c.kind == K_CONVT_FUSED) {
            int tot = 0;
            for (auto& q : pw) tot += q.cout;
            for (int ph = 0; ph < 4; ++ph) {
                int cb = 0;
                for (auto& q : pw) {
                    for (int co = 0; co < q.cout; ++co) {
                        float s1 = 1.f, t1 = q.b[co];
                        if (q.bn) {
                            const float g = q.bn[co], be = q.bn[q.cout + co], mu = q.bn[2 * q.cout + co], var = q.bn[3 * q.cout + co];
                            s1 = g / sqrtf(var + 1e-3f);
                            t1 = be - mu * s1 + q.b[co] * s1;
                        }
                        sc[ph * tot + cb + co] = s1 / wscale;
                        sh[ph * tot + cb + co] = t1;
                    }
                    cb += q.cout;
                }
            }
        }
        for (auto& q : pw) {
            if (c.kind == K_CONVT_FUSED) break;
            for (int co = 0; co < q.cout; ++co) {
                float s = 1.f, t = q.b[co];
                if (q.bn) {
                    const float g = q.bn[co], be = q.bn[q.cout + co], mu = q.bn[2 * q.cout + co], var = q.bn[3 * q.cout + co];
                    s = g / sqrtf(var + 1e-3f);  // keras BatchNormalization epsilon
                    t = be - mu * s + q.b[co] * s;
                }
                if (c.kcat) {               // scales live in the weights; the shifts of the parts add up
                    sc[co] = 1.f / wscale;
                    sh[co] += t;
                } else {
                    sc[n_base + co] = s / wscale;
                    sh[n_base + co] = t;
                }
            }
            if (!c.kcat) n_base += q.cout;
        }
        mc.packed.upload(packed.data(), packed.size());
        mc.scale.upload(sc.data(), sc.size());
        mc.shift.upload(sh.data(), sh.size());
        cuuint64_t dims[4] = {64, (cuuint64_t)c.Cout_pad, (cuuint64_t)np, (cuuint64_t)KI};
        cuuint64_t str[3] = {128, (cuuint64_t)c.Cout_pad * 128, (cuuint64_t)c.Cout_pad * 128 * np};
        cuuint32_t box[4] = {64, (cuuint32_t)c.BN, (cuuint32_t)np, 1};
        encode(&mc.mapB, mc.packed.p, 4, dims, str, box);
        mc.mapBmc[0] = mc.mapBmc[1] = mc.mapB;
        if (c.BN == 128) {
            cuuint32_t b2[4] = {64, 64, 1, 1}, b4[4] = {64, 32, 1, 1};
            encode(&mc.mapBmc[0], mc.packed.p, 4, dims, str, b2);
            encode(&mc.mapBmc[1], mc.packed.p, 4, dims, str, b4);
        }
        if (c.BN >= 128) {
            cuuint32_t boxh[4] = {64, (cuuint32_t)(c.BN / 2), (cuuint32_t)np, 1};
            encode(&mc.mapBh, mc.packed.p, 4, dims, str, boxh);
        } else {
            mc.mapBh = mc.mapB;
        }
    }
    P2P_CUDA(cudaDeviceSynchronize());
}

// ------------------------------------------------------------------------------------------------
// forward
namespace {

template <int BN, int NP>
void launch_conv_persistent(const CUtensorMap* mA, const CUtensorMap& mB, const CUtensorMap* mO, const CUtensorMap& mR,
                            const ConvParams& p, int ctas, cudaStream_t s) {
    using Cfg = ConvCfg<BN, NP>;
    static bool configured = false;
    if (!configured) {
        P2P_CUDA(cudaFuncSetAttribute(conv_tc_persistent_kernel<BN, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES_P));
        configured = true;
    }
    conv_tc_persistent_kernel<BN, NP><<<ctas, 64 + kEpiThreads, Cfg::SMEM_BYTES_P, s>>>(mA[0], mA[1], mA[2], mA[3], mB, mO[0], mO[1], mO[2], mO[3], mR, p);
    P2P_CUDA(cudaGetLastError());
}

template <int BN, int NP, int CL>
void launch_conv_slab(const CUtensorMap* mS, const CUtensorMap& mB, const CUtensorMap& mO, const ConvParams& p, int tiles, int num_sms,
                      cudaStream_t s) {
    using SC = SlabCfg<BN, NP>;
    static bool configured = false;
    if (!configured) {
        P2P_CUDA(cudaFuncSetAttribute(conv_tc_slab_kernel<BN, NP, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC::SMEM_BYTES));
        configured = true;
    }
    const int groups = std::min((tiles + CL - 1) / CL, num_sms / CL);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(groups * CL, 1, 1);
    cfg.blockDim = dim3(64 + kEpiThreads, 1, 1);
    cfg.dynamicSmemBytes = SC::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    P2P_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_slab_kernel<BN, NP, CL>, mS[0], mS[1], mB, mO, p));
}

// CTA-pair kernel: grid = 2 x (clusters that can be resident at once, at most one per tile pair)
template <int BN, int NP>
void launch_conv_pair(const CUtensorMap* mA, const CUtensorMap& mBh, const CUtensorMap* mO, const ConvParams& p, int pairs, cudaStream_t s) {
    using Cfg = PairCfg<BN, NP>;
    static int max_clusters = 0;
    if (!max_clusters) {
        P2P_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel<BN, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2, 1, 1);
        cfg.blockDim = dim3(64 + kEpiThreads, 1, 1);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, conv_tc_pair_kernel<BN, NP>, &cfg) != cudaSuccess || max_clusters < 1) {
            (void)cudaGetLastError();
            int dev = 0, sms = 0;
            P2P_CUDA(cudaGetDevice(&dev));
            P2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            max_clusters = std::max(1, sms / 2);
        }
        if (getenv("P2P_PROF_LAYERS")) fprintf(stderr, "[pair<%d,%d>] resident clusters %d\n", BN, NP, max_clusters);
    }
    const int clusters = std::min(pairs, max_clusters);
    conv_tc_pair_kernel<BN, NP><<<2 * clusters, 64 + kEpiThreads, Cfg::SMEM_BYTES, s>>>(mA[0], mA[1], mA[2], mA[3], mBh, mO[0], mO[1], mO[2], mO[3], p);
    P2P_CUDA(cudaGetLastError());
}

}  // namespace

void Engine::forward(const Model& m, const float* x_dev, int n, float* dec_dev, float* prob_dev, const int* n_active,
                     cudaStream_t s) {
    P2P_CHECK(n >= 1 && n <= cap, "forward: n=%d outside [1,%d]", n, cap);
    P2P_CHECK(m.engine == this, "model was packed for a different engine");
    std::vector<cudaEvent_t> pev;
    std::vector<int> pkind;
    auto mark = [&](int kind) {
        if (!prof) return;
        cudaEvent_t e;
        P2P_CUDA(cudaEventCreate(&e));
        P2P_CUDA(cudaEventRecord(e, s));
        pev.push_back(e);
        pkind.push_back(kind);
    };
    mark(-1);
    for (const Step& st : plan.steps) {
        if (st.kind == S_IM2COL) {
            const long long total = static_cast<long long>(n) * 64 * 64 * (st.d / 8);
            const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 1 << 22));
            im2col_stem_kernel<<<blocks, 256, 0, s>>>(x_dev, tensors[st.a].buf.p, tensors[st.a].plane, n, st.b, st.c, st.d, n_active);
            P2P_CUDA(cudaGetLastError());
            ++launches;
            mark(1);
        } else if (st.kind == S_STEMCVT) {
            const long long total = static_cast<long long>(n) * 128 * 128;
            const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 1 << 22));
            stem_convert_kernel<<<blocks, 256, 0, s>>>(x_dev, tensors[st.a].buf.p, tensors[st.a].plane, n, plan.tensors[st.a].W, st.b, n_active);
            P2P_CUDA(cudaGetLastError());
            ++launches;
            mark(1);
        } else if (st.kind == S_MAXPOOL) {
            const TensorSpec& ti = plan.tensors[st.a];
            const long long total = static_cast<long long>(n) * (ti.H / 2) * (ti.W / 2) * (ti.C / 8);
            const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 1 << 22));
            maxpool3x3s2_kernel<<<blocks, 256, 0, s>>>(tensors[st.a].buf.p, tensors[st.a].plane, tensors[st.b].buf.p,
                                                       tensors[st.b].plane, n, ti.H, ti.W, ti.C, n_active);
            P2P_CUDA(cudaGetLastError());
            ++launches;
            mark(1);
        } else {
            const ConvSpec& c = plan.convs[st.a];
            const ConvRt& rt = conv_rt[st.a];
            const ModelConv& mc = m.convs[st.a];
            ConvParams p;
            memset(&p, 0, sizeof(p));
            p.N = n; p.H = c.H; p.W = c.W; p.tw = c.tw; p.th = c.th; p.nb = c.nb;
            p.tiles_x = (c.W + c.tw - 1) / c.tw;
            p.tiles_y = (c.H + c.th - 1) / c.th;
            for (int i = 0; i < 5; ++i) p.kstart[i] = c.kstart[i];
            p.kit = rt.kit.p;
            p.sy = c.sy; p.sx = c.sx;
            for (int i = 0; i < 4; ++i) { p.oy_off[i] = c.oy_off[i]; p.ox_off[i] = c.ox_off[i]; }
            p.OH = c.H * c.sy; p.OW = c.W * c.sx;
            p.Cout = c.Cout; p.c_off = c.c_off;
            p.scale = mc.scale.p; p.shift = mc.shift.p; p.act = c.act;
            if (c.act == ACT_HEADS) {
                p.out_dec = dec_dev; p.out_prob = prob_dev;
            } else {
                const TensorSpec& to = plan.tensors[c.out_tensor];
                p.out_hi = tensors[c.out_tensor].buf.p; p.out_plane = tensors[c.out_tensor].plane; p.Ctot = to.C;
                if (c.kind == K_DENSE) { p.OH = 1; p.OW = 1; p.Ctot = to.H * to.W * to.C; }
            }
            if (c.res_tensor >= 0) {
                p.res_hi = tensors[c.res_tensor].buf.p; p.res_plane = tensors[c.res_tensor].plane;
                p.res_Ctot = plan.tensors[c.res_tensor].C;
            }
            p.n_active = n_active;
            P2P_CHECK(c.kit.size() <= sizeof(p.ksteps_tab), "conv %s: %zu k-iterations exceed the parameter table", c.name.c_str(), c.kit.size());
            for (size_t i = 0; i < c.kit.size(); ++i) p.ksteps_tab[i] = static_cast<uint8_t>(c.kit[i].x >> 8);
            int max_nk = 0;          // most k-iterations any tile of this launch runs
            {
                int max_steps = 0;   // longest accumulation chain of this launch, in k16 steps
                const int nz = c.splitk > 1 ? 1 : c.phases;
                for (int z = 0; z < nz; ++z) {
                    int steps = 0;
                    const int kb = c.kstart[z], ke = c.splitk > 1 ? std::min<int>(c.splitk_chunk, static_cast<int>(c.kit.size())) : c.kstart[z + 1];
                    for (int it = kb; it < ke; ++it) steps += c.kit[it].x >> 8;
                    max_steps = std::max(max_steps, steps);
                    max_nk = std::max(max_nk, ke - kb);
                }
                p.single_acc = (np == 2 && max_steps <= single_acc_steps) ? 1 : 0;
            }
            if (const char* e = getenv("P2P_DBG")) p.dbg = atoi(e);
            p.a_sw64 = (c.kind == K_STEM && stem_sw64()) ? 1 : 0;
            p.Cout_pad = c.Cout_pad;
            if (c.kind == K_CONVT_FUSED && c.act != ACT_HEADS) p.fused_cout = c.Cout / 4;
            if (c.splitk > 1) {
                p.splitk_chunk = c.splitk_chunk;
                p.out_partial = partial.p;
                p.partial_stride = static_cast<long long>(cap) * c.Cout_pad;
            }
            const int tiles_n = (n + c.nb - 1) / c.nb;
            dim3 grid(p.tiles_x * p.tiles_y * tiles_n, c.Cout_pad / c.BN, c.splitk > 1 ? c.splitk : c.phases);
            p.grid_m = grid.x; p.grid_n = grid.y; p.grid_z = grid.z;
            // N = 128 tiles stay on the single-CTA kernel: a cta_group::2 MMA seems to cost >= ~100 cycles whatever its width,
            // so three N = 128 pair MMAs per k-step ran 40-50 % slower than the widened two of the single-CTA kernel
            // (deconv3 1627 -> 2452 us); at N = 256 the pair wins (conv4 499 -> 438 us).  P2P_PAIR_BN128=1 = experiment.
            static const int pair_min_bn = getenv("P2P_PAIR_BN128") && atoi(getenv("P2P_PAIR_BN128")) ? 128 : 256;
            const bool use_pair = pair && c.BN >= pair_min_bn && c.splitk <= 1 && c.kind != K_DENSE && c.act != ACT_HEADS &&
                                  c.res_tensor < 0 && rt.has_out && tma_store && grid.x >= 2;
            if (slab && c.slab && c.act == ACT_HEADS) {
                p.tma_store = 0;
                p.nst = 0; p.epi_bufs = 1; p.res_tma = 0;
                p.slab_ksize = c.slab_ks;
                p.kit = rt.slabs.p;
                p.kstart[0] = 0;
                for (int i = 1; i < 5; ++i) p.kstart[i] = static_cast<int>(c.slabs.size());
                for (size_t i = 0; i < c.slabs.size(); ++i) p.ksteps_tab[i] = static_cast<uint8_t>(c.slabs[i].x >> 8);
                const int tiles = static_cast<int>(grid.x * grid.y);
                constexpr int res_kiters = SlabCfg<16, 2>::RES_KITERS;
                P2P_CHECK(static_cast<int>(c.slabs.size()) * c.slab_ks * c.slab_ks <= res_kiters,
                          "heads: %zu weight tiles do not fit the resident buffer", c.slabs.size() * c.slab_ks * c.slab_ks);
                if (np == 2) launch_conv_slab<16, 2, 1>(rt.mapSlab, mc.mapB, rt.mapSlab[0], p, tiles, num_sms, s);   // (no output map: fp32 stores)
                else launch_conv_slab<16, 1, 1>(rt.mapSlab, mc.mapB, rt.mapSlab[0], p, tiles, num_sms, s);
            } else if (slab && c.slab && rt.has_out && tma_store) {
                p.tma_store = 1;
                p.nst = 0; p.epi_bufs = 1; p.res_tma = 0;
                p.slab_ksize = c.slab_ks;
                p.kit = rt.slabs.p;
                p.kstart[0] = 0;
                for (int i = 1; i < 5; ++i) p.kstart[i] = static_cast<int>(c.slabs.size());
                for (size_t i = 0; i < c.slabs.size(); ++i) p.ksteps_tab[i] = static_cast<uint8_t>(c.slabs[i].x >> 8);
                const int tiles = static_cast<int>(grid.x * grid.y);
                const int cl = (tiles >= 2 * num_sms && c.BN == 128) ? slab_cluster : 1;  // small launches: no point in pairing up CTAs
                if (c.BN == 64) {
                    if (np == 2) launch_conv_slab<64, 2, 1>(rt.mapSlab, mc.mapB, rt.mapOut[0], p, tiles, num_sms, s);
                    else launch_conv_slab<64, 1, 1>(rt.mapSlab, mc.mapB, rt.mapOut[0], p, tiles, num_sms, s);
                } else if (np == 2) {
                    if (cl == 4) launch_conv_slab<128, 2, 4>(rt.mapSlab, mc.mapBmc[1], rt.mapOut[0], p, tiles, num_sms, s);
                    else if (cl == 2) launch_conv_slab<128, 2, 2>(rt.mapSlab, mc.mapBmc[0], rt.mapOut[0], p, tiles, num_sms, s);
                    else launch_conv_slab<128, 2, 1>(rt.mapSlab, mc.mapB, rt.mapOut[0], p, tiles, num_sms, s);
                } else {
                    if (cl == 4) launch_conv_slab<128, 1, 4>(rt.mapSlab, mc.mapBmc[1], rt.mapOut[0], p, tiles, num_sms, s);
                    else if (cl == 2) launch_conv_slab<128, 1, 2>(rt.mapSlab, mc.mapBmc[0], rt.mapOut[0], p, tiles, num_sms, s);
                    else launch_conv_slab<128, 1, 1>(rt.mapSlab, mc.mapB, rt.mapOut[0], p, tiles, num_sms, s);
                }
            } else if (use_pair) {
                p.tma_store = 1;
                p.nst = 0; p.epi_bufs = 1; p.res_tma = 0;
                const int pairs = static_cast<int>((grid.x + 1) / 2 * grid.y * grid.z);
                if (c.BN == 256) {
                    if (np == 2) launch_conv_pair<256, 2>(rt.mapA, mc.mapBh, rt.mapOut, p, pairs, s);
                    else launch_conv_pair<256, 1>(rt.mapA, mc.mapBh, rt.mapOut, p, pairs, s);
                } else {
                    if (np == 2) launch_conv_pair<128, 2>(rt.mapA, mc.mapBh, rt.mapOut, p, pairs, s);
                    else launch_conv_pair<128, 1>(rt.mapA, mc.mapBh, rt.mapOut, p, pairs, s);
                }
            } else {
                p.tma_store = (tma_store && rt.has_out) ? 1 : 0;
                {
                    // epilogue-bound layers (few k-iterations per tile) run on fewer operand stages and use the freed
                    // shared memory as extra output staging tiles, so TMA stores overlap the next slice's arithmetic
                    const int stage_bytes = np * 128 * 128 + np * c.BN * 128, epi_bytes = np * 128 * 128;
                    const int stages = std::min(6, 200 * 1024 / stage_bytes);
                    int nst = stages;
                    if (max_nk <= epi_nk && stages > 2) nst = 2;
                    else if (c.BN == 64 && stages > 3) nst = stages - 1;
                    p.nst = nst;
                    p.epi_bufs = (p.tma_store && c.BN >= 64) ? std::min(3, 1 + (stages - nst) * stage_bytes / epi_bytes) : 1;
                    if (res_tma && rt.has_res && p.tma_store && stages > 2 && (stages - 2) * stage_bytes >= 2 * epi_bytes) {
                        p.res_tma = 1;  // two residual tiles live in the operand stages above the second
                        p.nst = 2;
                        p.epi_bufs = 1;
                    }
                    if (const char* e = getenv("P2P_NST")) { p.nst = atoi(e); p.epi_bufs = 1; p.res_tma = 0; }
                }
                const int ctas = std::min<long long>(static_cast<long long>(grid.x) * grid.y * grid.z, num_sms);
                if (c.BN == 256) {
                    if (np == 2) launch_conv_persistent<256, 2>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                    else launch_conv_persistent<256, 1>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                } else if (np == 2) {
                    if (c.BN == 128) launch_conv_persistent<128, 2>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                    else if (c.BN == 64) launch_conv_persistent<64, 2>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                    else launch_conv_persistent<16, 2>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                } else {
                    if (c.BN == 128) launch_conv_persistent<128, 1>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                    else if (c.BN == 64) launch_conv_persistent<64, 1>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                    else launch_conv_persistent<16, 1>(rt.mapA, mc.mapB, rt.mapOut, rt.mapRes, p, ctas, s);
                }
            }
            ++launches;
            mark(0);
            if (c.splitk > 1) {
                const TensorSpec& to = plan.tensors[c.out_tensor];
                const long long total = static_cast<long long>(n) * c.Cout;
                const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, num_sms * 8));
                splitk_reduce_kernel<<<blocks, 256, 0, s>>>(partial.p, p.partial_stride, c.splitk, n, c.Cout, c.Cout_pad, mc.scale.p,
                                                            mc.shift.p, c.act, p.out_hi, p.out_plane, to.H * to.W * to.C, n_active);
                P2P_CUDA(cudaGetLastError());
                ++launches;
                mark(1);
            }
        }
    }
    if (prof) {
        P2P_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 1; i < pev.size(); ++i) {
            float ms = 0;
            P2P_CUDA(cudaEventElapsedTime(&ms, pev[i - 1], pev[i]));
            prof[pkind[i]] += ms;
            prof[2 + pkind[i]] += 1;
            if (prof_layers) fprintf(stderr, "%zu:%.0f ", i - 1, ms * 1000.f);  // per-launch microseconds (P2P_PROF_LAYERS=1)
        }
        if (prof_layers) fprintf(stderr, "\n");
        for (auto e : pev) cudaEventDestroy(e);
    }
}

void Engine::predict_host(const Model& m, const float* xh, int n, float* dech, float* probh) {
    P2P_CHECK(n >= 0, "negative batch");
    for (int b = 0; b < n; b += cap) {
        const int nb = std::min(cap, n - b);
        const size_t px = static_cast<size_t>(nb) * 128 * 128;
        P2P_CUDA(cudaMemcpyAsync(x.p, xh + static_cast<size_t>(b) * 128 * 128 * 3, px * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
        forward(m, x.p, nb, dec.p, prob.p, nullptr, stream);
        P2P_CUDA(cudaMemcpyAsync(dech + static_cast<size_t>(b) * 128 * 128 * 3, dec.p, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
        P2P_CUDA(cudaMemcpyAsync(probh + static_cast<size_t>(b) * 128 * 128, prob.p, px * sizeof(float), cudaMemcpyDeviceToHost, stream));
        P2P_CUDA(cudaStreamSynchronize(stream));
    }
}

void Engine::read_tensor(const std::string& name, int n, float* out) {
    const int id = plan.tensor_id(name);
    const TensorSpec& t = plan.tensors[id];
    P2P_CHECK(n >= 1 && n <= cap, "read_tensor: n=%d outside [1,%d]", n, cap);
    const size_t cnt = static_cast<size_t>(n) * t.H * t.W * t.C;
    std::vector<__half> hi(cnt), lo(cnt);
    P2P_CUDA(cudaStreamSynchronize(stream));
    P2P_CUDA(cudaMemcpy(hi.data(), tensors[id].buf.p, cnt * 2, cudaMemcpyDeviceToHost));
    if (tensors[id].plane) P2P_CUDA(cudaMemcpy(lo.data(), tensors[id].buf.p + tensors[id].plane, cnt * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < cnt; ++i) out[i] = __half2float(hi[i]) + (tensors[id].plane ? __half2float(lo[i]) : 0.f);
}

}  // namespace p2p
