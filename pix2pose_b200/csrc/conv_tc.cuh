// Tap-list implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One kernel serves every contraction of the Pix2Pose generator
// (reference: pix2pose_model/ae_model.py:175-240, resnet50_mod.py:40-118):
//   * k x k stride-1 'same' convs (1x1, 3x3, 5x5), input = concat of up to two channel-sliced tensors
//   * 5x5 / 1x1 stride-2 convs via four parity-phase tensor maps (TF 'same': pad 1 before / 2 after)
//   * 5x5 stride-2 transposed convs as four output-phase stride-1 convs (2x2 / 2x3 / 3x2 / 3x3 taps)
//   * the two Dense layers (a 1x1 "image" with C = 32768 / 256) and the im2col'd stems
//
// GEMM view per CTA: D[128 pixels, BN couts] = sum over k-iterations (source, tap, 64-channel chunk)
// of A[128 px, 64 ch] * B[BN, 64 ch]^T.  A tiles are fetched by TMA straight out of the NHWC
// activation tensor as a (64ch, tw, th, nb[, plane]) box shifted by the tap offset; out-of-image
// elements are zero-filled by TMA, which implements the zero padding.  B tiles come from weights
// pre-packed [k-iter][plane][Cout][64].  Both land in shared memory in the K-major 128-byte-swizzle
// layout tcgen05.mma consumes directly; accumulators live in TMEM.
//
// Precision: NP == 2 ("fp16x3"): every operand is an fp16 (hi, lo) pair and each k-step issues
// hi*hi into one fp32 accumulator and lo*hi + hi*lo into a second one (~22-bit operands; the generator
// output matches the fp32 reference to ~2e-4).  NP == 1: plain fp16 operands (one MMA per k-step, ~1e-2 max abs error
// through the whole network on random weights).
//
// This header holds what all kernels share: ConvParams, the operand-stage geometry (ConvCfg) and the split-K reduce
// kernel.  The kernels themselves: conv_tc_persistent.cuh (one CTA per SM looping over tiles; default),
// conv_tc_pair.cuh (CTA pairs, cta_group::2), conv_tc_slab.cuh (slab reuse for k x k stride-1 convs).
#pragma once
#include "sm100_ptx.cuh"

namespace p2p {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_HEADS = 3 };

struct ConvParams {
    // M tiling: grid of H x W positions per image, tile = tw x th x nb (= 128 rows)
    int N, H, W;
    int tw, th, nb;
    int tiles_x, tiles_y;
    int grid_m, grid_n, grid_z;  // tile grid (persistent kernel: tiles are enumerated m fastest, then n, then z)
    // K loop: k-iterations [kstart[z], kstart[z+1]) for phase z = blockIdx.z
    int kstart[5];
    const int4* kit;  // {map index | (k16 steps << 8), dy, dx, c0} per k-iteration
    // output tensor (NHWC, hi plane then lo plane `out_plane` elements later)
    __half* out_hi;
    long long out_plane;  // 0 => no lo plane (NP == 1 networks)
    int OH, OW, Ctot, c_off, Cout;
    int sy, sx;
    int oy_off[4], ox_off[4];
    const float* scale;  // [Cout_pad] folded BN scale / weight pre-scale
    const float* shift;  // [Cout_pad] folded BN shift + bias
    int act;
    const __half* res_hi;  // optional residual (same N,OH,OW; res_Ctot channels), added before the activation
    long long res_plane;
    int res_Ctot;
    float* out_dec;   // ACT_HEADS: (N,OH,OW,3) fp32 tanh
    float* out_prob;  // ACT_HEADS: (N,OH,OW,1) fp32 sigmoid
    const int* n_active;  // optional device-side count of live images (tiles past it exit)
    // split-K (Dense layers with a tiny M x N grid): blockIdx.z = K slice of `splitk_chunk` k-iterations; the raw
    // fp32 accumulators go to out_partial[z][row][Cout_pad] and splitk_reduce_kernel finishes the layer.
    // Phase-fused transposed conv with a regular epilogue: accumulator column c = phase * fused_cout + channel,
    // phase (a,b) -> output pixel (2y+a, 2x+b).  0 = off.
    int fused_cout;
    int tma_store;   // persistent kernel: stage 64-channel output slices in shared memory and write them with TMA stores
                     // (the per-lane 16-byte stores of a row-per-thread epilogue are uncoalesced: 32 lines per instruction)
    int single_acc;  // short K loops (<= 40 k16 steps): one TMEM accumulator instead of two (the round-toward-zero bias the
                     // split guards against grows with the chain length; the epilogue drain is 2x cheaper)
    int res_tma;     // persistent kernel: the residual arrives by TMA (tensor map mR) in two shared-memory tiles, two slices ahead
    int nst;         // persistent kernel: operand stages to use (0 = all).  Short-K layers run on fewer stages and hand the
    int epi_bufs;    // top ones to the epilogue as extra output staging tiles (epi_bufs = 1..3 tiles of EPI_BYTES)
    int dbg;         // bottleneck experiments on the persistent kernel (WRONG RESULTS): bit0 skip A loads, bit1 skip B loads,
                     // bit2 MMA issuer does not wait for operands
    int a_sw64;      // persistent kernel: A tiles are 128 rows x 64 bytes in the 64-byte swizzle (stem: K <= 32 per k-iteration)
    int slab_ksize;  // slab kernel (conv_tc_slab.cuh): kernel size; p.kit then holds the slab table, kstart[1] = #entries
    int splitk_chunk;
    float* out_partial;
    long long partial_stride;  // elements between K slices
    int Cout_pad;
    // k16 steps of every k-iteration (same numbers as kit[i].x >> 8), in the kernel parameters so that the MMA issuer
    // reads them through the constant bank into uniform registers
    uint8_t ksteps_tab[512];
};

template <int BN, int NP>
struct ConvCfg {
    static constexpr int A_BYTES = NP * 128 * 128;
    static constexpr int B_BYTES = NP * BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
    static constexpr int EPI_BYTES = NP * 128 * 128;  // output staging for the TMA-store epilogue (persistent kernel)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int CONST_BYTES = (BN >= 64 && BN <= 128) ? 2 * BN * 4 : 0;  // folded BN scale / shift of a tile column
    static constexpr int SMEM_BYTES_P = SMEM_BYTES + (BN >= 64 ? EPI_BYTES : 0) + CONST_BYTES;
    // TMEM plan (accumulator sets, columns): PersCfg in conv_tc_persistent.cuh -- fp16x3 keeps TWO accumulators per tile
    // (all hi*hi products in one chain, the 2^-11-times-smaller cross terms in the other; the tensor core rounds the fp32
    // accumulator toward zero after every MMA, and keeping the small terms out of the long chain is what holds the 1e-3
    // tolerance), double-buffered whenever two sets fit into the 512 columns.
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LRELU) return v > 0.f ? v : 0.3f * v;
    return v;
}

// Finishes a split-K layer: out[row][c] = act((sum_z partial[z][row][c]) * scale[c] + shift[c]) -> fp16 hi/lo planes.
static __global__ void splitk_reduce_kernel(const float* __restrict__ partial, long long partial_stride, int nsplit, int rows,
                                     int Cout, int Cout_pad, const float* __restrict__ scale, const float* __restrict__ shift,
                                     int act, __half* __restrict__ out_hi, long long out_plane, int Ctot,
                                     const int* __restrict__ n_active) {
    int n_limit = rows;
    if (n_active) n_limit = min(n_limit, *n_active);
    const long long total = static_cast<long long>(n_limit) * Cout;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cout);
        const long long r = i / Cout;
        float acc = 0.f;
        for (int z = 0; z < nsplit; ++z) acc += partial[z * partial_stride + r * Cout_pad + c];
        const float v = act_apply(acc * scale[c] + shift[c], act);
        const __half h = __float2half_rn(v);
        out_hi[r * Ctot + c] = h;
        if (out_plane) out_hi[r * Ctot + c + out_plane] = __float2half_rn(v - __half2float(h));
    }
}

}  // namespace p2p
