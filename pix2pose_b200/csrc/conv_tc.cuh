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
// hi*hi + lo*hi + hi*lo into the same fp32 accumulator (~22-bit operands; matches the fp32
// reference to ~1e-5).  NP == 1: plain fp16 operands (one MMA per k-step, ~1e-2 max abs error
// through the whole network on random weights).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
// (TMEM -> registers -> folded BN / bias / residual / activation -> NHWC global, hi/lo planes).
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
    int single_acc;  // short K loops (<= 40 k16 steps): one TMEM accumulator instead of three (the round-toward-zero bias the
                     // split guards against grows with the chain length; the epilogue drain is 3x cheaper)
    int res_tma;     // persistent kernel: the residual arrives by TMA (tensor map mR) in two shared-memory tiles, two slices ahead
    int nst;         // persistent kernel: operand stages to use (0 = all).  Short-K layers run on fewer stages and hand the
    int epi_bufs;    // top ones to the epilogue as extra output staging tiles (epi_bufs = 1..3 tiles of EPI_BYTES)
    int dbg;         // bottleneck experiments on the persistent kernel (WRONG RESULTS): bit0 skip A loads, bit1 skip B loads,
                     // bit2 MMA issuer does not wait for operands
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
    // fp16x3 keeps THREE accumulators in TMEM: two for the hi*hi products (even / odd k-steps) and one
    // for the 2^-11-times-smaller cross terms.  The tensor core rounds the fp32 accumulator toward zero
    // after every MMA; splitting the chains divides that bias (measured: resnet50 decode max error
    // 1.0e-3 with one accumulator).  The epilogue adds the three.
    // BN = 256 (N = 256 MMAs halve the shared-memory operand reads per FLOP -- the kernel is smem-bandwidth-bound at
    // N = 128: 8 KB of operands per 64-cycle MMA = the full 128 B/cycle, before TMA writes) only has room for two:
    // all hi*hi products in one chain, cross terms in the other.
    static constexpr int NACC = NP == 2 ? (BN == 256 ? 2 : 3) : 1;
    static constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;  // TMEM columns between accumulators
    static constexpr uint32_t TMEM_COLS = NACC * ACC_STRIDE <= 32 ? 32 : (NACC * ACC_STRIDE <= 64 ? 64 : (NACC * ACC_STRIDE <= 128 ? 128 : (NACC * ACC_STRIDE <= 256 ? 256 : 512)));
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LRELU) return v > 0.f ? v : 0.3f * v;
    return v;
}

template <int BN, int NP>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
               const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
               const __grid_constant__ CUtensorMap mB, const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BN, NP>;
    constexpr int STAGES = Cfg::STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int z = blockIdx.z;

    // ---- tile coordinates (uniform over the CTA)
    int n_limit = p.N;
    if (p.n_active != nullptr) {
        const int na = *p.n_active;
        n_limit = na < n_limit ? na : n_limit;
    }
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int tn = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - tn * tiles_per_img;
    const int ty = trem / p.tiles_x;
    const int tx = trem - ty * p.tiles_x;
    const int n0 = tn * p.nb, y0 = ty * p.th, x0 = tx * p.tw;
    if (n0 >= n_limit) return;
    const int nt0 = blockIdx.y * BN;
    const bool splitk = p.splitk_chunk > 0;
    const int zi = splitk ? 0 : z;  // index into the per-phase tables
    const int kbeg = splitk ? z * p.splitk_chunk : p.kstart[z];
    const int nk = splitk ? min(p.splitk_chunk, p.kstart[1] - kbeg) : p.kstart[z + 1] - kbeg;

    // ---- shared memory carve-up (1024-B aligned operand tiles for the 128-B swizzle)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int4 k = __ldg(&p.kit[kbeg + it]);
                uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                const int mi = k.x & 0xff;
                const CUtensorMap* mA = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                tma_load_5d(mA, &full_bar[s], sA, k.w, x0 + k.z, y0 + k.y, n0, 0);
                tma_load_4d(&mB, &full_bar[s], sB, 0, nt0, 0, kbeg + it);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN < 16 ? 16 : BN);
            int g = 0;  // k16 steps issued so far
            for (int it = 0; it < nk; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int ksteps = __ldg(&p.kit[kbeg + it].x) >> 8;  // 64-channel chunk: 4; a 32-channel skip slice: 2
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t aA = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t aB = aA + Cfg::A_BYTES;
#pragma unroll 1
                for (int kk = 0; kk < ksteps; ++kk, ++g) {
                    const uint64_t a_hi = umma_desc_sw128(aA + kk * 32);
                    const uint64_t b_hi = umma_desc_sw128(aB + kk * 32);
                    if (NP == 2) {
                        // k16 step parity: even -> accumulator 0, odd -> accumulator 1
                        umma_f16(tmem_base + (g & 1) * Cfg::ACC_STRIDE, a_hi, b_hi, idesc, g >= 2 ? 1u : 0u);
                        const uint64_t a_lo = umma_desc_sw128(aA + 128 * 128 + kk * 32);
                        const uint64_t b_lo = umma_desc_sw128(aB + BN * 128 + kk * 32);
                        umma_f16(tmem_base + 2 * Cfg::ACC_STRIDE, a_lo, b_hi, idesc, g > 0 ? 1u : 0u);
                        umma_f16(tmem_base + 2 * Cfg::ACC_STRIDE, a_hi, b_lo, idesc, 1u);
                    } else {
                        umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above retire
            }
            umma_commit(tmem_full_bar);  // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;
        const int wl = r % p.tw;
        const int hl = (r / p.tw) % p.th;
        const int nl = r / (p.tw * p.th);
        const int y = y0 + hl, x = x0 + wl, n = n0 + nl;
        const bool valid = (y < p.H) && (x < p.W) && (n < n_limit);
        const int oy = y * p.sy + p.oy_off[zi], ox = x * p.sx + p.ox_off[zi];
        const long long pix = (static_cast<long long>(n) * p.OH + oy) * p.OW + ox;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

        if (p.act == ACT_HEADS) {
            // Phase-fused transposed-conv heads: the 16 accumulator columns are (output phase a,b) x (x,y,z,prob);
            // this row's input pixel (y,x) produces the 2x2 output pixels (2y+a, 2x+b).
            uint32_t v[16];
            tmem_ld_32x16(taddr, v);
            tmem_ld_wait();
            if (Cfg::NACC == 3) {
                uint32_t v1[16], v2[16];
                tmem_ld_32x16(taddr + Cfg::ACC_STRIDE, v1);
                tmem_ld_32x16(taddr + 2 * Cfg::ACC_STRIDE, v2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = __float_as_uint((__uint_as_float(v[j]) + __uint_as_float(v1[j])) + __uint_as_float(v2[j]));
            }
            if (valid) {
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const long long opix = (static_cast<long long>(n) * p.OH + (2 * y + (ph >> 1))) * p.OW + (2 * x + (ph & 1));
                    float* d = p.out_dec + opix * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        d[c] = tanhf(__uint_as_float(v[ph * 4 + c]) * __ldg(&p.scale[ph * 4 + c]) + __ldg(&p.shift[ph * 4 + c]));
                    const float e = __uint_as_float(v[ph * 4 + 3]) * __ldg(&p.scale[ph * 4 + 3]) + __ldg(&p.shift[ph * 4 + 3]);
                    p.out_prob[opix] = 1.f / (1.f + expf(-e));
                }
            }
        } else {
            constexpr int NCH = BN >= 32 ? BN / 32 : 1;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (Cfg::NACC == 3) {
                    uint32_t v1[32];
                    tmem_ld_32x32(taddr + Cfg::ACC_STRIDE + ch * 32, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v1[j]));
                    tmem_ld_32x32(taddr + 2 * Cfg::ACC_STRIDE + ch * 32, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v1[j]));
                }
                const int c0 = nt0 + ch * 32;
                if (splitk) {
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(p.out_partial + z * p.partial_stride + pix * p.Cout_pad + c0);
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            dst[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                                 __uint_as_float(v[4 * g + 3]));
                    }
                } else if (valid && c0 < p.Cout) {
                    long long opix = pix;
                    int cch = c0;
                    if (p.fused_cout) {
                        const int phs = c0 / p.fused_cout;
                        cch = c0 - phs * p.fused_cout;
                        opix = (static_cast<long long>(n) * p.OH + (2 * y + (phs >> 1))) * p.OW + (2 * x + (phs & 1));
                    }
                    __half* o_hi = p.out_hi + opix * p.Ctot + p.c_off + cch;
                    const __half* r_hi = p.res_hi ? p.res_hi + opix * p.res_Ctot + cch : nullptr;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {  // 8 channels per 16-byte store
                        uint4 rh = make_uint4(0, 0, 0, 0), rl = make_uint4(0, 0, 0, 0);
                        if (r_hi) {
                            rh = __ldg(reinterpret_cast<const uint4*>(r_hi + g * 8));
                            if (p.res_plane) rl = __ldg(reinterpret_cast<const uint4*>(r_hi + p.res_plane + g * 8));
                        }
                        const __half* rhh = reinterpret_cast<const __half*>(&rh);
                        const __half* rlh = reinterpret_cast<const __half*>(&rl);
                        uint4 oh, ol;
                        __half* ohh = reinterpret_cast<__half*>(&oh);
                        __half* olh = reinterpret_cast<__half*>(&ol);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = c0 + g * 8 + j;
                            float val = __uint_as_float(v[g * 8 + j]) * __ldg(&p.scale[c]) + __ldg(&p.shift[c]);
                            if (r_hi) val += __half2float(rhh[j]) + __half2float(rlh[j]);
                            val = act_apply(val, p.act);
                            const __half h = __float2half_rn(val);
                            ohh[j] = h;
                            olh[j] = __float2half_rn(val - __half2float(h));
                        }
                        *reinterpret_cast<uint4*>(o_hi + g * 8) = oh;
                        if (p.out_plane) *reinterpret_cast<uint4*>(o_hi + p.out_plane + g * 8) = ol;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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
