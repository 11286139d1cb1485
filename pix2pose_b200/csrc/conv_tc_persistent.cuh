// Persistent variant of the tap-list implicit-GEMM convolution (see conv_tc.cuh for the GEMM view,
// operand layouts and precision scheme; ConvParams / ConvCfg are shared).
//
// One CTA per SM loops over output tiles (static stride scheduler).  What this buys over the
// one-tile-per-CTA kernel:
//   * barrier init, TMEM allocation and descriptor prefetch happen once per CTA, not once per tile;
//   * the TMA producer runs ahead across tile boundaries, so the smem pipeline never drains;
//   * the epilogue drains TMEM into registers (summing the three fp16x3 accumulators on the way),
//     hands TMEM back to the MMA warp, and only then does the BN / activation / split / store work,
//     which therefore overlaps the next tile's MMAs.
#pragma once
#include "conv_tc.cuh"

namespace p2p {

// TMEM plan and MMA schedule of the persistent kernel.
//
// fp16x3 keeps the hi*hi products and the 2^-11-times-smaller cross terms (lo*hi + hi*lo) in two separate accumulators
// (the tensor core rounds the fp32 accumulator toward zero after every MMA; keeping the small terms out of the big chain
// is what holds the 1e-3 tolerance, see ConvCfg).  For tiles of up to 128 columns the two accumulators sit side by side
// in TMEM, [hh: BN columns][cross: BN columns], and the weight tile sits in shared memory as [W_hi: BN rows][W_lo: BN
// rows], so ONE MMA of width 2 * BN computes A_hi * [W_hi; W_lo]^T = (hi*hi | hi*lo) into both, and a second MMA of
// width BN adds A_lo * W_hi^T to the cross half: two instructions and two A-operand reads per k16 step instead of
// three (narrow tiles are instruction-bound: the heads layer, N = 16, went from 521 to 3xx us per 256 crops).
// BN = 256 cannot be widened (N <= 256) and issues the three MMAs separately.
// Whenever two accumulator sets fit into the 512 TMEM columns they are double-buffered: the MMA warp starts tile i + 1
// in the other set while the epilogue warps drain tile i.
template <int BN, int NP>
struct PersCfg {
    static constexpr bool CONCAT = NP == 2 && BN <= 128;        // (hi*hi | hi*lo) in one MMA of width 2 * BN
    static constexpr int ACC_STRIDE = BN;                        // TMEM columns between the hh and the cross accumulator
    static constexpr int NACC = NP == 2 ? 2 : 1;
    static constexpr int BUF_COLS = NACC * ACC_STRIDE < 32 ? 32 : NACC * ACC_STRIDE;  // TMEM columns of one accumulator set
    static constexpr int NBUF = 2 * BUF_COLS <= 512 ? 2 : 1;     // accumulator sets (BN = 256 fp16x3 fills TMEM with one)
    static constexpr uint32_t TMEM_COLS = NBUF * BUF_COLS <= 32 ? 32 : (NBUF * BUF_COLS <= 64 ? 64 : (NBUF * BUF_COLS <= 128 ? 128 : (NBUF * BUF_COLS <= 256 ? 256 : 512)));
};

struct TileCoord {
    int n0, y0, x0, nt0, z, kbeg, nk;
    bool live;
};

// tile (m, ny, z) of the launch grid -> coordinates
template <int BN>
__device__ __forceinline__ TileCoord decode_mnz(const ConvParams& p, int m, int ny, int z, int n_limit) {
    TileCoord c;
    c.z = z;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int tn = m / tiles_per_img;
    const int trem = m - tn * tiles_per_img;
    const int ty = trem / p.tiles_x;
    const int tx = trem - ty * p.tiles_x;
    c.n0 = tn * p.nb; c.y0 = ty * p.th; c.x0 = tx * p.tw;
    c.nt0 = ny * BN;
    c.live = c.n0 < n_limit;
    if (p.splitk_chunk > 0) {
        c.kbeg = c.z * p.splitk_chunk;
        c.nk = min(p.splitk_chunk, p.kstart[1] - c.kbeg);
    } else {
        c.kbeg = p.kstart[c.z];
        c.nk = p.kstart[c.z + 1] - c.kbeg;
    }
    return c;
}

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t, int n_limit) {
    const int m = t % p.grid_m;
    const int r = t / p.grid_m;
    return decode_mnz<BN>(p, m, r % p.grid_n, r / p.grid_n, n_limit);
}

// Epilogue of one tile.  kEpiWarps = 8 warps (256 threads): warp w may only touch TMEM lanes (w % 4) * 32 .. + 31, so
// two warps share each 32-row quarter and split the COLUMNS: of every 64-channel slice, group h (= warps 2-5 / 6-9)
// takes the 32 columns h*32 .. h*32+31.  Per tile each thread (a) drains its columns of the (up to three) accumulators
// into registers, (b) releases TMEM to the MMA warp, (c) applies folded BN / bias / residual / activation and the hi/lo
// split, (d) stages the 64-channel slice in a swizzled shared-memory tile that one thread writes out with a TMA store
// (direct global stores for the split-K, heads and BN < 64 paths).  Eight warps, not four: the epilogue is
// latency-bound (residual / scale loads) and with one warp per scheduler nothing hides that latency.
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;

__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Geometry of 64-channel output slice `s` of a tile for this thread: first channel of the slice (c0s), output pixel and
// channel inside the output tensor (phase-fused transposed convs scatter slice -> output phase).
struct SliceGeom {
    int c0s, phs, cch;
    long long opix;
};
__device__ __forceinline__ SliceGeom slice_geom(const ConvParams& p, const TileCoord& tc, int s, int h, int n, int y, int x, long long pix) {
    SliceGeom g;
    g.c0s = tc.nt0 + s * 64;
    g.phs = 0;
    g.cch = g.c0s + h * 32;
    g.opix = pix;
    if (p.fused_cout) {
        g.phs = g.c0s / p.fused_cout;
        g.cch -= g.phs * p.fused_cout;
        g.opix = (static_cast<long long>(n) * p.OH + (2 * y + (g.phs >> 1))) * p.OW + (2 * x + (g.phs & 1));
    }
    return g;
}

// Shared-memory resources and running state of the epilogue warps (all values are uniform over the 256 threads).
//   stage0     output staging tile 0 (tiles 1.. lie EPI_BYTES below each other, in operand stages the layer does not use)
//   res0       residual tiles (p.res_tma): two tiles of EPI_BYTES; the TMA producer warp fills tile (j & 1) with slice j's
//              residual as soon as slice j - 2 has released it (res_bar[0..1] = full, res_bar[2..3] = empty)
//   s_scale/s_shift  folded BN scale / shift of the current tile column (BN <= 128)
struct EpiState {
    uint8_t* stage0;
    uint8_t* res0;
    uint64_t* res_bar;
    float* s_scale;
    float* s_shift;
    int epi_buf;      // staging tile the next stored slice uses
    int cur_nt0;      // tile column whose scale / shift are in shared memory
    uint32_t res_cnt; // residual slices consumed so far (buffer = cnt & 1, parity = (cnt >> 1) & 1)
};

template <int BN, int NP>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, int lane, uint64_t* tmem_full_bar,
                                              uint32_t full_parity, uint64_t* tmem_empty_bar, const TileCoord& tc, int hl, int wl,
                                              int nl, int n_limit, EpiState& e, const CUtensorMap* mO0, const CUtensorMap* mO1,
                                              const CUtensorMap* mO2, const CUtensorMap* mO3, int etid,
                                              uint32_t tmem_empty_cluster_addr = 0) {
    using Cfg = ConvCfg<BN, NP>;
    using PC = PersCfg<BN, NP>;
    constexpr int NCOL = BN < 32 ? 16 : (BN > 128 ? 128 : BN);  // accumulator columns handled per pass (all threads)
    constexpr int NPASS = BN > 128 ? BN / 128 : 1;
    constexpr int NSL = NCOL >= 64 ? NCOL / 64 : 1;               // 64-channel slices per pass
    constexpr int MYC = BN >= 64 ? 32 * NSL : NCOL;               // columns this thread owns per pass
    constexpr bool SMEM_CONST = BN >= 64 && BN <= 128;            // scale / shift of the tile column staged in shared memory
    const int h = etid >> 7;                                     // column group
    const bool splitk = p.splitk_chunk > 0;
    const int z = tc.z, zi = splitk ? 0 : z;
    const int y = tc.y0 + hl, x = tc.x0 + wl, n = tc.n0 + nl;
    const bool valid = (y < p.H) && (x < p.W) && (n < n_limit);
    const int oy = y * p.sy + p.oy_off[zi], ox = x * p.sx + p.ox_off[zi];
    const long long pix = (static_cast<long long>(n) * p.OH + oy) * p.OW + ox;
    const bool tstore = BN >= 64 && p.tma_store && !splitk && p.act != ACT_HEADS;
    const bool nacc2 = PC::NACC == 2 && !p.single_acc;
    const float slope = p.act == ACT_LRELU ? 0.3f : (p.act == ACT_RELU ? 0.f : 1.f);  // act(v) = max(v, slope * v)
    const int nbuf = p.epi_bufs;
    const bool res_tma = BN >= 64 && p.res_tma != 0;
    const bool res_ldg = BN >= 64 && !res_tma && p.res_hi != nullptr && !splitk;

    if (SMEM_CONST && !splitk && tc.nt0 != e.cur_nt0) {
        // new tile column (at most grid_n * grid_z times per CTA): restage its folded BN constants
        named_bar_sync(1, kEpiThreads);
        if (etid < BN) {
            e.s_scale[etid] = __ldg(p.scale + tc.nt0 + etid);
            e.s_shift[etid] = __ldg(p.shift + tc.nt0 + etid);
        }
        named_bar_sync(1, kEpiThreads);
        e.cur_nt0 = tc.nt0;
    }
    mbar_wait(tmem_full_bar, full_parity);
    tc_fence_after();
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
        __syncwarp();  // lanes may arrive diverged (per-lane paths of the previous tile); tcgen05.ld is .aligned
        float acc[MYC];
        if (BN < 64) {
            // small tiles (heads BN = 16): group 0 does all the work
            if (h == 0) {
                uint32_t v[16];
                tmem_ld_32x16(taddr, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
                if (nacc2) {
                    uint32_t v1[16];
                    tmem_ld_32x16(taddr + PC::ACC_STRIDE, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v1[j]);
                }
            }
        } else {
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
                const uint32_t tcol = taddr + pass * NCOL + sl * 64 + h * 32;
                uint32_t v[32];
                tmem_ld_32x32(tcol, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[sl * 32 + j] = __uint_as_float(v[j]);
                if (nacc2) {
                    tmem_ld_32x32(tcol + PC::ACC_STRIDE, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]);
                }
            }
        }
        if (pass == NPASS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {  // TMEM is free: the next tile's MMAs may start
                if (tmem_empty_cluster_addr) mbar_arrive_cluster(tmem_empty_cluster_addr);  // CTA pair: the leader's barrier
                else mbar_arrive(tmem_empty_bar);
            }
        }

        if (BN < 64) {
            if (h == 0 && valid && p.act == ACT_HEADS) {
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const long long opix = (static_cast<long long>(n) * p.OH + (2 * y + (ph >> 1))) * p.OW + (2 * x + (ph & 1));
                    float* d = p.out_dec + opix * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        d[c] = tanhf(acc[ph * 4 + c] * __ldg(&p.scale[ph * 4 + c]) + __ldg(&p.shift[ph * 4 + c]));
                    const float ev = acc[ph * 4 + 3] * __ldg(&p.scale[ph * 4 + 3]) + __ldg(&p.shift[ph * 4 + 3]);
                    p.out_prob[opix] = 1.f / (1.f + expf(-ev));
                }
            }
            continue;
        }
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
            const int s = pass * NSL + sl;
            const SliceGeom g = slice_geom(p, tc, s, h, n, y, x, pix);
            const int c0 = g.c0s + h * 32;  // first of this thread's 32 accumulator columns
            if (splitk) {
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(p.out_partial + z * p.partial_stride + pix * p.Cout_pad + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        dst[q] = make_float4(acc[sl * 32 + 4 * q], acc[sl * 32 + 4 * q + 1], acc[sl * 32 + 4 * q + 2], acc[sl * 32 + 4 * q + 3]);
                }
                continue;
            }
            if (g.c0s >= p.Cout) continue;  // uniform over the CTA (never with res_tma: Cout is a multiple of BN there)
            __half* o_hi = p.out_hi + g.opix * p.Ctot + p.c_off + g.cch;
            const __half* r_hi = p.res_hi + g.opix * p.res_Ctot + g.cch;
            const int r = (nl * p.th + hl) * p.tw + wl;  // MMA row == row of the staging / residual box
            uint8_t* stage = e.stage0 - e.epi_buf * Cfg::EPI_BYTES;
            const uint8_t* rbuf = e.res0 - (e.res_cnt & 1u) * Cfg::EPI_BYTES;
            if (res_tma) mbar_wait(&e.res_bar[e.res_cnt & 1u], (e.res_cnt >> 1) & 1u);  // this slice's residual has landed
            if (tstore && nbuf == 1) {
                if (etid == 0) bulk_wait_read<0>();  // the previous TMA store has finished reading the only staging tile
                named_bar_sync(1, kEpiThreads);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t c16 = static_cast<uint32_t>(h * 4 + q);  // 16-byte chunk inside the 128-byte row
                const uint32_t off = static_cast<uint32_t>(r) * 128 + ((c16 ^ (static_cast<uint32_t>(r) & 7u)) << 4);
                float sc8[8], sh8[8];
                if (SMEM_CONST) {
                    const int cl = s * 64 + h * 32 + q * 8;  // column inside the tile
                    *reinterpret_cast<float4*>(sc8) = *reinterpret_cast<const float4*>(e.s_scale + cl);
                    *reinterpret_cast<float4*>(sc8 + 4) = *reinterpret_cast<const float4*>(e.s_scale + cl + 4);
                    *reinterpret_cast<float4*>(sh8) = *reinterpret_cast<const float4*>(e.s_shift + cl);
                    *reinterpret_cast<float4*>(sh8 + 4) = *reinterpret_cast<const float4*>(e.s_shift + cl + 4);
                } else {
                    *reinterpret_cast<float4*>(sc8) = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + q * 8));
                    *reinterpret_cast<float4*>(sc8 + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + q * 8 + 4));
                    *reinterpret_cast<float4*>(sh8) = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + q * 8));
                    *reinterpret_cast<float4*>(sh8 + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + q * 8 + 4));
                }
                uint4 rhv = make_uint4(0, 0, 0, 0), rlv = make_uint4(0, 0, 0, 0);
                if (res_tma) {
                    rhv = *reinterpret_cast<const uint4*>(rbuf + off);
                    if (NP == 2) rlv = *reinterpret_cast<const uint4*>(rbuf + 128 * 128 + off);
                } else if (res_ldg && valid) {
                    rhv = __ldg(reinterpret_cast<const uint4*>(r_hi + q * 8));
                    if (p.res_plane) rlv = __ldg(reinterpret_cast<const uint4*>(r_hi + p.res_plane + q * 8));
                }
                const uint32_t rh[4] = {rhv.x, rhv.y, rhv.z, rhv.w};
                const uint32_t rl[4] = {rlv.x, rlv.y, rlv.z, rlv.w};
                uint32_t oh[4], ol[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v0 = acc[sl * 32 + q * 8 + 2 * j] * sc8[2 * j] + sh8[2 * j];
                    float v1 = acc[sl * 32 + q * 8 + 2 * j + 1] * sc8[2 * j + 1] + sh8[2 * j + 1];
                    if (res_tma || res_ldg) {
                        const float2 a = h2_to_f2(rh[j]), b = h2_to_f2(rl[j]);
                        v0 += a.x + b.x;
                        v1 += a.y + b.y;
                    }
                    v0 = fmaxf(v0, slope * v0);
                    v1 = fmaxf(v1, slope * v1);
                    oh[j] = f2_to_h2(v0, v1);
                    const float2 back = h2_to_f2(oh[j]);
                    ol[j] = f2_to_h2(v0 - back.x, v1 - back.y);
                }
                const uint4 ohv = make_uint4(oh[0], oh[1], oh[2], oh[3]), olv = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                if (tstore) {
                    *reinterpret_cast<uint4*>(stage + off) = ohv;
                    if (NP == 2) *reinterpret_cast<uint4*>(stage + 128 * 128 + off) = olv;
                } else if (valid) {
                    *reinterpret_cast<uint4*>(o_hi + q * 8) = ohv;
                    if (p.out_plane) *reinterpret_cast<uint4*>(o_hi + p.out_plane + q * 8) = olv;
                }
            }
            if (tstore) {
                fence_proxy_async();
                // with >= 2 staging tiles one barrier per slice is enough: before it, the issuing thread makes sure the
                // store that last used the NEXT slice's tile has finished reading it
                if (etid == 0) {
                    if (nbuf == 2) bulk_wait_read<0>();
                    else if (nbuf == 3) bulk_wait_read<1>();
                }
                named_bar_sync(1, kEpiThreads);
                if (etid == 0) {
                    const CUtensorMap* mo = mO0;
                    if (p.sy == 2) {
                        const int mi = p.fused_cout ? g.phs : z;
                        mo = mi == 0 ? mO0 : (mi == 1 ? mO1 : (mi == 2 ? mO2 : mO3));
                    }
                    tma_store_5d(mo, stage, p.c_off + (g.cch - h * 32), tc.x0, tc.y0, tc.n0, 0);
                    bulk_commit_group();
                }
                e.epi_buf = e.epi_buf + 1 == nbuf ? 0 : e.epi_buf + 1;
            } else if (res_tma) {
                named_bar_sync(1, kEpiThreads);  // everyone is done with the residual tile
            }
            if (res_tma) {
                if (etid == 0) mbar_arrive(&e.res_bar[2 + (e.res_cnt & 1u)]);  // tile released: the producer refills it (slice + 2)
                ++e.res_cnt;
            }
        }
    }
}

template <int BN, int NP>
__global__ void __launch_bounds__(64 + kEpiThreads, 1)
conv_tc_persistent_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                          const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
                          const __grid_constant__ CUtensorMap mB, const __grid_constant__ CUtensorMap mO0,
                          const __grid_constant__ CUtensorMap mO1, const __grid_constant__ CUtensorMap mO2,
                          const __grid_constant__ CUtensorMap mO3, const __grid_constant__ CUtensorMap mR,
                          const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BN, NP>;
    using PC = PersCfg<BN, NP>;
    constexpr int STAGES = Cfg::STAGES;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int n_limit = p.N;
    if (p.n_active != nullptr) {
        const int na = *p.n_active;
        n_limit = na < n_limit ? na : n_limit;
    }
    const int total = p.grid_m * p.grid_n * p.grid_z;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;  // output staging (BN >= 64), 1024-B aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + (BN >= 64 ? Cfg::EPI_BYTES : 0));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]: one per accumulator set
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
    uint64_t* res_bar = tmem_empty_bar + 2;  // residual tiles (p.res_tma): [0..1] full, [2..3] empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);
    float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [BN] + [BN] (BN <= 128, Cfg::CONST_BYTES)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], kEpiWarps);  // one arrival per epilogue warp
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc<PC::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nst = p.nst > 0 && p.nst < STAGES ? p.nst : STAGES;  // short-K layers give their top stages to the epilogue
    if (warp == 0) {
        // ===================== TMA producer (whole warp convergent, TMA issue under elect_one; see the MMA issuer) =====================
        int st = 0;
        uint32_t ph = 0;  // running stage / phase (no per-iteration division: the stage count is a run-time value)
        uint32_t res_j = 0;  // residual slices issued
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            int4 k = __ldg(&p.kit[tc.kbeg]);
            for (int it = 0; it < tc.nk; ++it) {
                const int4 kn = __ldg(&p.kit[tc.kbeg + (it + 1 < tc.nk ? it + 1 : it)]);  // next entry, in flight during the wait
                mbar_wait(&empty_bar[st], ph ^ 1);
                uint8_t* sA = smem + st * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                const uint32_t tx = ((p.dbg & 1) ? 0 : (p.a_sw64 ? Cfg::A_BYTES / 2 : Cfg::A_BYTES)) + ((p.dbg & 2) ? 0 : Cfg::B_BYTES);
                const int mi = k.x & 0xff;
                const CUtensorMap* mA = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                if (elect_one()) {
                    if (tx == 0) {
                        mbar_arrive(&full_bar[st]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[st], tx);
                        if (!(p.dbg & 1)) tma_load_5d(mA, &full_bar[st], sA, k.w, tc.x0 + k.z, tc.y0 + k.y, tc.n0, 0);
                        if (!(p.dbg & 2)) tma_load_4d(&mB, &full_bar[st], sB, 0, tc.nt0, 0, tc.kbeg + it);
                    }
                }
                k = kn;
                if (++st == nst) { st = 0; ph ^= 1; }
            }
            if (BN >= 64 && p.res_tma) {
                // this tile's residual, one 64-channel slice per shared-memory tile, as soon as the epilogue has released it
#pragma unroll 1
                for (int sl = 0; sl < BN / 64; ++sl, ++res_j) {
                    const uint32_t b = res_j & 1u;
                    mbar_wait(&res_bar[2 + b], ((res_j >> 1) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&res_bar[b], Cfg::EPI_BYTES);
                        tma_load_5d(&mR, &res_bar[b], sEpi - (1 + b) * Cfg::EPI_BYTES, tc.nt0 + sl * 64, tc.x0, tc.y0, tc.n0, 0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The WHOLE warp runs this loop convergently and only the tcgen05 instructions sit under elect_one(): every
        // operand of tcgen05.mma must live in uniform registers, and inside an `if (lane == 0)` region the compiler keeps
        // the loop state in per-thread registers and wraps each MMA in R2UR moves plus an ELECT / BRA.U.ANY loop -- about
        // as long as the MMA itself runs, i.e. the kernel was issue-bound (98 instead of 75 cycles per MMA).  For the
        // same reason the per-iteration k16-step counts come from the kernel parameters (constant bank -> uniform
        // registers), not from global memory.
        constexpr uint32_t idesc = umma_idesc_f16(128, BN < 16 ? 16 : BN);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, PC::CONCAT ? 2 * BN : BN);  // A_hi * [W_hi; W_lo]^T
        const bool one_acc = p.single_acc != 0;
        const uint32_t cross = one_acc ? 0u : PC::ACC_STRIDE;
        const uint32_t smem_a0 = smem_u32(smem);
        int st = 0, tile_i = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            // accumulator set of this tile; the epilogue has drained the tile that used it last (NBUF tiles ago)
            const uint32_t buf = PC::NBUF == 2 ? (tile_i & 1) : 0;
            const uint32_t use = PC::NBUF == 2 ? (tile_i >> 1) : tile_i;
            mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * PC::BUF_COLS;
            int g = 0;
            for (int it = 0; it < tc.nk; ++it) {
                const int ksteps = p.ksteps_tab[tc.kbeg + it];
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t aA = smem_a0 + st * Cfg::STAGE_BYTES;
                const uint32_t aB = aA + Cfg::A_BYTES;
                // One elected region per k-iteration (<= 4 k16 steps, unrolled): the ELECT / BRA.DIV / BSYNC sequence and the
                // descriptor set-up are paid once, not per k16 step -- narrow tiles (heads N = 16, stem, ResNet 1x1) were bound
                // by this issue loop (~100 cycles per MMA), not by the tensor pipe.  Descriptors advance by 32 bytes = 2 units.
                const uint64_t dA = p.a_sw64 ? umma_desc_sw64(aA) : umma_desc_sw128(aA), dB = umma_desc_sw128(aB);
                const uint64_t dAlo = p.a_sw64 ? umma_desc_sw64(aA + 128 * 64) : umma_desc_sw128(aA + 128 * 128), dBlo = umma_desc_sw128(aB + BN * 128);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (kk < ksteps) {
                            const uint64_t a_hi = dA + 2 * kk, b_hi = dB + 2 * kk;
                            const uint32_t acc_g = (g + kk) > 0 ? 1u : 0u;
                            if (NP == 2) {
                                const uint64_t a_lo = dAlo + 2 * kk;
                                if (PC::CONCAT && !one_acc) {
                                    umma_f16(tacc, a_hi, b_hi, idesc2, acc_g);          // hi*hi -> [0, BN), hi*lo -> [BN, 2 BN)
                                    umma_f16(tacc + cross, a_lo, b_hi, idesc, 1u);     // lo*hi -> [BN, 2 BN)
                                } else {
                                    // same order of additions into the cross accumulator as the widened scheme above
                                    // (hi*lo, then lo*hi), so every kernel variant produces identical bits
                                    const uint64_t b_lo = dBlo + 2 * kk;
                                    umma_f16(tacc, a_hi, b_hi, idesc, acc_g);
                                    umma_f16(tacc + cross, a_hi, b_lo, idesc, ((g + kk) > 0 || one_acc) ? 1u : 0u);
                                    umma_f16(tacc + cross, a_lo, b_hi, idesc, 1u);
                                }
                            } else {
                                umma_f16(tacc, a_hi, b_hi, idesc, acc_g);
                            }
                        }
                    }
                }
                g += ksteps;
                if (elect_one()) umma_commit(&empty_bar[st]);
                if (++st == nst) { st = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full_bar[buf]);
            ++tile_i;
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int wl = r % p.tw;
        const int hl = (r / p.tw) % p.th;
        const int nl = r / (p.tw * p.th);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int etid = static_cast<int>(threadIdx.x) - 64;
        EpiState e;
        e.stage0 = sEpi;
        e.res0 = sEpi - Cfg::EPI_BYTES;  // residual tiles 0 / 1 below the staging tile (operand stages >= nst are unused)
        e.res_bar = res_bar;
        e.s_scale = s_scale;
        e.s_shift = s_scale + BN;
        e.epi_buf = 0;
        e.cur_nt0 = -1;
        e.res_cnt = 0;
        int tile_i = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            const uint32_t buf = PC::NBUF == 2 ? (tile_i & 1) : 0;
            const uint32_t use = PC::NBUF == 2 ? (tile_i >> 1) : tile_i;
            epilogue_tile<BN, NP>(p, taddr + buf * PC::BUF_COLS, lane, &tmem_full_bar[buf], use & 1, &tmem_empty_bar[buf], tc, hl, wl, nl,
                                  n_limit, e, &mO0, &mO1, &mO2, &mO3, etid);
            ++tile_i;
        }
        if (threadIdx.x == 64) bulk_wait_all();  // outstanding TMA stores complete before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<PC::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
