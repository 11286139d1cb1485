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

struct TileCoord {
    int n0, y0, x0, nt0, z, kbeg, nk;
    bool live;
};

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t, int n_limit) {
    TileCoord c;
    const int m = t % p.grid_m;
    const int r = t / p.grid_m;
    const int ny = r % p.grid_n;
    c.z = r / p.grid_n;
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

// Epilogue of one tile.  kEpiWarps = 8 warps (256 threads): warp w may only touch TMEM lanes (w % 4) * 32 .. + 31, so
// two warps share each 32-row quarter and split the COLUMNS: of every 64-channel slice, group h (= warps 2-5 / 6-9)
// takes the 32 columns h*32 .. h*32+31.  Per tile each thread (a) drains its columns of the (up to three) accumulators
// into registers, (b) releases TMEM to the MMA warp, (c) applies folded BN / bias / residual / activation and the hi/lo
// split, (d) stages the 64-channel slice in a swizzled shared-memory tile that one thread writes out with a TMA store
// (direct global stores for the split-K, heads and BN < 64 paths).  Eight warps, not four: the epilogue is
// latency-bound (residual / scale loads) and with one warp per scheduler nothing hides that latency.
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;

template <int BN, int NP>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, int lane, uint64_t* tmem_empty_bar,
                                              const TileCoord& tc, int hl, int wl, int nl, int n_limit, uint8_t* stage,
                                              const CUtensorMap* const* omaps, int etid) {
    using Cfg = ConvCfg<BN, NP>;
    constexpr int NCOL = BN < 32 ? 16 : (BN > 128 ? 128 : BN);  // accumulator columns handled per pass (all threads)
    constexpr int NPASS = BN > 128 ? BN / 128 : 1;
    constexpr int NSL = NCOL >= 64 ? NCOL / 64 : 1;               // 64-channel slices per pass
    constexpr int MYC = BN >= 64 ? 32 * NSL : NCOL;               // columns this thread owns per pass
    const int h = etid >> 7;                                     // column group
    const bool splitk = p.splitk_chunk > 0;
    const int z = tc.z, zi = splitk ? 0 : z;
    const int y = tc.y0 + hl, x = tc.x0 + wl, n = tc.n0 + nl;
    const bool valid = (y < p.H) && (x < p.W) && (n < n_limit);
    const int oy = y * p.sy + p.oy_off[zi], ox = x * p.sx + p.ox_off[zi];
    const long long pix = (static_cast<long long>(n) * p.OH + oy) * p.OW + ox;
    const bool tstore = BN >= 64 && p.tma_store && !splitk && p.act != ACT_HEADS;
    const bool nacc3 = Cfg::NACC == 3 && !p.single_acc, nacc2 = Cfg::NACC >= 2 && !p.single_acc;
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
                if (nacc3) {
                    uint32_t v1[16], v2[16];
                    tmem_ld_32x16(taddr + Cfg::ACC_STRIDE, v1);
                    tmem_ld_32x16(taddr + 2 * Cfg::ACC_STRIDE, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = (acc[j] + __uint_as_float(v1[j])) + __uint_as_float(v2[j]);
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
                    tmem_ld_32x32(tcol + Cfg::ACC_STRIDE, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]);
                }
                if (nacc3) {
                    tmem_ld_32x32(tcol + 2 * Cfg::ACC_STRIDE, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]);
                }
            }
        }
        if (pass == NPASS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);  // TMEM is free: the next tile's MMAs may start
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
                    const float e = acc[ph * 4 + 3] * __ldg(&p.scale[ph * 4 + 3]) + __ldg(&p.shift[ph * 4 + 3]);
                    p.out_prob[opix] = 1.f / (1.f + expf(-e));
                }
            }
            continue;
        }
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
            const int c0s = tc.nt0 + pass * NCOL + sl * 64;  // first channel of the 64-wide slice
            const int c0 = c0s + h * 32;                     // first of this thread's 32 channels
            if (splitk) {
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(p.out_partial + z * p.partial_stride + pix * p.Cout_pad + c0);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        dst[g] = make_float4(acc[sl * 32 + 4 * g], acc[sl * 32 + 4 * g + 1], acc[sl * 32 + 4 * g + 2], acc[sl * 32 + 4 * g + 3]);
                }
                continue;
            }
            if (c0s >= p.Cout) continue;  // uniform over the CTA
            int phs = 0, cch = c0;
            long long opix = pix;
            if (p.fused_cout) {
                phs = c0s / p.fused_cout;
                cch = c0 - phs * p.fused_cout;
                opix = (static_cast<long long>(n) * p.OH + (2 * y + (phs >> 1))) * p.OW + (2 * x + (phs & 1));
            }
            const __half* r_hi = (p.res_hi && valid) ? p.res_hi + opix * p.res_Ctot + cch : nullptr;
            __half* o_hi = p.out_hi + opix * p.Ctot + p.c_off + cch;
            const int r = (nl * p.th + hl) * p.tw + wl;  // MMA row == row of the staging box
            if (tstore) {
                if (etid == 0) bulk_wait_read0();  // the previous TMA store has finished reading the staging tile
                named_bar_sync(1, kEpiThreads);
            }
            // residual and scale / shift loads first (independent), then the arithmetic
            uint4 rh[4], rl[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                rh[g] = make_uint4(0, 0, 0, 0); rl[g] = make_uint4(0, 0, 0, 0);
                if (r_hi) {
                    rh[g] = __ldg(reinterpret_cast<const uint4*>(r_hi + g * 8));
                    if (p.res_plane) rl[g] = __ldg(reinterpret_cast<const uint4*>(r_hi + p.res_plane + g * 8));
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float sc8[8], sh8[8];
                *reinterpret_cast<float4*>(sc8) = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + g * 8));
                *reinterpret_cast<float4*>(sc8 + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + g * 8 + 4));
                *reinterpret_cast<float4*>(sh8) = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + g * 8));
                *reinterpret_cast<float4*>(sh8 + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + g * 8 + 4));
                const __half* rhh = reinterpret_cast<const __half*>(&rh[g]);
                const __half* rlh = reinterpret_cast<const __half*>(&rl[g]);
                uint4 oh, ol;
                __half* ohh = reinterpret_cast<__half*>(&oh);
                __half* olh = reinterpret_cast<__half*>(&ol);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float val = acc[sl * 32 + g * 8 + j] * sc8[j] + sh8[j];
                    if (r_hi) val += __half2float(rhh[j]) + __half2float(rlh[j]);
                    val = act_apply(val, p.act);
                    const __half hv = __float2half_rn(val);
                    ohh[j] = hv;
                    olh[j] = __float2half_rn(val - __half2float(hv));
                }
                if (tstore) {
                    const uint32_t c16 = static_cast<uint32_t>(h * 4 + g);  // 16-byte chunk inside the 128-byte row
                    const uint32_t off = static_cast<uint32_t>(r) * 128 + ((c16 ^ (static_cast<uint32_t>(r) & 7u)) << 4);
                    *reinterpret_cast<uint4*>(stage + off) = oh;
                    if (NP == 2) *reinterpret_cast<uint4*>(stage + 128 * 128 + off) = ol;
                } else if (valid) {
                    *reinterpret_cast<uint4*>(o_hi + g * 8) = oh;
                    if (p.out_plane) *reinterpret_cast<uint4*>(o_hi + p.out_plane + g * 8) = ol;
                }
            }
            if (tstore) {
                fence_proxy_async();
                named_bar_sync(1, kEpiThreads);
                if (etid == 0) {
                    tma_store_5d(omaps[p.sy == 2 ? (p.fused_cout ? phs : z) : 0], stage, p.c_off + (cch - h * 32), tc.x0, tc.y0, tc.n0, 0);
                    bulk_commit_group();
                }
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
                          const __grid_constant__ CUtensorMap mO3, const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BN, NP>;
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
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, kEpiWarps);  // one arrival per epilogue warp
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nst = (p.dbg >> 8) > 0 && (p.dbg >> 8) < STAGES ? (p.dbg >> 8) : STAGES;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int itg = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                for (int it = 0; it < tc.nk; ++it, ++itg) {
                    const int s = itg % nst;
                    const uint32_t ph = (itg / nst) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int4 k = __ldg(&p.kit[tc.kbeg + it]);
                    uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sB = sA + Cfg::A_BYTES;
                    const uint32_t tx = ((p.dbg & 1) ? 0 : Cfg::A_BYTES) + ((p.dbg & 2) ? 0 : Cfg::B_BYTES);
                    if (tx == 0) { mbar_arrive(&full_bar[s]); continue; }
                    mbar_arrive_expect_tx(&full_bar[s], tx);
                    const int mi = k.x & 0xff;
                    const CUtensorMap* mA = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                    if (!(p.dbg & 1)) tma_load_5d(mA, &full_bar[s], sA, k.w, tc.x0 + k.z, tc.y0 + k.y, tc.n0, 0);
                    if (!(p.dbg & 2)) tma_load_4d(&mB, &full_bar[s], sB, 0, tc.nt0, 0, tc.kbeg + it);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN < 16 ? 16 : BN);
            int itg = 0, tile_i = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                mbar_wait(tmem_empty_bar, (tile_i & 1) ^ 1);  // epilogue has drained the previous tile's accumulators
                tc_fence_after();
                int g = 0;
                for (int it = 0; it < tc.nk; ++it, ++itg) {
                    const int s = itg % nst;
                    const uint32_t ph = (itg / nst) & 1;
                    const int ksteps = __ldg(&p.kit[tc.kbeg + it].x) >> 8;
                    if (!(p.dbg & 4)) mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t aA = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t aB = aA + Cfg::A_BYTES;
#pragma unroll 1
                    for (int kk = 0; kk < ksteps; ++kk, ++g) {
                        const uint64_t a_hi = umma_desc_sw128(aA + kk * 32);
                        const uint64_t b_hi = umma_desc_sw128(aB + kk * 32);
                        if (NP == 2) {
                            const uint32_t cross = p.single_acc ? 0u : (Cfg::NACC - 1) * Cfg::ACC_STRIDE;
                            if (Cfg::NACC == 3 && !p.single_acc) umma_f16(tmem_base + (g & 1) * Cfg::ACC_STRIDE, a_hi, b_hi, idesc, g >= 2 ? 1u : 0u);
                            else umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                            const uint64_t a_lo = umma_desc_sw128(aA + 128 * 128 + kk * 32);
                            const uint64_t b_lo = umma_desc_sw128(aB + BN * 128 + kk * 32);
                            umma_f16(tmem_base + cross, a_lo, b_hi, idesc, (g > 0 || p.single_acc) ? 1u : 0u);
                            umma_f16(tmem_base + cross, a_hi, b_lo, idesc, 1u);
                        } else {
                            umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(tmem_full_bar);
                ++tile_i;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int wl = r % p.tw;
        const int hl = (r / p.tw) % p.th;
        const int nl = r / (p.tw * p.th);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const CUtensorMap* const omaps[4] = {&mO0, &mO1, &mO2, &mO3};
        int tile_i = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            mbar_wait(tmem_full_bar, tile_i & 1);
            tc_fence_after();
            epilogue_tile<BN, NP>(p, taddr, lane, tmem_empty_bar, tc, hl, wl, nl, n_limit, BN >= 64 ? sEpi : nullptr, omaps,
                                  static_cast<int>(threadIdx.x) - 64);
            ++tile_i;
        }
        if (threadIdx.x == 64) bulk_wait_all();  // outstanding TMA stores complete before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
