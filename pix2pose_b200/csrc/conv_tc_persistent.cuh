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

// Epilogue of one tile for one epilogue thread (TMEM lane = accumulator row): drains the accumulators into
// registers (summing the three fp16x3 accumulators), releases TMEM to the MMA warp, then applies folded BN /
// bias / residual / activation, splits into fp16 hi/lo and stores NHWC.  Shared by the persistent kernels.
template <int BN, int NP>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, int lane, uint64_t* tmem_empty_bar,
                                              const TileCoord& tc, int hl, int wl, int nl, int n_limit) {
    using Cfg = ConvCfg<BN, NP>;
    constexpr int NCOL = BN < 32 ? 16 : (BN > 128 ? 128 : BN);  // accumulator columns drained per pass
    constexpr int NPASS = BN > 128 ? BN / 128 : 1;
    const bool splitk = p.splitk_chunk > 0;
    const int z = tc.z, zi = splitk ? 0 : z;
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t tcol = taddr + pass * NCOL;
    __syncwarp();  // lanes may arrive here diverged (per-lane store paths of the previous tile); tcgen05.ld is .aligned
    do {
            // ---- drain TMEM into registers, summing the fp16x3 accumulators
            float acc[NCOL];
            if (BN < 32) {
                uint32_t v[16];
                tmem_ld_32x16(taddr, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
                if (Cfg::NACC == 3) {
                    uint32_t v1[16], v2[16];
                    tmem_ld_32x16(taddr + Cfg::ACC_STRIDE, v1);
                    tmem_ld_32x16(taddr + 2 * Cfg::ACC_STRIDE, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = (acc[j] + __uint_as_float(v1[j])) + __uint_as_float(v2[j]);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < NCOL / 32; ++ch) {
                    uint32_t v[32];
                    tmem_ld_32x32(tcol + ch * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[ch * 32 + j] = __uint_as_float(v[j]);
                    if (Cfg::NACC >= 2) {
                        tmem_ld_32x32(tcol + Cfg::ACC_STRIDE + ch * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[ch * 32 + j] += __uint_as_float(v[j]);
                    }
                    if (Cfg::NACC == 3) {
                        tmem_ld_32x32(tcol + 2 * Cfg::ACC_STRIDE + ch * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[ch * 32 + j] += __uint_as_float(v[j]);
                    }
                }
            }
            if (pass == NPASS - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar);  // TMEM is free: the next tile's MMAs may start
            }

            // ---- BN / bias / residual / activation / hi-lo split / stores, from registers
            const int y = tc.y0 + hl, x = tc.x0 + wl, n = tc.n0 + nl;
            const bool valid = (y < p.H) && (x < p.W) && (n < n_limit);
            if (!valid) break;
            const int oy = y * p.sy + p.oy_off[zi], ox = x * p.sx + p.ox_off[zi];
            const long long pix = (static_cast<long long>(n) * p.OH + oy) * p.OW + ox;
            if (p.act == ACT_HEADS) {
                if (BN < 32) {
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
                break;
            }
            if (BN >= 32) {
#pragma unroll
                for (int ch = 0; ch < NCOL / 32; ++ch) {
                    const int c0 = tc.nt0 + pass * NCOL + ch * 32;
                    if (splitk) {
                        float4* dst = reinterpret_cast<float4*>(p.out_partial + z * p.partial_stride + pix * p.Cout_pad + c0);
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            dst[g] = make_float4(acc[ch * 32 + 4 * g], acc[ch * 32 + 4 * g + 1], acc[ch * 32 + 4 * g + 2], acc[ch * 32 + 4 * g + 3]);
                        continue;
                    }
                    if (c0 >= p.Cout) continue;
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
                    for (int g = 0; g < 4; ++g) {
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
                            float val = acc[ch * 32 + g * 8 + j] * __ldg(&p.scale[c]) + __ldg(&p.shift[c]);
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
    } while (false);
    }
}

template <int BN, int NP>
__global__ void __launch_bounds__(192, 1)
conv_tc_persistent_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                          const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
                          const __grid_constant__ CUtensorMap mB, const __grid_constant__ ConvParams p) {
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
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
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
        mbar_init(tmem_empty_bar, 4);  // one arrival per epilogue warp
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
            int itg = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                for (int it = 0; it < tc.nk; ++it, ++itg) {
                    const int s = itg % STAGES;
                    const uint32_t ph = (itg / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int4 k = __ldg(&p.kit[tc.kbeg + it]);
                    uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sB = sA + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    const int mi = k.x & 0xff;
                    const CUtensorMap* mA = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                    tma_load_5d(mA, &full_bar[s], sA, k.w, tc.x0 + k.z, tc.y0 + k.y, tc.n0, 0);
                    tma_load_4d(&mB, &full_bar[s], sB, 0, tc.nt0, 0, tc.kbeg + it);
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
                    const int s = itg % STAGES;
                    const uint32_t ph = (itg / STAGES) & 1;
                    const int ksteps = __ldg(&p.kit[tc.kbeg + it].x) >> 8;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t aA = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t aB = aA + Cfg::A_BYTES;
#pragma unroll 1
                    for (int kk = 0; kk < ksteps; ++kk, ++g) {
                        const uint64_t a_hi = umma_desc_sw128(aA + kk * 32);
                        const uint64_t b_hi = umma_desc_sw128(aB + kk * 32);
                        if (NP == 2) {
                            constexpr uint32_t cross = (Cfg::NACC - 1) * Cfg::ACC_STRIDE;
                            if (Cfg::NACC == 3) umma_f16(tmem_base + (g & 1) * Cfg::ACC_STRIDE, a_hi, b_hi, idesc, g >= 2 ? 1u : 0u);
                            else umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                            const uint64_t a_lo = umma_desc_sw128(aA + 128 * 128 + kk * 32);
                            const uint64_t b_lo = umma_desc_sw128(aB + BN * 128 + kk * 32);
                            umma_f16(tmem_base + cross, a_lo, b_hi, idesc, g > 0 ? 1u : 0u);
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
        int tile_i = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            mbar_wait(tmem_full_bar, tile_i & 1);
            tc_fence_after();
            epilogue_tile<BN, NP>(p, taddr, lane, tmem_empty_bar, tc, hl, wl, nl, n_limit);
            ++tile_i;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
