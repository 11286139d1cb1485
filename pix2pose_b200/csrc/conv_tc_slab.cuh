// Slab-reuse variant of the persistent implicit-GEMM convolution for k x k stride-1 'same' convs (k = 3, 5) with
// 128 output channels: deconv3 of the decoder (ae_model.py:227-230, 23 % of the network's FLOPs) and the 3x3 convs of
// the stage-3 ResNet bottlenecks (resnet50_mod.py:62-65).
//
// Why: with N = 128 the generic kernel moves 64 KB of operands into the SM per k-iteration (32 KB activations + 32 KB
// weights for 768 MMA cycles) and is bound by that traffic, not by the tensor pipe (profiles/r01b_traffic_experiments.log:
// deconv3 1637 us, 1143 us with the activation loads switched off, 1107 us with no loads at all).  The generic kernel
// re-fetches the 128-pixel activation tile for each of the k*k taps (shifted copies, all L2 hits).  Here the output tile
// is 8 (w) x 16 (h) pixels and for each 64-channel chunk and each horizontal tap offset dx ONE slab of 8 x (16 + 2*pad)
// pixels is loaded; the k vertical taps dy are then plain descriptor offsets into it (MMA row r = h*8 + w  <->  slab row
// dy*8 + r: contiguous 128-byte rows whose start is 1024-byte aligned for every dy, so the ordinary K-major
// SWIZZLE_128B descriptor applies; shifting by dx inside a wider slab does not work, an 8-row core-matrix group may not
// straddle a swizzle atom).  Activation traffic drops by k * 16 / (16 + 2*pad): 4x for 5x5.  Weights stream per tap
// through their own ring of stages.
//
// Everything else is the persistent kernel's: convergent producer / MMA warps, widened (hi*hi | hi*lo) MMA, two
// double-buffered accumulator sets in TMEM, eight epilogue warps with TMA-store staging (epilogue_tile).
#pragma once
#include "conv_tc_persistent.cuh"

namespace p2p {

template <int BN, int NP>
struct SlabCfg {
    static constexpr int SLAB_W = 8;
    static constexpr int TILE_H = 16;
    // BN = 16 (the heads: 3 x 3 taps, 128 input channels, 16 accumulator columns).  A slab feeds only 3 taps x 8 narrow MMAs
    // (a few hundred cycles of tensor work), far less than a TMA round trip, so the layer was bound by how many operand
    // stages are in flight, not by MMAs or bytes: the generic 6-stage ring gave ~690 cycles per tap (heads 360 us per 256
    // crops; 414 us on this kernel with a 2-deep slab ring and the 6-stage weight ring).  Here ALL the layer's weights
    // (18 k-iterations x 4 KB) are loaded once per CTA and stay resident (RES_B), so the only ring is the slab ring, 4 deep.
    static constexpr bool RES_B = BN < 64;
    static constexpr int RES_KITERS = 18;                             // resident weight tiles (2 chunks x 9 taps)
    static constexpr int MAX_ROWS = RES_B ? TILE_H + 2 : TILE_H + 4;   // halo rows above and below: k = 3 / k <= 5
    static constexpr int SLAB_BYTES = NP * SLAB_W * MAX_ROWS * 128;   // 40 KB (fp16x3, k <= 5), 36 KB (heads)
    static constexpr int SLAB_BUFS = RES_B ? 4 : 2;
    static constexpr int B_BYTES = NP * BN * 128;
    static constexpr int EPI_BYTES = RES_B ? 0 : NP * 128 * 128;      // the heads store fp32 straight from registers
    static constexpr int B_ROOM = 224 * 1024 - SLAB_BUFS * SLAB_BYTES - EPI_BYTES - 2048;
    static constexpr int B_STAGES = RES_B ? RES_KITERS : (B_ROOM / B_BYTES > 6 ? 6 : B_ROOM / B_BYTES);
    static constexpr int CONST_BYTES = 2 * (BN < 64 ? 64 : BN) * 4;
    static constexpr int SMEM_BYTES = SLAB_BUFS * SLAB_BYTES + B_STAGES * B_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/ + CONST_BYTES;
    static_assert(B_STAGES >= 2 && SMEM_BYTES <= 227 * 1024, "shared-memory plan does not fit");
    static_assert((2 * (RES_B ? 1 : B_STAGES) + 2 * SLAB_BUFS + 4) * 8 + 4 <= 256, "barrier area");
};

// slab table entry (p.kit): {map index | (k16 steps << 8), first channel, first packed-weight k-iteration of
// (source, tap 0, chunk), chunks of the source}; p.kstart[1] = number of entries; p.slab_ksize = k.
//
// CL > 1: the kernel runs in clusters of CL CTAs (launch attribute).  Every CTA computes its own tile, but all tiles of a
// layer use the same weight tiles, so each CTA fetches only BN / CL rows of every weight stage and TMA-multicasts them into
// the shared memory of all CL CTAs (mB then has a box of BN / CL rows of ONE plane).  A weight stage is free again when the
// MMA warps of ALL CL CTAs have committed it (multicast commit, empty barrier count CL).  The CTAs of a cluster therefore
// run the same number of tiles: tile group g = tiles g*CL .. g*CL + CL-1, a group that sticks out past the last tile
// recomputes the last tile (identical bits stored twice).
template <int BN, int NP, int CL>
__global__ void __launch_bounds__(64 + kEpiThreads, 1)
conv_tc_slab_kernel(const __grid_constant__ CUtensorMap mS0, const __grid_constant__ CUtensorMap mS1,
                    const __grid_constant__ CUtensorMap mB, const __grid_constant__ CUtensorMap mO,
                    const __grid_constant__ ConvParams p) {
    static_assert(BN == 128 || BN == 64 || BN == 16, "slab tiles are 128 pixels x 128 / 64 channels (k x k convs) or x 16 accumulator columns (the heads)");
    static_assert(CL == 1 || CL == 2 || CL == 4, "cluster size");
    using SC = SlabCfg<BN, NP>;
    using PC = PersCfg<BN, NP>;
    constexpr int BS = SC::B_STAGES;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int n_limit = p.N;
    if (p.n_active != nullptr) {
        const int na = *p.n_active;
        n_limit = na < n_limit ? na : n_limit;
    }
    const int total = p.grid_m * p.grid_n;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const int group0 = static_cast<int>(blockIdx.x) / CL, ngroups = static_cast<int>(gridDim.x) / CL;
    const int total_groups = (total + CL - 1) / CL;
    constexpr uint16_t kAll = static_cast<uint16_t>((1u << CL) - 1u);
    const int ks = p.slab_ksize, pad = (ks - 1) / 2;
    const int n_slabs = p.kstart[1];
    const int rows = SC::TILE_H + 2 * pad;
    const uint32_t plane_off = SC::SLAB_W * rows * 128;   // the lo plane follows the hi box
    const uint32_t slab_tx = NP * plane_off;              // bytes one slab load delivers

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sSlab = smem;
    uint8_t* sB = smem + SC::SLAB_BUFS * SC::SLAB_BYTES;
    uint8_t* sEpi = sB + BS * SC::B_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(sEpi + SC::EPI_BYTES);
    constexpr int NBB = SC::RES_B ? 1 : BS;   // weight-ring barriers (resident weights: one "all landed" barrier)
    uint64_t* b_empty = b_full + NBB;
    uint64_t* slab_full = b_empty + NBB;
    uint64_t* slab_empty = slab_full + SC::SLAB_BUFS;
    uint64_t* tmem_full_bar = slab_empty + SC::SLAB_BUFS;  // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(b_full) + 256);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mS0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < NBB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], CL);  // the MMA warp of every CTA in the cluster
        }
        for (int s = 0; s < SC::SLAB_BUFS; ++s) {
            mbar_init(&slab_full[s], 1);
            mbar_init(&slab_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], kEpiWarps);
        }
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc<PC::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peers' barriers exist before any multicast touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tile of this CTA in tile group g (clamped: see the note on CL above)
    auto tile_of = [&](int g) { const int t = g * CL + static_cast<int>(crank); return t < total ? t : total - 1; };

    if (warp == 0) {
        // ===================== TMA producer (whole warp convergent, TMA issue under elect_one) =====================
        int bs = 0, ss = 0;
        uint32_t bph = 0, sph = 0;
        if (SC::RES_B) {
            // every weight tile of the layer, once (the host checked that they fit: n_slabs * ks * ks <= RES_KITERS)
            const int nk = n_slabs * ks * ks;
            if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(nk) * SC::B_BYTES);
                for (int kiter = 0; kiter < nk; ++kiter) tma_load_4d(&mB, &b_full[0], sB + kiter * SC::B_BYTES, 0, 0, 0, kiter);
            }
        }
        for (int g = group0; g < total_groups; g += ngroups) {
            const TileCoord tc = decode_tile<BN>(p, tile_of(g), n_limit);
            if (!decode_tile<BN>(p, g * CL, n_limit).live) continue;  // one decision per cluster
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int4 k = __ldg(&p.kit[sl]);
                const CUtensorMap* mS = (k.x & 0xff) == 0 ? &mS0 : &mS1;
                for (int dx = 0; dx < ks; ++dx) {
                    mbar_wait(&slab_empty[ss], sph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&slab_full[ss], slab_tx);
                        tma_load_5d(mS, &slab_full[ss], sSlab + ss * SC::SLAB_BYTES, k.y, tc.x0 - pad + dx, tc.y0 - pad, tc.n0, 0);
                    }
                    if (++ss == SC::SLAB_BUFS) { ss = 0; sph ^= 1; }
                    if (SC::RES_B) continue;
                    for (int dy = 0; dy < ks; ++dy) {
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&b_full[bs], SC::B_BYTES);
                            const int kiter = k.z + (dy * ks + dx) * k.w;
                            if (CL == 1) {
                                tma_load_4d(&mB, &b_full[bs], sB + bs * SC::B_BYTES, 0, tc.nt0, 0, kiter);
                            } else {
                                constexpr int RB = BN / CL;  // rows this CTA fetches for everybody
#pragma unroll
                                for (int pl = 0; pl < NP; ++pl)
                                    tma_load_4d_mc(&mB, &b_full[bs], sB + bs * SC::B_BYTES + (pl * BN + static_cast<int>(crank) * RB) * 128, 0,
                                                   tc.nt0 + static_cast<int>(crank) * RB, pl, kiter, kAll);
                            }
                        }
                        if (++bs == BS) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp convergent, see conv_tc_persistent.cuh) =====================
        constexpr uint32_t idesc = umma_idesc_f16(128, BN);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, PC::CONCAT ? 2 * BN : BN);
        const bool one_acc = p.single_acc != 0;
        const uint32_t cross = one_acc ? 0u : PC::ACC_STRIDE;
        const uint32_t slab0 = smem_u32(sSlab), b0 = smem_u32(sB);
        int bs = 0, ss = 0, tile_i = 0;
        uint32_t bph = 0, sph = 0;
        if (SC::RES_B) {
            mbar_wait(&b_full[0], 0);   // the resident weights have landed
            tc_fence_after();
        }
        for (int g = group0; g < total_groups; g += ngroups) {
            if (!decode_tile<BN>(p, g * CL, n_limit).live) continue;
            const uint32_t buf = tile_i & 1, use = tile_i >> 1;
            mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * PC::BUF_COLS;
            int gk = 0;  // k16 steps issued for this tile
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int ksteps = p.ksteps_tab[sl];
                const int4 kent = __ldg(&p.kit[sl]);
                for (int dx = 0; dx < ks; ++dx) {
                    mbar_wait(&slab_full[ss], sph);
                    tc_fence_after();
                    const uint32_t aSlab = slab0 + ss * SC::SLAB_BYTES;
                    for (int dy = 0; dy < ks; ++dy) {
                        if (!SC::RES_B) {
                            mbar_wait(&b_full[bs], bph);
                            tc_fence_after();
                        }
                        const uint32_t aA = aSlab + dy * SC::SLAB_W * 128;  // 1024-byte aligned for every dy
                        const uint32_t aB = b0 + (SC::RES_B ? kent.z + (dy * ks + dx) * kent.w : bs) * SC::B_BYTES;
                        const uint64_t dA = umma_desc_sw128(aA), dB = umma_desc_sw128(aB);
                        const uint64_t dAlo = umma_desc_sw128(aA + plane_off), dBlo = umma_desc_sw128(aB + BN * 128);
                        if (elect_one()) {   // one elected region per tap (see conv_tc_persistent.cuh)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (kk < ksteps) {
                                    const uint64_t a_hi = dA + 2 * kk, b_hi = dB + 2 * kk;
                                    const uint32_t acc_g = (gk + kk) > 0 ? 1u : 0u;
                                    if (NP == 2) {
                                        const uint64_t a_lo = dAlo + 2 * kk;
                                        if (!one_acc) {
                                            umma_f16(tacc, a_hi, b_hi, idesc2, acc_g);       // hi*hi -> [0, BN), hi*lo -> [BN, 2 BN)
                                            umma_f16(tacc + cross, a_lo, b_hi, idesc, 1u);  // lo*hi -> [BN, 2 BN)
                                        } else {
                                            const uint64_t b_lo = dBlo + 2 * kk;
                                            umma_f16(tacc, a_hi, b_hi, idesc, acc_g);
                                            umma_f16(tacc, a_hi, b_lo, idesc, 1u);
                                            umma_f16(tacc, a_lo, b_hi, idesc, 1u);
                                        }
                                    } else {
                                        umma_f16(tacc, a_hi, b_hi, idesc, acc_g);
                                    }
                                }
                            }
                        }
                        gk += ksteps;
                        if (!SC::RES_B) {
                            if (elect_one()) {
                                if (CL == 1) umma_commit(&b_empty[bs]);
                                else umma_commit_mc(&b_empty[bs], kAll);  // this stage is shared: free it in every CTA
                            }
                            if (++bs == BS) { bs = 0; bph ^= 1; }
                        }
                    }
                    if (elect_one()) umma_commit(&slab_empty[ss]);  // frees the slab once its k vertical taps have retired
                    if (++ss == SC::SLAB_BUFS) { ss = 0; sph ^= 1; }
                }
            }
            if (elect_one()) umma_commit(&tmem_full_bar[buf]);
            ++tile_i;
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int wl = r % p.tw;
        const int hl = (r / p.tw) % p.th;
        const int nl = r / (p.tw * p.th);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int etid = static_cast<int>(threadIdx.x) - 64;
        EpiState e;
        e.stage0 = sEpi;
        e.res0 = sEpi;
        e.res_bar = nullptr;
        e.s_scale = s_scale;
        e.s_shift = s_scale + BN;
        e.epi_buf = 0;
        e.cur_nt0 = -1;
        e.res_cnt = 0;
        int tile_i = 0;
        for (int g = group0; g < total_groups; g += ngroups) {
            const TileCoord tc = decode_tile<BN>(p, tile_of(g), n_limit);
            if (!decode_tile<BN>(p, g * CL, n_limit).live) continue;
            const uint32_t buf = tile_i & 1, use = tile_i >> 1;
            epilogue_tile<BN, NP>(p, taddr + buf * PC::BUF_COLS, lane, &tmem_full_bar[buf], use & 1, &tmem_empty_bar[buf], tc, hl, wl, nl,
                                  n_limit, e, &mO, &mO, &mO, &mO, etid);
            ++tile_i;
        }
        if (threadIdx.x == 64) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // nobody retires while a peer may still multicast into its shared memory / barriers
    if (warp == 2) tmem_dealloc<PC::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
