// CTA-pair variant of the persistent tap-list implicit-GEMM convolution (conv_tc_persistent.cuh): a cluster of two
// CTAs on the two SMs of one TPC computes a 256-pixel x BN tile with tcgen05.mma.cta_group::2 (M = 256).
//
// Why: the wide decoder convolutions are bound by operand traffic into the SM, not by the tensor pipe (measured by
// switching the TMA loads off, profiles/r01b_traffic_experiments.log: deconv3 1637 -> 1143 us without the activation
// loads, deconv2 1568 -> 1255 us without the weight loads).  In a pair every SM still fetches its own 128-pixel
// activation tile, but only HALF of the weight tile (BN / 2 rows): the tensor cores of both SMs read the two halves
// from both shared memories.  Per k-iteration and SM that is 32 + 16 KB instead of 32 + 32 KB at BN = 128 and
// 32 + 32 KB instead of 32 + 64 KB at BN = 256, it halves the shared-memory reads of the B operand, and the smaller
// stages leave room for one more pipeline stage.
//
// Protocol (rank 0 = leader):
//   * producer warp of EACH CTA: waits its own empty barrier, then announces its bytes on the LEADER's full barrier
//     (remote mbarrier.arrive.expect_tx) and issues its two TMA loads with cta_group::2, whose transaction bytes are
//     counted on the leader's barrier too (full barrier: 2 arrivals + both CTAs' bytes);
//   * MMA warp of the LEADER only: waits the full barrier, issues the M = 256 MMAs (hi*hi, hi*lo, lo*hi; the widened
//     (hi*hi | hi*lo) instruction of the single-CTA kernel does not exist here because each CTA holds one half of the
//     N rows), and commits with a multicast arrive on BOTH CTAs' empty barriers / accumulator-full barriers;
//   * epilogue warps of each CTA drain their own TMEM (rows of their own 128 pixels) exactly like the single-CTA
//     kernel and release the accumulator set on the leader's barrier (2 x 8 remote / local arrivals).
// Accumulator layout, epilogue, tensor maps of the activations and of the outputs are shared with the single-CTA kernel.
#pragma once
#include "conv_tc_persistent.cuh"

namespace p2p {

template <int BN, int NP>
struct PairCfg {
    static constexpr int A_BYTES = NP * 128 * 128;
    static constexpr int BH_BYTES = NP * (BN / 2) * 128;  // this CTA's half of the weight tile (hi plane, lo plane)
    static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
    static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > 6 ? 6 : (196 * 1024 / STAGE_BYTES);
    static constexpr int EPI_BYTES = NP * 128 * 128;
    static constexpr int CONST_BYTES = BN <= 128 ? 2 * BN * 4 : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + CONST_BYTES;
};

template <int BN, int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + kEpiThreads, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                    const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
                    const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mO0,
                    const __grid_constant__ CUtensorMap mO1, const __grid_constant__ CUtensorMap mO2,
                    const __grid_constant__ CUtensorMap mO3, const __grid_constant__ ConvParams p) {
    static_assert(BN == 128 || BN == 256, "pair tiles are 256 x 128 or 256 x 256");
    using Cfg = PairCfg<BN, NP>;
    using PC = PersCfg<BN, NP>;
    constexpr int STAGES = Cfg::STAGES;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    int n_limit = p.N;
    if (p.n_active != nullptr) {
        const int na = *p.n_active;
        n_limit = na < n_limit ? na : n_limit;
    }
    const int gm2 = (p.grid_m + 1) >> 1;  // tile pairs along M (an odd last tile is paired with an out-of-range one)
    const int total = gm2 * p.grid_n * p.grid_z;
    const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + Cfg::EPI_BYTES);  // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2], used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mBh);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 2);   // one arrive.expect_tx per CTA of the pair
            mbar_init(&empty_bar[s], 1);  // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps);  // every epilogue warp of both CTAs
        }
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc_pair<PC::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile of this CTA inside pair-tile tp
    auto decode = [&](int tp, TileCoord& tc) {
        const int mp = tp % gm2;
        const int r = tp / gm2;
        const int ny = r % p.grid_n;
        const int z = r / p.grid_n;
        tc = decode_mnz<BN>(p, 2 * mp + static_cast<int>(rank), ny, z, n_limit);
        const TileCoord lead = decode_mnz<BN>(p, 2 * mp, ny, z, n_limit);
        tc.live = lead.live;  // both CTAs of a pair take the same decision
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int st = 0;
        uint32_t ph = 0;
        for (int tp = pair0; tp < total; tp += npairs) {
            TileCoord tc;
            decode(tp, tc);
            if (!tc.live) continue;
            int4 k = __ldg(&p.kit[tc.kbeg]);
            for (int it = 0; it < tc.nk; ++it) {
                const int4 kn = __ldg(&p.kit[tc.kbeg + (it + 1 < tc.nk ? it + 1 : it)]);
                mbar_wait(&empty_bar[st], ph ^ 1);
                uint8_t* sA = smem + st * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                const int mi = k.x & 0xff;
                const CUtensorMap* mA = mi == 0 ? &mA0 : (mi == 1 ? &mA1 : (mi == 2 ? &mA2 : &mA3));
                const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[st]), 0);
                if (elect_one()) {
                    mbar_arrive_expect_tx_cluster(lead_full, Cfg::STAGE_BYTES);
                    tma_load_5d_pair(mA, lead_full, sA, k.w, tc.x0 + k.z, tc.y0 + k.y, tc.n0, 0);
                    tma_load_4d_pair(&mBh, lead_full, sB, 0, tc.nt0 + static_cast<int>(rank) * (BN / 2), 0, tc.kbeg + it);
                }
                k = kn;
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only; whole warp convergent, see conv_tc_persistent.cuh) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, BN);
            const bool one_acc = p.single_acc != 0;
            const uint32_t cross = one_acc ? 0u : PC::ACC_STRIDE;
            const uint32_t smem_a0 = smem_u32(smem);
            int st = 0, tile_i = 0;
            uint32_t ph = 0;
            for (int tp = pair0; tp < total; tp += npairs) {
                TileCoord tc;
                decode(tp, tc);
                if (!tc.live) continue;
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
                    const uint64_t dA = umma_desc_sw128(aA), dB = umma_desc_sw128(aB);
                    const uint64_t dAlo = umma_desc_sw128(aA + 128 * 128), dBlo = umma_desc_sw128(aB + (BN / 2) * 128);
                    if (elect_one()) {   // one elected region per k-iteration (see conv_tc_persistent.cuh)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < ksteps) {
                                const uint64_t a_hi = dA + 2 * kk, b_hi = dB + 2 * kk;
                                const uint32_t acc_g = (g + kk) > 0 ? 1u : 0u;
                                if (NP == 2) {
                                    // hi*hi, hi*lo, lo*hi: the order of the single-CTA kernel (bit-identical results)
                                    const uint64_t a_lo = dAlo + 2 * kk, b_lo = dBlo + 2 * kk;
                                    umma_f16_pair(tacc, a_hi, b_hi, idesc, acc_g);
                                    umma_f16_pair(tacc + cross, a_hi, b_lo, idesc, ((g + kk) > 0 || one_acc) ? 1u : 0u);
                                    umma_f16_pair(tacc + cross, a_lo, b_hi, idesc, 1u);
                                } else {
                                    umma_f16_pair(tacc, a_hi, b_hi, idesc, acc_g);
                                }
                            }
                        }
                    }
                    g += ksteps;
                    if (elect_one()) umma_commit_pair(&empty_bar[st], 3);  // frees this stage in both CTAs
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&tmem_full_bar[buf], 3);
                ++tile_i;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
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
        for (int tp = pair0; tp < total; tp += npairs) {
            TileCoord tc;
            decode(tp, tc);
            if (!tc.live) continue;
            const uint32_t buf = PC::NBUF == 2 ? (tile_i & 1) : 0;
            const uint32_t use = PC::NBUF == 2 ? (tile_i >> 1) : tile_i;
            const uint32_t lead_empty = mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0);
            epilogue_tile<BN, NP>(p, taddr + buf * PC::BUF_COLS, lane, &tmem_full_bar[buf], use & 1, &tmem_empty_bar[buf], tc, hl, wl, nl,
                                  n_limit, e, &mO0, &mO1, &mO2, &mO3, etid, lead_empty);
            ++tile_i;
        }
        if (threadIdx.x == 64) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody retires while the peer may still signal its barriers or read its shared memory
    if (warp == 2) tmem_dealloc_pair<PC::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
