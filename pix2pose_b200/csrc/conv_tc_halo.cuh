// Slab-reuse variant of the implicit-GEMM convolution for k x k stride-1 'same' convs (k = 3, 5):
// deconv1/2/3 of the decoder (66 % of the network's FLOPs, ae_model.py:209, 218, 228) and the 3x3
// convs of the ResNet bottlenecks (resnet50_mod.py:62-65).
//
// Idea: the generic kernel re-fetches the 128-pixel activation tile for every tap (25 shifted copies for 5x5, all L2
// hits).  Here the output tile is 8 (w) x 16 (h) pixels, and for each 64-channel chunk and each horizontal tap offset dx
// ONE slab of 8 x (16 + 2*pad) pixels is loaded; the k vertical taps dy are then plain descriptor offsets into it
// (MMA row r = h*8 + w  <->  slab pixel dy*8 + r: contiguous 128-byte rows, start 1024-B aligned for every dy, so the
// ordinary K-major SWIZZLE_128B descriptor applies -- a first version that also shifted by dx inside a 16-wide slab
// produced wrong results: an 8-row core-matrix group may not straddle a 1024-byte swizzle atom).  Activation traffic
// drops by k (5x for 5x5); weights stream per tap as before.
//
// Status: bit-for-bit as accurate as the generic kernel (decode max error 8.5e-5), but NOT faster (11.32 vs 11.25 ms per
// 256-crop forward): the experiments in profiles/r01_summary.md show that the big convolutions are bound by the tensor
// pipe's instruction rate, not by operand supply.  Kept behind P2P_HALO=1 as the starting point for an N = 256 /
// cta_group::2 version where the freed shared memory matters.
#pragma once
#include "conv_tc_persistent.cuh"

namespace p2p {

template <int BN, int NP>
struct HaloCfg {
    static constexpr int SLAB_W = 8;
    static constexpr int MAX_ROWS = 20;                               // 16 + 2*2
    static constexpr int SLAB_BYTES = NP * SLAB_W * MAX_ROWS * 128;   // 40 KB (fp16x3)
    static constexpr int SLAB_BUFS = 2;
    static constexpr int B_BYTES = NP * BN * 128;
    static constexpr int B_ROOM = 208 * 1024 - SLAB_BUFS * SLAB_BYTES;
    static constexpr int B_STAGES = B_ROOM / B_BYTES > 8 ? 8 : B_ROOM / B_BYTES;
    static constexpr int SMEM_BYTES = SLAB_BUFS * SLAB_BYTES + B_STAGES * B_BYTES + 1024 + 256 + 2048;  // + scale / shift staging
};

// slab table entry (p.kit): {map index | (k16 steps << 8), c0, first packed-weight k-iteration of (source, tap 0, chunk), chunks of the source}
template <int BN, int NP>
__global__ void __launch_bounds__(64 + kEpiThreads, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                    const __grid_constant__ CUtensorMap mB, const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BN, NP>;
    using HC = HaloCfg<BN, NP>;
    constexpr int BS = HC::B_STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int n_limit = p.N;
    if (p.n_active != nullptr) {
        const int na = *p.n_active;
        n_limit = na < n_limit ? na : n_limit;
    }
    const int total = p.grid_m * p.grid_n;
    const int ks = p.halo_ksize, pad = (ks - 1) / 2;
    const int n_slabs = p.kstart[1];
    const int rows = 16 + 2 * pad;
    const uint32_t slab_tx = static_cast<uint32_t>(NP) * HC::SLAB_W * rows * 128;  // bytes one slab load delivers
    const uint32_t plane_off = HC::SLAB_W * rows * 128;                             // lo plane follows the hi box

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sSlab = smem;
    uint8_t* sB = smem + HC::SLAB_BUFS * HC::SLAB_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(sB + BS * HC::B_BYTES);
    uint64_t* b_empty = b_full + BS;
    uint64_t* slab_full = b_empty + BS;
    uint64_t* slab_empty = slab_full + HC::SLAB_BUFS;
    uint64_t* tmem_full_bar = slab_empty + HC::SLAB_BUFS;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < BS; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < HC::SLAB_BUFS; ++s) {
            mbar_init(&slab_full[s], 1);
            mbar_init(&slab_empty[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, kEpiWarps);
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
            int bg = 0, sg = 0;  // weight-stage and slab use counters
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                for (int sl = 0; sl < n_slabs; ++sl) {
                    const int4 k = __ldg(&p.kit[sl]);
                    const CUtensorMap* mA = (k.x & 0xff) == 0 ? &mA0 : &mA1;
                    for (int dx = 0; dx < ks; ++dx, ++sg) {
                        const int sb = sg % HC::SLAB_BUFS;
                        mbar_wait(&slab_empty[sb], ((sg / HC::SLAB_BUFS) & 1) ^ 1);
                        mbar_arrive_expect_tx(&slab_full[sb], slab_tx);
                        tma_load_5d(mA, &slab_full[sb], sSlab + sb * HC::SLAB_BYTES, k.y, tc.x0 - pad + dx, tc.y0 - pad, tc.n0, 0);
                        for (int dy = 0; dy < ks; ++dy, ++bg) {
                            const int s = bg % BS;
                            mbar_wait(&b_empty[s], ((bg / BS) & 1) ^ 1);
                            mbar_arrive_expect_tx(&b_full[s], HC::B_BYTES);
                            tma_load_4d(&mB, &b_full[s], sB + s * HC::B_BYTES, 0, tc.nt0, 0, k.z + (dy * ks + dx) * k.w);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            constexpr uint32_t cross = (Cfg::NACC - 1) * Cfg::ACC_STRIDE;
            int bg = 0, sg = 0, tile_i = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                mbar_wait(tmem_empty_bar, (tile_i & 1) ^ 1);
                tc_fence_after();
                int g = 0;
                for (int sl = 0; sl < n_slabs; ++sl) {
                    const int ksteps = __ldg(&p.kit[sl].x) >> 8;
                    for (int dx = 0; dx < ks; ++dx, ++sg) {
                        const int sb = sg % HC::SLAB_BUFS;
                        mbar_wait(&slab_full[sb], (sg / HC::SLAB_BUFS) & 1);
                        tc_fence_after();
                        const uint32_t aSlab = smem_u32(sSlab + sb * HC::SLAB_BYTES);
                        for (int dy = 0; dy < ks; ++dy, ++bg) {
                            const int s = bg % BS;
                            mbar_wait(&b_full[s], (bg / BS) & 1);
                            tc_fence_after();
                            const uint32_t aA = aSlab + dy * HC::SLAB_W * 128;  // 1024-B aligned for every dy
                            const uint32_t aB = smem_u32(sB + s * HC::B_BYTES);
#pragma unroll 1
                            for (int kk = 0; kk < ksteps; ++kk, ++g) {
                                const uint64_t a_hi = umma_desc_sw128(aA + kk * 32);
                                const uint64_t b_hi = umma_desc_sw128(aB + kk * 32);
                                if (NP == 2) {
                                    if (Cfg::NACC == 3) umma_f16(tmem_base + (g & 1) * Cfg::ACC_STRIDE, a_hi, b_hi, idesc, g >= 2 ? 1u : 0u);
                                    else umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                                    const uint64_t a_lo = umma_desc_sw128(aA + plane_off + kk * 32);
                                    const uint64_t b_lo = umma_desc_sw128(aB + BN * 128 + kk * 32);
                                    umma_f16(tmem_base + cross, a_lo, b_hi, idesc, g > 0 ? 1u : 0u);
                                    umma_f16(tmem_base + cross, a_hi, b_lo, idesc, 1u);
                                } else {
                                    umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                                }
                            }
                            umma_commit(&b_empty[s]);
                        }
                        umma_commit(&slab_empty[sb]);  // frees the slab once its k taps have retired
                    }
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
        EpiState e;
        e.stage0 = nullptr; e.res0 = nullptr; e.res_bar = nullptr;
        e.s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~static_cast<uintptr_t>(15));
        e.s_shift = e.s_scale + BN;
        e.epi_buf = 0; e.cur_nt0 = -1; e.res_cnt = 0;
        int tile_i = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tc = decode_tile<BN>(p, t, n_limit);
            if (!tc.live) continue;
            epilogue_tile<BN, NP>(p, taddr, lane, tmem_full_bar, tile_i & 1, tmem_empty_bar, tc, hl, wl, nl, n_limit, e, nullptr, nullptr,
                                  nullptr, nullptr, static_cast<int>(threadIdx.x) - 64);
            ++tile_i;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace p2p
