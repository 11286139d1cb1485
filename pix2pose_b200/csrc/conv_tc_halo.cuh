// Halo-reuse variant of the implicit-GEMM convolution for k x k stride-1 'same' convs (k = 3, 5):
// deconv1/2/3 of the decoder (66 % of the network's FLOPs, ae_model.py:209, 218, 228) and the 3x3
// convs of the ResNet bottlenecks (resnet50_mod.py:62-65).
//
// The generic kernel re-fetches the 128-pixel A tile for every tap (25 shifted copies for 5x5): all L2
// hits, but 32 KB per 768 MMA cycles on top of the weights makes the kernel TMA-supply-bound
// (profiles/r01_summary.md: 66 % tensor-pipe active).  Here one TMA box per 64-channel slab brings the
// input tile WITH its halo -- 16 x (16 + 2*pad) pixels x 64 ch x {hi,lo} = 80 KB -- into shared memory once,
// and every tap's A operand is a UMMA descriptor pointing into that slab:
//   output tile = 8 (w) x 16 (h) pixels; MMA row r = h*8 + w; the 8 rows of a core-matrix group are the 8
//   consecutive pixels of one halo row (128 B apart), groups are one halo row (16 px = 2048 B) apart (SBO);
//   tap (dy,dx) starts at halo pixel dy*16 + dx, i.e. 128-B-aligned but not 1024-B-aligned, so the descriptor
//   carries base_offset = (addr >> 7) & 7 for the 128-byte swizzle phase.
// Only the weights stream per tap (4 stages x 32 KB).  A traffic drops ~10x (5x5) / ~4x (3x3).
#pragma once
#include "conv_tc_persistent.cuh"

namespace p2p {

// smem descriptor for a K-major 128B-swizzled operand whose 8-row groups are `sbo` bytes apart and whose
// start address is only 128-B aligned (base_offset carries the swizzle phase of the first row).
__device__ __forceinline__ uint64_t umma_desc_sw128_halo(uint32_t smem_addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;  // base offset
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <int BN, int NP>
struct HaloCfg {
    static constexpr int HALO_W = 16;
    static constexpr int MAX_ROWS = 20;                                // 16 + 2*2
    static constexpr int PLANE_BYTES = HALO_W * MAX_ROWS * 128;        // 40 KB per plane
    static constexpr int HALO_BYTES = NP * PLANE_BYTES;
    static constexpr int B_BYTES = NP * BN * 128;
    static constexpr int B_STAGES = (212 * 1024 - HALO_BYTES) / B_BYTES > 8 ? 8 : (212 * 1024 - HALO_BYTES) / B_BYTES;
    static constexpr int SMEM_BYTES = HALO_BYTES + B_STAGES * B_BYTES + 1024 + 256;
};

// slab table entry (p.kit): {map index | (k16 steps << 8), c0, first packed-weight k-iteration of (source, tap 0, chunk), chunks of the source}
template <int BN, int NP>
__global__ void __launch_bounds__(192, 1)
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
    const int ks = p.halo_ksize, pad = (ks - 1) / 2, ntaps = ks * ks;
    const int n_slabs = p.kstart[1];
    const uint32_t halo_tx = static_cast<uint32_t>(NP) * HC::HALO_W * (16 + 2 * pad) * 128;  // bytes one slab load delivers
    const uint32_t plane_off = HC::HALO_W * (16 + 2 * pad) * 128;                             // lo plane follows the hi box

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sHalo = smem;
    uint8_t* sB = smem + HC::HALO_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(sB + BS * HC::B_BYTES);
    uint64_t* b_empty = b_full + BS;
    uint64_t* halo_full = b_empty + BS;
    uint64_t* halo_empty = halo_full + 1;
    uint64_t* tmem_full_bar = halo_empty + 1;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mA0);
        tma_prefetch_desc(&mB);
        for (int s = 0; s < BS; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(halo_full, 1);
        mbar_init(halo_empty, 1);
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, 4);
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
            int bg = 0, hg = 0;  // weight-stage and halo use counters
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                for (int sl = 0; sl < n_slabs; ++sl, ++hg) {
                    const int4 k = __ldg(&p.kit[sl]);
                    mbar_wait(halo_empty, (hg & 1) ^ 1);
                    mbar_arrive_expect_tx(halo_full, halo_tx);
                    const CUtensorMap* mA = (k.x & 0xff) == 0 ? &mA0 : &mA1;
                    tma_load_5d(mA, halo_full, sHalo, k.y, tc.x0 - pad, tc.y0 - pad, tc.n0, 0);
                    for (int tap = 0; tap < ntaps; ++tap, ++bg) {
                        const int s = bg % BS;
                        mbar_wait(&b_empty[s], ((bg / BS) & 1) ^ 1);
                        mbar_arrive_expect_tx(&b_full[s], HC::B_BYTES);
                        tma_load_4d(&mB, &b_full[s], sB + s * HC::B_BYTES, 0, tc.nt0, 0, k.z + tap * k.w);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            const uint32_t aHalo = smem_u32(sHalo);
            int bg = 0, hg = 0, tile_i = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile<BN>(p, t, n_limit);
                if (!tc.live) continue;
                mbar_wait(tmem_empty_bar, (tile_i & 1) ^ 1);
                tc_fence_after();
                int g = 0;
                for (int sl = 0; sl < n_slabs; ++sl, ++hg) {
                    const int ksteps = __ldg(&p.kit[sl].x) >> 8;
                    mbar_wait(halo_full, hg & 1);
                    tc_fence_after();
                    for (int tap = 0; tap < ntaps; ++tap, ++bg) {
                        const int s = bg % BS;
                        mbar_wait(&b_full[s], (bg / BS) & 1);
                        tc_fence_after();
                        const int dy = tap / ks, dx = tap - dy * ks;
                        const uint32_t aA = aHalo + (dy * HC::HALO_W + dx) * 128;
                        const uint32_t aB = smem_u32(sB + s * HC::B_BYTES);
#pragma unroll 1
                        for (int kk = 0; kk < ksteps; ++kk, ++g) {
                            const uint64_t a_hi = umma_desc_sw128_halo(aA + kk * 32, HC::HALO_W * 128);
                            const uint64_t b_hi = umma_desc_sw128(aB + kk * 32);
                            if (NP == 2) {
                                umma_f16(tmem_base + (g & 1) * Cfg::ACC_STRIDE, a_hi, b_hi, idesc, g >= 2 ? 1u : 0u);
                                const uint64_t a_lo = umma_desc_sw128_halo(aA + plane_off + kk * 32, HC::HALO_W * 128);
                                const uint64_t b_lo = umma_desc_sw128(aB + BN * 128 + kk * 32);
                                umma_f16(tmem_base + 2 * Cfg::ACC_STRIDE, a_lo, b_hi, idesc, g > 0 ? 1u : 0u);
                                umma_f16(tmem_base + 2 * Cfg::ACC_STRIDE, a_hi, b_lo, idesc, 1u);
                            } else {
                                umma_f16(tmem_base, a_hi, b_hi, idesc, g > 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&b_empty[s]);
                    }
                    umma_commit(halo_empty);  // all taps of this slab have been issued; frees the halo when they retire
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
