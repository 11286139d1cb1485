// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) on one CTA per SM, for
//   N = 128 / 256, A from shared memory (SS) or from TMEM (TS), plain or the fp16x3 issue pattern,
//   with or without a concurrent stream of bulk global->shared copies (emulating the TMA operand feed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pix2pose_b200/csrc/sm100_ptx.cuh"

using namespace p2p;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// mode: 0 = SS x1, 1 = SS fp16x3 pattern (a_hi*b_hi, a_lo*b_hi, a_hi*b_lo), 2 = TS x1 (A in TMEM)
template <int N>
__global__ void __launch_bounds__(128, 1) bench(int iters, int mode, int feed_kb_per_iter, const uint8_t* gsrc, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                 // hi 16 KB + lo 16 KB
    uint8_t* sB = smem + 32768;         // hi N*128 + lo N*128
    uint8_t* sFeed = sB + 2 * N * 128;  // 64 KB landing zone for the copy stream
    uint64_t* bar = reinterpret_cast<uint64_t*>(sFeed + 65536);
    uint64_t* fbar = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(fbar + 1);
    for (int i = threadIdx.x; i < (32768 + 2 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(fbar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    if (mode >= 10 && threadIdx.x < 32) {
        // warp-uniform control flow, MMAs issued inside elect.sync (the CUTLASS / DeepGEMM idiom)
        const int m = mode - 10;
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t a_hi = umma_desc_sw128(aA + kk * 32), b_hi = umma_desc_sw128(aB + kk * 32);
                    if (m == 0) {
                        umma_f16(tm, a_hi, b_hi, idesc, 1u);
                    } else {
                        const uint64_t a_lo = umma_desc_sw128(aA + 16384 + kk * 32), b_lo = umma_desc_sw128(aB + N * 128 + kk * 32);
                        umma_f16(tm, a_hi, b_hi, idesc, 1u);
                        umma_f16(tm + 256, a_lo, b_hi, idesc, 1u);
                        umma_f16(tm + 256, a_hi, b_lo, idesc, 1u);
                    }
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    } else if (mode < 10 && threadIdx.x == 0) {
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t a_hi = umma_desc_sw128(aA + kk * 32), b_hi = umma_desc_sw128(aB + kk * 32);
                if (mode == 0) {
                    umma_f16(tm, a_hi, b_hi, idesc, 1u);
                } else if (mode == 1) {
                    const uint64_t a_lo = umma_desc_sw128(aA + 16384 + kk * 32), b_lo = umma_desc_sw128(aB + N * 128 + kk * 32);
                    umma_f16(tm, a_hi, b_hi, idesc, 1u);
                    umma_f16(tm + 256, a_lo, b_hi, idesc, 1u);
                    umma_f16(tm + 256, a_hi, b_lo, idesc, 1u);
                } else if (mode == 3) {
                    umma_f16(tm + (kk & 1) * 256, a_hi, b_hi, idesc, 1u);   // two independent accumulation chains
                } else if (mode == 4) {
                    umma_f16(tm + kk * 64, a_hi, b_hi, idesc, 1u);          // four independent chains
                } else {
                    umma_f16_ts(tm, tm + 256 + kk * 8, b_hi, idesc, 1u);
                }
            }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    } else if (threadIdx.x == 32 && feed_kb_per_iter > 0 && mode < 10) {
        // emulate the TMA operand feed: feed_kb_per_iter KB of bulk copies per "k-iteration"
        const long long t0 = clock64();
        const int per_mma = mode == 1 ? 12 : 4;
        for (int it = 0; it < iters; ++it) {
            mbar_arrive_expect_tx(fbar, feed_kb_per_iter * 1024);
            for (int c = 0; c < feed_kb_per_iter; c += 16)
                bulk_g2s(sFeed + (c % 64) * 1024, gsrc + (static_cast<size_t>(blockIdx.x) * 64 + (c % 64)) * 1024, 16384, fbar);
            mbar_wait(fbar, it & 1);
            // pace roughly like the MMA stream (per_mma MMAs of >= N/2 cycles each)
            while (clock64() - t0 < static_cast<long long>(it + 1) * per_mma * (N / 2)) {}
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N>
void run(const char* name, int mode, int feed, const uint8_t* gsrc, long long* dout, int sms) {
    const int iters = 2000;
    const int smem = 1024 + 32768 + 2 * N * 128 + 65536 + 64;
    cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<N><<<sms, 128, smem>>>(100, mode, feed, gsrc, dout);
    bench<N><<<sms, 128, smem>>>(iters, mode, feed, gsrc, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, dout, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0, sum = 0;
    for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    const double mmas = static_cast<double>(iters) * ((mode % 10) == 1 ? 12 : 4);   // modes 3 / 4: 4 MMAs per iteration as well
    printf("%-44s N=%3d feed %2d KB/iter : %.1f cycles/MMA (avg over SMs), %.1f (slowest SM); ideal %d\n", name, N, feed, sum / (double)sms / mmas,
           mx / mmas, N / 2);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint8_t* gsrc; long long* dout;
    cudaMalloc(&gsrc, static_cast<size_t>(sms) * 64 * 1024);
    cudaMemset(gsrc, 0, static_cast<size_t>(sms) * 64 * 1024);
    cudaMalloc(&dout, sizeof(long long) * 256);
    // narrow tiles: is there a floor per MMA, and is it the dependency through the accumulator?
    run<16>("SS, 1 chain", 0, 0, gsrc, dout, sms);
    run<16>("SS, 2 independent chains", 3, 0, gsrc, dout, sms);
    run<16>("SS, 4 independent chains", 4, 0, gsrc, dout, sms);
    run<32>("SS, 1 chain", 0, 0, gsrc, dout, sms);
    run<32>("SS, 2 independent chains", 3, 0, gsrc, dout, sms);
    run<64>("SS, 1 chain", 0, 0, gsrc, dout, sms);
    run<64>("SS, 2 independent chains", 3, 0, gsrc, dout, sms);
    run<128>("SS, 2 independent chains", 3, 0, gsrc, dout, sms);
    for (int feed : {0, 64}) {
        run<128>("SS, 1 MMA per k-step", 0, feed == 0 ? 0 : feed / 2, gsrc, dout, sms);
        run<128>("SS, fp16x3 pattern (3 MMAs per k-step)", 1, feed, gsrc, dout, sms);
        run<256>("SS, 1 MMA per k-step", 0, feed == 0 ? 0 : feed / 2, gsrc, dout, sms);
        run<256>("SS, fp16x3 pattern", 1, feed, gsrc, dout, sms);
    }
    run<128>("SS elect.sync idiom, 1 MMA per k-step", 10, 0, gsrc, dout, sms);
    run<128>("SS elect.sync idiom, fp16x3 pattern", 11, 0, gsrc, dout, sms);
    run<256>("SS elect.sync idiom, 1 MMA per k-step", 10, 0, gsrc, dout, sms);
    run<256>("SS elect.sync idiom, fp16x3 pattern", 11, 0, gsrc, dout, sms);
    run<128>("TS (A in TMEM), 1 MMA per k-step", 2, 0, gsrc, dout, sms);
    run<256>("TS (A in TMEM), 1 MMA per k-step", 2, 0, gsrc, dout, sms);
    return 0;
}
