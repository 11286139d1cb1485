// CUDA-core helper kernels of the generator: stem im2col (Cin = 3) and the ResNet max-pool.
// Reference: resnet50_mod.py:200-204 (ZeroPadding2D(3)+Conv2D(7x7,s2,valid), MaxPooling2D(3x3,s2,same)),
// ae_model.py:74-80 (paper backbone conv1_1/conv1_2: Conv2D(5x5,s2,same) on the RGB crop).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace p2p {

// x: (N,128,128,3) fp32 NHWC.  patches: (N,64,64,Kpad) fp16 hi/lo, K index = (kh*ks + kw)*3 + c,
// input row = 2*oy + kh - pad (zero outside), columns likewise; K >= ks*ks*3 is zero padding.
// One thread produces 8 consecutive K entries (one 16-byte store per plane).
__global__ void im2col_stem_kernel(const float* __restrict__ x, __half* __restrict__ out, long long plane,
                                   int N, int ks, int pad, int Kpad, const int* __restrict__ n_active) {
    int n_limit = N;
    if (n_active) n_limit = min(n_limit, *n_active);
    const int kg = Kpad / 8;
    const long long total = static_cast<long long>(n_limit) * 64 * 64 * kg;
    const int kreal = ks * ks * 3;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k0 = static_cast<int>(i % kg) * 8;
        const long long pix = i / kg;
        const int ox = static_cast<int>(pix & 63), oy = static_cast<int>((pix >> 6) & 63);
        const int n = static_cast<int>(pix >> 12);
        uint4 hi4, lo4;
        __half* hh = reinterpret_cast<__half*>(&hi4);
        __half* lh = reinterpret_cast<__half*>(&lo4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            float v = 0.f;
            if (k < kreal) {
                const int c = k % 3, kw = (k / 3) % ks, kh = k / (3 * ks);
                const int iy = 2 * oy + kh - pad, ix = 2 * ox + kw - pad;
                if (iy >= 0 && iy < 128 && ix >= 0 && ix < 128) v = __ldg(&x[((static_cast<long long>(n) * 128 + iy) * 128 + ix) * 3 + c]);
            }
            const __half h = __float2half_rn(v);
            hh[j] = h;
            lh[j] = __float2half_rn(v - __half2float(h));
        }
        const long long o = pix * Kpad + k0;
        *reinterpret_cast<uint4*>(out + o) = hi4;
        if (plane) *reinterpret_cast<uint4*>(out + o + plane) = lo4;
    }
}

// Stem input for the direct (no im2col) first convolution: x (N,128,128,3) fp32 -> (N,128,Wp,4) fp16 hi/lo, channel 3 = 0,
// real pixels at columns pl .. pl+127 (the columns around them stay zero from the allocation memset: they are the
// horizontal zero padding).  With 4 channels a pixel is 8 bytes, so the 16-byte-strided, overlapping TMA view
// "window of 16 pixels starting at pixel 2*ox" (engine.cu, K_STEM) presents every kernel ROW (kw, c) of a stride-2 stem
// convolution as the contiguous K slice of output pixel ox -- the 805 MB patch matrix of the im2col path is never built.
__global__ void stem_convert_kernel(const float* __restrict__ x, __half* __restrict__ out, long long plane, int N, int Wp, int pl,
                                    const int* __restrict__ n_active) {
    int n_limit = N;
    if (n_active) n_limit = min(n_limit, *n_active);
    const long long total = static_cast<long long>(n_limit) * 128 * 128;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(i & 127);
        const long long row = i >> 7;  // n * 128 + y
        const float r = __ldg(x + i * 3), g = __ldg(x + i * 3 + 1), b = __ldg(x + i * 3 + 2);
        const __half2 h01 = __floats2half2_rn(r, g), h23 = __floats2half2_rn(b, 0.f);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(r - f01.x, g - f01.y), l23 = __floats2half2_rn(b - f23.x, 0.f);
        const long long o = (row * Wp + pl + xx) * 4;
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(out + o) = hv;
        if (plane) *reinterpret_cast<uint2*>(out + o + plane) = lv;
    }
}

// MaxPooling2D(3x3, s2, 'same') on (N,64,64,64) -> (N,32,32,64): TF pads 0 before / 1 after with -inf.
// One thread handles 8 channels (16-byte loads / stores per plane).
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, long long in_plane, __half* __restrict__ out,
                                    long long out_plane, int N, int H, int W, int C, const int* __restrict__ n_active) {
    int n_limit = N;
    if (n_active) n_limit = min(n_limit, *n_active);
    const int OH = H / 2, OW = W / 2, cg = C / 8;
    const long long total = static_cast<long long>(n_limit) * OH * OW * cg;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c0 = static_cast<int>(i % cg) * 8;
        long long t = i / cg;
        const int ox = static_cast<int>(t % OW);
        t /= OW;
        const int oy = static_cast<int>(t % OH);
        const int n = static_cast<int>(t / OH);
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix >= W) continue;
                const long long j0 = ((static_cast<long long>(n) * H + iy) * W + ix) * C + c0;
                const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(in + j0));
                uint4 l4 = make_uint4(0, 0, 0, 0);
                if (in_plane) l4 = __ldg(reinterpret_cast<const uint4*>(in + j0 + in_plane));
                const __half* hh = reinterpret_cast<const __half*>(&h4);
                const __half* lh = reinterpret_cast<const __half*>(&l4);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __half2float(hh[j]) + __half2float(lh[j]));
            }
        }
        uint4 oh, ol;
        __half* ohh = reinterpret_cast<__half*>(&oh);
        __half* olh = reinterpret_cast<__half*>(&ol);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __half h = __float2half_rn(m[j]);
            ohh[j] = h;
            olh[j] = __float2half_rn(m[j] - __half2float(h));
        }
        const long long o = i * 8;
        *reinterpret_cast<uint4*>(out + o) = oh;
        if (out_plane) *reinterpret_cast<uint4*>(out + o + out_plane) = ol;
    }
}

}  // namespace p2p
