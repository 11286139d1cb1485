// CUDA-core helper kernels of the generator: stem im2col (Cin = 3) and the ResNet max-pool.
// Reference: resnet50_mod.py:200-204 (ZeroPadding2D(3)+Conv2D(7x7,s2,valid), MaxPooling2D(3x3,s2,same)),
// ae_model.py:74-80 (paper backbone conv1_1/conv1_2: Conv2D(5x5,s2,same) on the RGB crop).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace p2p {

__device__ __forceinline__ void split_store(__half* hi, long long plane, long long idx, float v) {
    const __half h = __float2half_rn(v);
    hi[idx] = h;
    if (plane) hi[idx + plane] = __float2half_rn(v - __half2float(h));
}

// x: (N,128,128,3) fp32 NHWC.  patches: (N,64,64,Kpad) fp16 hi/lo, K index = (kh*ks + kw)*3 + c,
// input row = 2*oy + kh - pad (zero outside), columns likewise; K >= ks*ks*3 is zero padding.
__global__ void im2col_stem_kernel(const float* __restrict__ x, __half* __restrict__ out, long long plane,
                                   int N, int ks, int pad, int Kpad, const int* __restrict__ n_active) {
    int n_limit = N;
    if (n_active) n_limit = min(n_limit, *n_active);
    const long long total = static_cast<long long>(n_limit) * 64 * 64 * Kpad;
    const int kreal = ks * ks * 3;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kpad);
        const long long pix = i / Kpad;
        const int ox = static_cast<int>(pix & 63), oy = static_cast<int>((pix >> 6) & 63);
        const int n = static_cast<int>(pix >> 12);
        float v = 0.f;
        if (k < kreal) {
            const int c = k % 3, kw = (k / 3) % ks, kh = k / (3 * ks);
            const int iy = 2 * oy + kh - pad, ix = 2 * ox + kw - pad;
            if (iy >= 0 && iy < 128 && ix >= 0 && ix < 128) v = x[((static_cast<long long>(n) * 128 + iy) * 128 + ix) * 3 + c];
        }
        split_store(out, plane, i, v);
    }
}

// MaxPooling2D(3x3, s2, 'same') on (N,64,64,64) -> (N,32,32,64): TF pads 0 before / 1 after with -inf.
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, long long in_plane, __half* __restrict__ out,
                                    long long out_plane, int N, int H, int W, int C, const int* __restrict__ n_active) {
    int n_limit = N;
    if (n_active) n_limit = min(n_limit, *n_active);
    const int OH = H / 2, OW = W / 2;
    const long long total = static_cast<long long>(n_limit) * OH * OW * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long t = i / C;
        const int ox = static_cast<int>(t % OW);
        t /= OW;
        const int oy = static_cast<int>(t % OH);
        const int n = static_cast<int>(t / OH);
        float m = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix >= W) continue;
                const long long j = ((static_cast<long long>(n) * H + iy) * W + ix) * C + c;
                float v = __half2float(in[j]);
                if (in_plane) v += __half2float(in[j + in_plane]);
                m = fmaxf(m, v);
            }
        }
        split_store(out, out_plane, i, m);
    }
}

}  // namespace p2p
