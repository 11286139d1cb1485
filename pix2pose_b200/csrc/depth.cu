// Depth -> camera-frame points and surface normals on the device: the per-pixel part of
// pix2pose_util/common_util.py:13-90 (getXYZ, get_normal) that tools/5_evaluation_bop_icp3d.py:78-79 and
// ros_kinetic/ros_pix2pose.py:180-181, 296-297 call to build the ICP point sets.  Everything is evaluated in double, in
// the reference's order of operations.  What stays on the host: cv2.inpaint (Navier-Stokes hole filling, get_normal
// refine=True, common_util.py:43-47) -- the Python wrapper pix2pose_b200/depth.py calls OpenCV for it, then this file
// does the Gaussian smoothing (:48), the gradient (:69) and the normals (:70-87).
#include "common.cuh"

#include <math.h>

#include <vector>

namespace p2p {
namespace {

// numpy writes `np.arange(n) - c` (float64) into an int16 array: C conversion, truncation toward zero (:16-18, :52-55)
__device__ __forceinline__ double uv_i16(int i, double c) { return static_cast<double>(static_cast<short>(static_cast<int>(static_cast<double>(i) - c))); }

// getXYZ (:13-30): xyz[v,u] = (u_tab * depth * 1 / fx, v_tab * depth * 1 / fy, depth) over the bbox [v0,u0,v1,u1)
__global__ void depth_xyz_kernel(const double* __restrict__ depth, int W, double fx, double fy, double cx, double cy, int v0, int u0,
                                 int h, int w, double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int i = p / w, j = p - i * w;
    const int v = v0 + i, u = u0 + j;
    const double d = depth[static_cast<long long>(v) * W + u];
    out[3 * p] = __ddiv_rn(__dmul_rn(__dmul_rn(uv_i16(u, cx), d), 1.0), fx);
    out[3 * p + 1] = __ddiv_rn(__dmul_rn(__dmul_rn(uv_i16(v, cy), d), 1.0), fy);
    out[3 * p + 2] = d;
}

// scipy.ndimage.gaussian_filter(x, sigma) (:48): separable correlation with the normalised kernel exp(-i^2 / (2 sigma^2)),
// radius int(4 sigma + 0.5), mode 'reflect' (d c b a | a b c d | d c b a); axis 0 first, then axis 1.
__device__ __forceinline__ int reflect_i(int i, int n) {
    // 'reflect' of scipy = half-sample symmetric, any distance
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}
__global__ void gauss_axis_kernel(const double* __restrict__ in, double* __restrict__ out, int H, int W, int axis,
                                  const double* __restrict__ wts, int radius) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int y = p / W, x = p - y * W;
    // scipy's correlate1d for symmetric kernels: centre tap, then pairs (left + right) * w from the inside out
    double acc = __dmul_rn(in[p], wts[0]);
    for (int k = radius; k >= 1; --k) {   // NI_Correlate1D walks the symmetric pairs from the outermost tap inwards
        double a, b;
        if (axis == 0) {
            a = in[static_cast<long long>(reflect_i(y - k, H)) * W + x];
            b = in[static_cast<long long>(reflect_i(y + k, H)) * W + x];
        } else {
            a = in[static_cast<long long>(y) * W + reflect_i(x - k, W)];
            b = in[static_cast<long long>(y) * W + reflect_i(x + k, W)];
        }
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), wts[k]));
    }
    out[p] = acc;
}

// np.gradient(f, 2, edge_order=2) along one axis of the h x w crop
__device__ __forceinline__ double grad2(const double* f, int idx, int n, long long stride, long long base) {
    const double h = 2.0;
    if (n < 3) return nan("");  // numpy raises for edge_order=2 with fewer than 3 samples; callers guard (bbox >= 5 px)
    // numpy (uniform spacing): interior (f[i+1] - f[i-1]) / (2h); edges with a = -1.5/h, b = 2/h, c = -0.5/h:
    // out[0] = a f0 + b f1 + c f2, out[-1] = -c f[-3] - b f[-2] - a f[-1]   (numpy/lib/function_base.py gradient)
    const double a = -1.5 / h, b = 2.0 / h, c = -0.5 / h;
    if (idx == 0) return __dadd_rn(__dadd_rn(__dmul_rn(a, f[base]), __dmul_rn(b, f[base + stride])), __dmul_rn(c, f[base + 2 * stride]));
    if (idx == n - 1)
        return __dadd_rn(__dadd_rn(__dmul_rn(-c, f[base + (idx - 2) * stride]), __dmul_rn(-b, f[base + (idx - 1) * stride])),
                         __dmul_rn(-a, f[base + idx * stride]));
    return __ddiv_rn(__dsub_rn(f[base + (idx + 1) * stride], f[base + (idx - 1) * stride]), 2.0 * h);
}

// get_normal (:50-89) after the optional refinement: tangent vectors from the depth gradient, normalised cross product
__global__ void depth_normal_kernel(const double* __restrict__ depth, int W, double fx, double fy, double cx, double cy, int v0,
                                    int u0, int h, int w, double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int i = p / w, j = p - i * w;
    const double kx = 1.0 / fx, ky = 1.0 / fy;
    const double us = uv_i16(u0 + j, cx), vs = uv_i16(v0 + i, cy);
    const long long org = static_cast<long long>(v0) * W + u0;
    const double d = depth[org + static_cast<long long>(i) * W + j];
    const double g0 = grad2(depth, i, h, W, org + j);                                   // d/dv (axis 0)
    const double g1 = grad2(depth, j, w, 1, org + static_cast<long long>(i) * W);      // d/du (axis 1)
    // no FMA contraction: numpy rounds every product and sum
    const double vy[3] = {__dmul_rn(__dmul_rn(us, kx), g0), __dadd_rn(__dmul_rn(d, ky), __dmul_rn(__dmul_rn(vs, ky), g0)), g0};
    const double vx[3] = {__dadd_rn(__dmul_rn(d, kx), __dmul_rn(__dmul_rn(us, kx), g1)), __dmul_rn(__dmul_rn(vs, ky), g1), g1};
    double c[3] = {__dsub_rn(__dmul_rn(vx[1], vy[2]), __dmul_rn(vx[2], vy[1])), __dsub_rn(__dmul_rn(vx[2], vy[0]), __dmul_rn(vx[0], vy[2])),
                   __dsub_rn(__dmul_rn(vx[0], vy[1]), __dmul_rn(vx[1], vy[0]))};
    double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2])));
    if (nrm == 0.0) nrm = 1.0;
    for (int k = 0; k < 3; ++k) {
        double v = __ddiv_rn(c[k], nrm);
        if (isnan(v)) v = 0.0;                                    // np.nan_to_num
        else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
        out[3 * p + k] = v;
    }
}

void check_bbox(const int* bbox, int H, int W, int& v0, int& u0, int& h, int& w) {
    v0 = 0; u0 = 0; h = H; w = W;
    if (bbox) {
        v0 = bbox[0]; u0 = bbox[1]; h = bbox[2] - bbox[0]; w = bbox[3] - bbox[1];
        P2P_CHECK(v0 >= 0 && u0 >= 0 && h >= 0 && w >= 0 && v0 + h <= H && u0 + w <= W, "bbox [%d,%d,%d,%d] outside the %d x %d image",
                  bbox[0], bbox[1], bbox[2], bbox[3], H, W);
    }
}

}  // namespace

void depth_xyz(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double* out) {
    require_device();
    P2P_CHECK(depth && out && H > 0 && W > 0, "bad argument");
    int v0, u0, h, w;
    check_bbox(bbox, H, W, v0, u0, h, w);
    if (h * w == 0) return;
    DevBuf<double> d, o;
    d.upload(depth, static_cast<size_t>(H) * W);
    o.alloc(static_cast<size_t>(h) * w * 3);
    depth_xyz_kernel<<<(h * w + 255) / 256, 256>>>(d.p, W, fx, fy, cx, cy, v0, u0, h, w, o.p);
    P2P_CUDA(cudaGetLastError());
    P2P_CUDA(cudaMemcpy(out, o.p, sizeof(double) * h * w * 3, cudaMemcpyDeviceToHost));
}

void depth_normals(const double* depth, int H, int W, double fx, double fy, double cx, double cy, const int* bbox, double sigma,
                   double* out) {
    require_device();
    P2P_CHECK(depth && out && H > 0 && W > 0, "bad argument");
    int v0, u0, h, w;
    check_bbox(bbox, H, W, v0, u0, h, w);
    if (h * w == 0) return;
    DevBuf<double> d, t, o, wd;
    d.upload(depth, static_cast<size_t>(H) * W);
    const int blocks = (H * W + 255) / 256;
    if (sigma > 0) {
        const int radius = static_cast<int>(4.0 * sigma + 0.5);
        std::vector<double> wts(radius + 1);
        double sum = 0;
        for (int k = -radius; k <= radius; ++k) sum += exp(-0.5 / (sigma * sigma) * k * k);
        for (int k = 0; k <= radius; ++k) wts[k] = exp(-0.5 / (sigma * sigma) * k * k) / sum;
        wd.upload(wts.data(), wts.size());
        P2P_CUDA(cudaStreamSynchronize(0));   // wts is a host temporary
        t.alloc(static_cast<size_t>(H) * W);
        gauss_axis_kernel<<<blocks, 256>>>(d.p, t.p, H, W, 0, wd.p, radius);
        gauss_axis_kernel<<<blocks, 256>>>(t.p, d.p, H, W, 1, wd.p, radius);
        P2P_CUDA(cudaGetLastError());
    }
    o.alloc(static_cast<size_t>(h) * w * 3);
    depth_normal_kernel<<<(h * w + 255) / 256, 256>>>(d.p, W, fx, fy, cx, cy, v0, u0, h, w, o.p);
    P2P_CUDA(cudaGetLastError());
    P2P_CUDA(cudaMemcpy(out, o.p, sizeof(double) * h * w * 3, cudaMemcpyDeviceToHost));
}

}  // namespace p2p
