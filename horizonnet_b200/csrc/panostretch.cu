// pano_stretch: fused equirectangular stretch warp (reference misc/panostretch.py:81-102).
//
//   u0   = atan2(sin u * kx/ky, cos u)                       (:92)
//   v0   = atan(tan v * sin(u0)/sin u * ky)                  (:93)
//   refx = (u0/2pi + .5) W - .5 ,  refy = (v0/pi + .5) H - .5 (:95-96)
//   out  = scipy.ndimage.map_coordinates(img[...,c], [refy, refx], order, mode='wrap')   (:99-102)
//
// refx and g = sin(u0)/sin u * ky depend on the column only, so a prologue kernel builds two fp64
// tables of W entries per (kx, ky) pair and the main kernel spends one fp64 atan per FOUR pixels
// (rows y and H-1-y have tan v of opposite sign, so v0 is odd: refy(H-1-y) = H-1-refy(y); columns x
// and W-1-x have u of opposite sign, so g is even).
// Coordinates and the bilinear accumulation are fp64 like scipy (result cast to fp32); scipy's
// legacy 'wrap' folds coordinates with period n-1.  HBM-bound: 2*H*W*C*4 bytes per panorama.
#include "hn_common.cuh"

namespace hn {

namespace {

constexpr double PI_D = 3.14159265358979323846;

__global__ void stretch_tables_kernel(const double* __restrict__ kx, const double* __restrict__ ky,
                                      double* __restrict__ refx, double* __restrict__ gcol,
                                      double* __restrict__ tanv, int n, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * W) {
        const int img = i / W, x = i - img * W;
        const double u = (((double)x + 0.5) / (double)W - 0.5) * 2.0 * PI_D;     // panostretch.py:9
        const double su = sin(u), cu = cos(u);
        const double u0 = atan2(su * kx[img] / ky[img], cu);                      // :92
        refx[i] = (u0 / (2.0 * PI_D) + 0.5) * (double)W - 0.5;                    // :95
        gcol[i] = sin(u0) / su * ky[img];                                         // column factor of :93
    }
    if (i < H) {
        const double v = (((double)i + 0.5) / (double)H - 0.5) * PI_D;            // :10
        tanv[i] = tan(v);                                                         // :19
    }
}

// scipy ni_interpolation.c map_coordinate(), NI_EXTEND_WRAP (legacy wrap, period len-1)
__device__ __forceinline__ double legacy_wrap(double c, int len) {
    if (len <= 1) return 0.0;
    const double sz = (double)(len - 1);
    if (c < 0.0) c += sz * (floor(-c / sz) + 1.0);
    else if (c > sz) c -= sz * floor(c / sz);
    return c;
}

template <int C>
__device__ __forceinline__ void sample_store(const float* __restrict__ img, float* __restrict__ dst,
                                             double cy, double cx, int H, int W, int order) {
    cy = legacy_wrap(cy, H);
    cx = legacy_wrap(cx, W);
    if (order == 0) {
        const int iy = min((int)floor(cy + 0.5), H - 1), ix = min((int)floor(cx + 0.5), W - 1);
        const float* s = img + ((size_t)iy * W + ix) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = __ldg(s + c);
        return;
    }
    const int y0 = (int)floor(cy), x0 = (int)floor(cx);
    const double ty = cy - (double)y0, tx = cx - (double)x0;
    // index y0+1 == H only happens with weight exactly 0; fold it like scipy does (period n-1)
    const int y1 = (y0 + 1 > H - 1) ? (H > 1 ? y0 + 1 - (H - 1) : 0) : y0 + 1;
    const int x1 = (x0 + 1 > W - 1) ? (W > 1 ? x0 + 1 - (W - 1) : 0) : x0 + 1;
    const double w00 = __dmul_rn(1.0 - ty, 1.0 - tx), w01 = __dmul_rn(1.0 - ty, tx);
    const double w10 = __dmul_rn(ty, 1.0 - tx), w11 = __dmul_rn(ty, tx);
    const float* r0 = img + (size_t)y0 * W * C;
    const float* r1 = img + (size_t)y1 * W * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        // same accumulation order as scipy (no FMA contraction): (y0,x0) (y0,x1) (y1,x0) (y1,x1)
        double acc = __dmul_rn(w00, (double)__ldg(r0 + (size_t)x0 * C + c));
        acc = __dadd_rn(acc, __dmul_rn(w01, (double)__ldg(r0 + (size_t)x1 * C + c)));
        acc = __dadd_rn(acc, __dmul_rn(w10, (double)__ldg(r1 + (size_t)x0 * C + c)));
        acc = __dadd_rn(acc, __dmul_rn(w11, (double)__ldg(r1 + (size_t)x1 * C + c)));
        dst[c] = (float)acc;
    }
}

// grid: (ceil(ceil(W/2)/128), ceil(H/2), n); thread = column pair (x, W-1-x) x row pair (y, H-1-y).
// u(W-1-x) = -u(x) and v(H-1-y) = -v(y), so g is even in the column and v0 is odd in the row: one fp64
// atan serves four pixels (the kernel is bound by the fp64 pipe, not by HBM: ~90 DP ops per pixel otherwise).
template <int C>
__global__ void __launch_bounds__(128) stretch_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                      const double* __restrict__ refx,
                                                      const double* __restrict__ gcol,
                                                      const double* __restrict__ tanv, int H, int W, int order) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    const int xm = W - 1 - x;
    if (x > xm) return;
    const size_t plane = (size_t)H * W * C;
    const float* src = img + (size_t)n * plane;
    float* dst = out + (size_t)n * plane;
    const double v0 = atan(tanv[y] * gcol[(size_t)n * W + x]);                   // panostretch.py:93
    const double ry = (v0 / PI_D + 0.5) * (double)H - 0.5;                        // :96
    const double rym = (double)(H - 1) - ry;
    const int ym = H - 1 - y;
    float px[C];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int xx = side ? xm : x;
        if (side && xm == x) break;
        const double rx = refx[(size_t)n * W + xx];
        sample_store<C>(src, px, ry, rx, H, W, order);
#pragma unroll
        for (int c = 0; c < C; ++c) dst[((size_t)y * W + xx) * C + c] = px[c];
        if (ym != y) {
            sample_store<C>(src, px, rym, rx, H, W, order);
#pragma unroll
            for (int c = 0; c < C; ++c) dst[((size_t)ym * W + xx) * C + c] = px[c];
        }
    }
}

}  // namespace

// img/out: n images [H][W][C] fp32 on the device; kx/ky: n doubles on the device;
// scratch: (2*n*W + H) doubles on the device.
int pano_stretch_device(const float* img, float* out, int n, int H, int W, int C, const double* kx_dev,
                        const double* ky_dev, double* scratch, int order, cudaStream_t st) {
    HN_CHECK(order == 0 || order == 1, "pano_stretch: only order 0/1 are on the hot path (panostretch.py:86)");
    HN_CHECK(C >= 1 && C <= 4, "pano_stretch: 1..4 channels supported");
    HN_CHECK(n >= 0 && H >= 1 && W >= 1, "pano_stretch: bad geometry");
    if (n == 0) return 0;
    double* refx = scratch;
    double* gcol = scratch + (size_t)n * W;
    double* tanv = scratch + 2 * (size_t)n * W;
    const int tot = (n * W > H) ? n * W : H;
    stretch_tables_kernel<<<(tot + 255) / 256, 256, 0, st>>>(kx_dev, ky_dev, refx, gcol, tanv, n, H, W);
    HN_LAUNCH_OK();
    dim3 g(((W + 1) / 2 + 127) / 128, (H + 1) / 2, n);
    HN_CHECK(n <= 65535, "pano_stretch: at most 65535 images per call");
    switch (C) {
        case 1: stretch_kernel<1><<<g, 128, 0, st>>>(img, out, refx, gcol, tanv, H, W, order); break;
        case 2: stretch_kernel<2><<<g, 128, 0, st>>>(img, out, refx, gcol, tanv, H, W, order); break;
        case 3: stretch_kernel<3><<<g, 128, 0, st>>>(img, out, refx, gcol, tanv, H, W, order); break;
        default: stretch_kernel<4><<<g, 128, 0, st>>>(img, out, refx, gcol, tanv, H, W, order); break;
    }
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
