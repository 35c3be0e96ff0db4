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
#include <cstdlib>
#include "hn_common.cuh"

namespace hn {

namespace {

constexpr double PI_D = 3.14159265358979323846;

// scipy ni_interpolation.c map_coordinate(), NI_EXTEND_WRAP (legacy wrap, period len-1)
__device__ __forceinline__ double legacy_wrap(double c, int len) {
    if (len <= 1) return 0.0;
    const double sz = (double)(len - 1);
    if (c < 0.0) c += sz * (floor(-c / sz) + 1.0);
    else if (c > sz) c -= sz * floor(c / sz);
    return c;
}

// Everything that depends on the column only, per (image, x): the row-coordinate factor g and the fully
// resolved horizontal taps (wrapped refx -> x0, x1, tx), so the per-pixel work is the vertical coordinate
// and the 2x2xC blend.
struct ColEntry {
    double g;        // sin(u0)/sin(u) * ky                        (column factor of panostretch.py:93)
    double tx;       // horizontal interpolation weight of x1
    int x0, x1;      // horizontal taps after scipy's legacy wrap (element offsets x*C are formed in the kernel)
    int xn;          // nearest column (order 0)
    int pad;
};

__global__ void stretch_tables_kernel(const double* __restrict__ kx, const double* __restrict__ ky,
                                      ColEntry* __restrict__ cols, double* __restrict__ tanv, int n, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * W) {
        const int img = i / W, x = i - img * W;
        const double u = (((double)x + 0.5) / (double)W - 0.5) * 2.0 * PI_D;     // panostretch.py:9
        const double su = sin(u), cu = cos(u);
        const double u0 = atan2(su * kx[img] / ky[img], cu);                      // :92
        const double refx = (u0 / (2.0 * PI_D) + 0.5) * (double)W - 0.5;          // :95
        const double cx = legacy_wrap(refx, W);
        const int x0 = (int)floor(cx);
        ColEntry e;
        e.g = sin(u0) / su * ky[img];
        e.tx = cx - (double)x0;
        e.x0 = x0;
        // index x0+1 == W only happens with weight exactly 0; fold it like scipy does (period n-1)
        e.x1 = (x0 + 1 > W - 1) ? (W > 1 ? x0 + 1 - (W - 1) : 0) : x0 + 1;
        e.xn = min((int)floor(cx + 0.5), W - 1);
        e.pad = 0;
        cols[i] = e;
    }
    if (i < H) {
        const double v = (((double)i + 0.5) / (double)H - 0.5) * PI_D;            // :10
        tanv[i] = tan(v);                                                         // :19
    }
}

struct RowTaps {       // vertical taps of one output row: resolved once, shared by both mirrored columns
    int r0, r1;        // element offsets y0*W*C, y1*W*C
    int rn;            // nearest row offset (order 0)
    double ty;
};

template <int C>
__device__ __forceinline__ RowTaps resolve_row(double cy, int H, int W) {
    cy = legacy_wrap(cy, H);
    const int y0 = (int)floor(cy);
    RowTaps t;
    t.ty = cy - (double)y0;
    // index y0+1 == H only happens with weight exactly 0; fold it like scipy does (period n-1)
    const int y1 = (y0 + 1 > H - 1) ? (H > 1 ? y0 + 1 - (H - 1) : 0) : y0 + 1;
    t.r0 = y0 * W * C;
    t.r1 = y1 * W * C;
    t.rn = min((int)floor(cy + 0.5), H - 1) * W * C;
    return t;
}

// grid: (ceil(ceil(W/2)/128), ceil(H/2), n); thread = column pair (x, W-1-x) x row pair (y, H-1-y).
// u(W-1-x) = -u(x) and v(H-1-y) = -v(y), so g is even in the column and v0 is odd in the row: one fp64
// atan serves four pixels.  ncu showed the kernel latency-bound (long-scoreboard stalls on the gathers, 30 %
// DRAM), so all taps of the four pixels (4 x 4 x C loads) are issued back to back before any blending.
template <int C>
__global__ void __launch_bounds__(128) stretch_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                      const ColEntry* __restrict__ cols,
                                                      const double* __restrict__ tanv, int H, int W, int order) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    const int xm = W - 1 - x;
    if (x > xm) return;
    const size_t plane = (size_t)H * W * C;
    const float* src = img + (size_t)n * plane;
    float* dst = out + (size_t)n * plane;
    const ColEntry* ce = cols + (size_t)n * W;
    const ColEntry e[2] = {ce[x], ce[xm]};
    const double v0 = atan(tanv[y] * e[0].g);                                    // panostretch.py:93
    const double ry = (v0 / PI_D + 0.5) * (double)H - 0.5;                        // :96
    const int ym = H - 1 - y;
    const RowTaps rt[2] = {resolve_row<C>(ry, H, W), resolve_row<C>((double)(H - 1) - ry, H, W)};
    const int orow[2] = {y * W * C, ym * W * C};
    const int ocol[2] = {x * C, xm * C};
    const int nside = (xm == x) ? 1 : 2, nvert = (ym == y) ? 1 : 2;
    if (order == 0) {
        for (int sd = 0; sd < nside; ++sd)
            for (int vt = 0; vt < nvert; ++vt)
#pragma unroll
                for (int c = 0; c < C; ++c) dst[orow[vt] + ocol[sd] + c] = __ldg(src + rt[vt].rn + e[sd].xn * C + c);
        return;
    }
    float tap[2][2][4][C];                       // [side][vert][tap][channel]
#pragma unroll
    for (int sd = 0; sd < 2; ++sd)
#pragma unroll
        for (int vt = 0; vt < 2; ++vt) {
            const float* p00 = src + rt[vt].r0 + e[sd].x0 * C;
            const float* p01 = src + rt[vt].r0 + e[sd].x1 * C;
            const float* p10 = src + rt[vt].r1 + e[sd].x0 * C;
            const float* p11 = src + rt[vt].r1 + e[sd].x1 * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                tap[sd][vt][0][c] = __ldg(p00 + c);
                tap[sd][vt][1][c] = __ldg(p01 + c);
                tap[sd][vt][2][c] = __ldg(p10 + c);
                tap[sd][vt][3][c] = __ldg(p11 + c);
            }
        }
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        if (sd >= nside) break;
#pragma unroll
        for (int vt = 0; vt < 2; ++vt) {
            if (vt >= nvert) break;
            const double ty = rt[vt].ty, tx = e[sd].tx;
            const double w00 = __dmul_rn(1.0 - ty, 1.0 - tx), w01 = __dmul_rn(1.0 - ty, tx);
            const double w10 = __dmul_rn(ty, 1.0 - tx), w11 = __dmul_rn(ty, tx);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // same accumulation order as scipy (no FMA contraction): (y0,x0) (y0,x1) (y1,x0) (y1,x1)
                double acc = __dmul_rn(w00, (double)tap[sd][vt][0][c]);
                acc = __dadd_rn(acc, __dmul_rn(w01, (double)tap[sd][vt][1][c]));
                acc = __dadd_rn(acc, __dmul_rn(w10, (double)tap[sd][vt][2][c]));
                acc = __dadd_rn(acc, __dmul_rn(w11, (double)tap[sd][vt][3][c]));
                dst[orow[vt] + ocol[sd] + c] = (float)acc;
            }
        }
    }
}

}  // namespace

// img/out: n images [H][W][C] fp32 on the device; kx/ky: n doubles on the device;
// scratch: (4*n*W + H) doubles on the device (n*W ColEntry records of 32 bytes, then H doubles).
int pano_stretch_device(const float* img, float* out, int n, int H, int W, int C, const double* kx_dev,
                        const double* ky_dev, double* scratch, int order, cudaStream_t st) {
    HN_CHECK(order == 0 || order == 1, "pano_stretch: only order 0/1 are on the hot path (panostretch.py:86)");
    HN_CHECK(C >= 1 && C <= 4, "pano_stretch: 1..4 channels supported");
    HN_CHECK(n >= 0 && H >= 1 && W >= 1, "pano_stretch: bad geometry");
    if (n == 0) return 0;
    HN_CHECK((long long)H * W * C < (1ll << 31), "pano_stretch: image too large");
    static_assert(sizeof(ColEntry) == 32, "ColEntry layout");
    ColEntry* cols = reinterpret_cast<ColEntry*>(scratch);
    double* tanv = scratch + 4 * (size_t)n * W;
    const int tot = (n * W > H) ? n * W : H;
    stretch_tables_kernel<<<(tot + 255) / 256, 256, 0, st>>>(kx_dev, ky_dev, cols, tanv, n, H, W);
    HN_LAUNCH_OK();
    dim3 g(((W + 1) / 2 + 127) / 128, (H + 1) / 2, n);
    HN_CHECK(n <= 65535, "pano_stretch: at most 65535 images per call");
    switch (C) {
        case 1: stretch_kernel<1><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        case 2: stretch_kernel<2><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        case 3: stretch_kernel<3><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        default: stretch_kernel<4><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
    }
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
