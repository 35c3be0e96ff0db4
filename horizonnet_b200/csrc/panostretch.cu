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
// Coordinates and the bilinear accumulation are fp64 like scipy (result cast to the image dtype); scipy's
// legacy 'wrap' folds coordinates with period n-1.  HBM-bound: 2*H*W*C*4 bytes per panorama; measured: bound by the L1
// data pipe (89.7 % of its wavefront peak: 4-byte gathers and stores at a 12-byte pixel stride), 0.43-0.50 of HBM peak.
// Round-2 negative result: a variant that staged each tile's source patch in shared memory (16-byte cp.async, LDS
// gathers; bit-identical, verified over the 49-pair grid) ran 0.38 ms vs 0.28 ms per 64 images -- the patch fill cannot
// overlap the gathers inside a block and every smem pass costs the same L1 wavefronts it saves; removed.
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
    double ty;
};

// floor() of a coordinate in [0, 2^31) without the conversion unit: adding 1.5 * 2^52 leaves round-to-nearest(c) in
// the low mantissa word; one compare turns round-to-nearest into floor.  (ncu, round 1: F2F/F2I/I2F/FRND on the XU pipe
// were the top stall reason of this kernel -- 23 XU instructions per pixel.)
constexpr double MAGIC_RN = 6755399441055744.0;
__device__ __forceinline__ int floor_nonneg(double c, double& fl) {
    const double s = c + MAGIC_RN;
    int i = __double2loint(s);
    fl = s - MAGIC_RN;
    if (fl > c) { i -= 1; fl -= 1.0; }
    return i;
}

template <int C>
__device__ __forceinline__ RowTaps resolve_row(double cy, int H, int W) {
    cy = legacy_wrap(cy, H);
    double fl;
    const int y0 = floor_nonneg(cy, fl);
    RowTaps t;
    t.ty = cy - fl;
    // index y0+1 == H only happens with weight exactly 0; fold it like scipy does (period n-1)
    const int y1 = (y0 + 1 > H - 1) ? (H > 1 ? y0 + 1 - (H - 1) : 0) : y0 + 1;
    t.r0 = y0 * W * C;
    t.r1 = y1 * W * C;
    return t;
}

template <int C>
__device__ __forceinline__ int nearest_row(double cy, int H, int W) {
    double fl;
    return min(floor_nonneg(legacy_wrap(cy, H) + 0.5, fl), H - 1) * W * C;
}

// grid: (ceil(ceil(W/2)/128), ceil(H/2), n); thread = column pair (x, W-1-x) x row pair (y, H-1-y).
// u(W-1-x) = -u(x) and v(H-1-y) = -v(y), so g is even in the column and v0 is odd in the row: one fp64
// atan serves four pixels.  ncu showed the kernel latency-bound (long-scoreboard stalls on the gathers, 30 %
// DRAM), so all taps of the four pixels (4 x 4 x C loads) are issued back to back before any blending.
template <int C, typename T>
__global__ void __launch_bounds__(128) stretch_kernel(const T* __restrict__ img, T* __restrict__ out,
                                                      const ColEntry* __restrict__ cols,
                                                      const double* __restrict__ tanv, int H, int W, int order) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    const int xm = W - 1 - x;
    if (x > xm) return;
    const size_t plane = (size_t)H * W * C;
    const T* src = img + (size_t)n * plane;
    T* dst = out + (size_t)n * plane;
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
        const int rn[2] = {nearest_row<C>(ry, H, W), nearest_row<C>((double)(H - 1) - ry, H, W)};
        for (int sd = 0; sd < nside; ++sd)
            for (int vt = 0; vt < nvert; ++vt)
#pragma unroll
                for (int c = 0; c < C; ++c) dst[orow[vt] + ocol[sd] + c] = __ldg(src + rn[vt] + e[sd].xn * C + c);
        return;
    }
    T tap[2][2][4][C];                           // [side][vert][tap][channel]
#pragma unroll
    for (int sd = 0; sd < 2; ++sd)
#pragma unroll
        for (int vt = 0; vt < 2; ++vt) {
            const T* p00 = src + rt[vt].r0 + e[sd].x0 * C;
            const T* p01 = src + rt[vt].r0 + e[sd].x1 * C;
            const T* p10 = src + rt[vt].r1 + e[sd].x0 * C;
            const T* p11 = src + rt[vt].r1 + e[sd].x1 * C;
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
            const double uy = 1.0 - ty, ux = 1.0 - tx;
            const double w00 = uy * ux, w01 = uy * tx, w10 = ty * ux, w11 = ty * tx;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // scipy accumulates (y0,x0) (y0,x1) (y1,x0) (y1,x1) in double with separately rounded products; the fused
                // chain differs from that by < 2^-52 relative, i.e. after the cast to fp32 in at most a last-bit rounding
                // flip (the contract is 1 fp32 ulp = 1.2e-7), and costs 4 instead of 7 fp64 instructions per channel
                double acc = w00 * (double)tap[sd][vt][0][c];
                acc = fma(w01, (double)tap[sd][vt][1][c], acc);
                acc = fma(w10, (double)tap[sd][vt][2][c], acc);
                acc = fma(w11, (double)tap[sd][vt][3][c], acc);
                dst[orow[vt] + ocol[sd] + c] = (T)acc;       // map_coordinates output dtype = input dtype
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// augment_kernel ("next" row f3): the image path of the reference's training augmentation, reference
// dataset.py:53 + :69-105 + :124, in ONE gather pass from the uint8 HWC panorama to the float32 CHW network input:
//     img = uint8 / 255 (float32)  ->  pano_stretch(kx, ky)  ->  flip  ->  roll(dx)  ->  img ** p  ->  CHW
// The reference materialises a float32 HWC image after every step in a DataLoader worker (~70 ms per image on one CPU
// core, dominated by pano_stretch).  Here output pixel (y, xo) is traced back through roll and flip to column xs of
// the stretched image, sampled bilinearly from the uint8 source (fp64 coordinates and blend, scipy legacy 'wrap'; the
// taps are read through a 256-entry table of double(float32(v) / 255f), so there is no per-tap conversion), raised to
// the power p in fp32 and written to the three channel planes.  As in stretch_kernel one atan serves the four mirror pixels.
// Algorithmic bytes per panorama: H*W*3 (uint8 in) + H*W*3*4 (float32 out).
struct AugParams {
    int flip;        // dataset.py:88-91
    int dx;          // dataset.py:95-98, 0 <= dx < W
    float gamma;     // dataset.py:102-105 exponent p; <= 0: no gamma
    int stretch;     // 0: no stretch (self.stretch False): kx, ky are not used
};

__global__ void __launch_bounds__(128) augment_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                      const ColEntry* __restrict__ cols,
                                                      const double* __restrict__ tanv,
                                                      const AugParams* __restrict__ params, int H, int W) {
    __shared__ double lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (double)__fdiv_rn((float)i, 255.f);   // dataset.py:53
    __syncthreads();
    // thread = column pair (xs, W-1-xs) of the STRETCHED image x row pair (y, H-1-y): one atan serves four pixels, as in
    // stretch_kernel; each stretched column is then carried to its output column through flip and roll
    const int xs0 = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    const int xs1 = W - 1 - xs0;
    if (xs0 > xs1) return;
    const AugParams pr = params[n];
    const size_t plane = (size_t)H * W;
    const unsigned char* src = img + (size_t)n * plane * 3;
    float* dst = out + (size_t)n * plane * 3;
    const int ym = H - 1 - y;
    const int nside = (xs1 == xs0) ? 1 : 2, nvert = (ym == y) ? 1 : 2;
    const int xsd[2] = {xs0, xs1};
    int xo[2];                                   // np.flip: st[j] -> flipped[W-1-j]; np.roll(.., dx): in[j] -> out[(j + dx) mod W]
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        int j = pr.flip ? W - 1 - xsd[sd] : xsd[sd];
        j += pr.dx;
        xo[sd] = j >= W ? j - W : j;
    }
    float val[2][2][3];                          // [side][vert][channel]
    if (!pr.stretch) {
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
#pragma unroll
            for (int vt = 0; vt < 2; ++vt) {
                const unsigned char* p = src + ((size_t)(vt ? ym : y) * W + xsd[sd]) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) val[sd][vt][c] = (float)lut[p[c]];
            }
    } else {
        const ColEntry* ce = cols + (size_t)n * W;
        const ColEntry e[2] = {ce[xs0], ce[xs1]};
        const double v0 = atan(tanv[y] * e[0].g);                                 // panostretch.py:93
        const double ry = (v0 / PI_D + 0.5) * (double)H - 0.5;                     // :96
        const RowTaps rt[2] = {resolve_row<3>(ry, H, W), resolve_row<3>((double)(H - 1) - ry, H, W)};
        unsigned char tap[2][2][4][3];
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
#pragma unroll
            for (int vt = 0; vt < 2; ++vt) {
                const unsigned char* p00 = src + rt[vt].r0 + e[sd].x0 * 3;
                const unsigned char* p01 = src + rt[vt].r0 + e[sd].x1 * 3;
                const unsigned char* p10 = src + rt[vt].r1 + e[sd].x0 * 3;
                const unsigned char* p11 = src + rt[vt].r1 + e[sd].x1 * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    tap[sd][vt][0][c] = __ldg(p00 + c); tap[sd][vt][1][c] = __ldg(p01 + c);
                    tap[sd][vt][2][c] = __ldg(p10 + c); tap[sd][vt][3][c] = __ldg(p11 + c);
                }
            }
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
#pragma unroll
            for (int vt = 0; vt < 2; ++vt) {
                const double ty = rt[vt].ty, tx = e[sd].tx;
                const double uy = 1.0 - ty, ux = 1.0 - tx;
                const double w00 = uy * ux, w01 = uy * tx, w10 = ty * ux, w11 = ty * tx;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double acc = w00 * lut[tap[sd][vt][0][c]];
                    acc = fma(w01, lut[tap[sd][vt][1][c]], acc);
                    acc = fma(w10, lut[tap[sd][vt][2][c]], acc);
                    acc = fma(w11, lut[tap[sd][vt][3][c]], acc);
                    val[sd][vt][c] = (float)acc;                                  // map_coordinates output dtype = float32
                }
            }
    }
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        if (sd >= nside) break;
#pragma unroll
        for (int vt = 0; vt < 2; ++vt) {
            if (vt >= nvert) break;
            const int yy = vt ? ym : y;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = val[sd][vt][c];
                // dataset.py:105 (float32 ** float32) for v in [0, 1], p > 0: exp2f(p * log2f(v)) with the full-precision (1 ulp)
                // functions is within 1e-7 of the correctly rounded power here and a third of powf's instructions
                if (pr.gamma > 0.f) v = (v > 0.f) ? exp2f(pr.gamma * log2f(v)) : 0.f;
                dst[(size_t)c * plane + (size_t)yy * W + xo[sd]] = v;              // dataset.py:124 HWC -> CHW
            }
        }
    }
}

}  // namespace

// img/out: n images [H][W][C] fp32 on the device; kx/ky: n doubles on the device;
// scratch: (4*n*W + H) doubles on the device (n*W ColEntry records of 32 bytes, then H doubles).
template <typename T>
static int pano_stretch_device_t(const T* img, T* out, int n, int H, int W, int C, const double* kx_dev,
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
    HN_CHECK(n <= 65535, "pano_stretch: at most 65535 images per call");
    dim3 g(((W + 1) / 2 + 127) / 128, (H + 1) / 2, n);
    switch (C) {
        case 1: stretch_kernel<1, T><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        case 2: stretch_kernel<2, T><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        case 3: stretch_kernel<3, T><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
        default: stretch_kernel<4, T><<<g, 128, 0, st>>>(img, out, cols, tanv, H, W, order); break;
    }
    HN_LAUNCH_OK();
    return 0;
}

int pano_stretch_device(const float* img, float* out, int n, int H, int W, int C, const double* kx_dev,
                        const double* ky_dev, double* scratch, int order, cudaStream_t st) {
    return pano_stretch_device_t<float>(img, out, n, H, W, C, kx_dev, ky_dev, scratch, order, st);
}
// float64 images (the reference's CLI feeds float64 0-255, misc/panostretch.py:171): taps and result stay double
int pano_stretch_device_f64(const double* img, double* out, int n, int H, int W, int C, const double* kx_dev,
                            const double* ky_dev, double* scratch, int order, cudaStream_t st) {
    return pano_stretch_device_t<double>(img, out, n, H, W, C, kx_dev, ky_dev, scratch, order, st);
}

// img: n uint8 images [H][W][3]; out: n float32 images [3][H][W]; kx/ky: n doubles on the device (entries of images without
// stretch must still be positive); params_dev: n AugParams {flip, dx, gamma, stretch}; scratch as for pano_stretch_device.
int augment_device(const unsigned char* img, float* out, int n, int H, int W, const double* kx_dev, const double* ky_dev,
                   const int* params_dev, double* scratch, cudaStream_t st) {
    HN_CHECK(n >= 0 && H >= 1 && W >= 1, "augment: bad geometry");
    if (n == 0) return 0;
    HN_CHECK((long long)H * W * 3 < (1ll << 31), "augment: image too large");
    HN_CHECK(n <= 65535 && H <= 2 * 65535, "augment: at most 65535 images per call");
    static_assert(sizeof(AugParams) == 16, "AugParams layout");
    ColEntry* cols = reinterpret_cast<ColEntry*>(scratch);
    double* tanv = scratch + 4 * (size_t)n * W;
    const int tot = (n * W > H) ? n * W : H;
    stretch_tables_kernel<<<(tot + 255) / 256, 256, 0, st>>>(kx_dev, ky_dev, cols, tanv, n, H, W);
    HN_LAUNCH_OK();
    dim3 g(((W + 1) / 2 + 127) / 128, (H + 1) / 2, n);
    augment_kernel<<<g, 128, 0, st>>>(img, out, cols, tanv, reinterpret_cast<const AugParams*>(params_dev), H, W);
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
