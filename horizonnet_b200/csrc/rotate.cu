// rotate_panorama ("next" row f4): reference misc/pano_lsd_align.py:125-171 rotatePanorama + :101-122 warpImageFast,
// the vanishing-point alignment warp of preprocess.py:65-66.  Same kernel family as pano_stretch (an equirectangular
// gather), with a 3x3 rotation instead of the stretch map:
//   target pixel (TX, TY) (1-based) -> angles (:136-137) -> unit vector (uv2xyzN, :76-79) -> R^-1 (:146) ->
//   angles (xyz2uvN, :53-69) -> source pixel (Px, Py) (:149-150) -> bilinear sample (scipy map_coordinates order 1,
//   mode 'constant') of the image padded by one pixel on every side (:156-168, including the quirk of :163: the right
//   half of the bottom padding row copies the image's FIRST row).
// Everything in fp64 like the reference (its padded image and result are float64).  The padded image is never
// materialised: pad(r, c) maps a padded index to a source pixel.
#include <cstdlib>
#include "hn_common.cuh"

namespace hn {

namespace {

constexpr double PI_R = 3.14159265358979323846;

struct RotArgs {
    double rinv[9];     // row-major inverse of R
    int H, W, C;
};

// padded (H+2) x (W+2) image index -> source pixel offset (row * W + col); pano_lsd_align.py:156-168
__device__ __forceinline__ int pad_index(int r, int c, int H, int W) {
    if (r >= 1 && r <= H) {
        const int cc = (c == 0) ? W - 1 : (c == W + 1 ? 0 : c - 1);             // :158-159 wrap columns
        return (r - 1) * W + cc;
    }
    const bool top = (r == 0);
    if (c == 0) return top ? 0 : (H - 1) * W;                                    // :165, :168
    if (c == W + 1) return top ? (W - 1) : (H - 1) * W + (W - 1);                // :167, :166
    // :160-163: both halves of the top row, and the left half of the bottom row, mirror column W - c of the first /
    // last image row; the right half of the bottom row (c > W/2) copies the FIRST row (reference quirk, kept)
    const int row = (top || c > W / 2) ? 0 : H - 1;
    return row * W + (W - c);
}

template <typename T>
__global__ void __launch_bounds__(128) rotate_kernel(const T* __restrict__ img, double* __restrict__ out, const RotArgs a) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= a.W) return;
    const int H = a.H, W = a.W, C = a.C;
    const size_t plane = (size_t)H * W * C;
    const T* src = img + (size_t)blockIdx.z * plane;
    double* dst = out + (size_t)blockIdx.z * plane + ((size_t)y * W + x) * C;
    const double angx = ((double)(x + 1) - (double)W / 2.0 - 0.5) / (double)W * PI_R * 2.0;      // :136
    const double angy = -((double)(y + 1) - (double)H / 2.0 - 0.5) / (double)H * PI_R;           // :137
    double sx, cx, sy, cy;
    sincos(angx, &sx, &cx);
    sincos(angy, &sy, &cy);
    const double n0 = cy * sx, n1 = cy * cx, n2 = sy;                                            // uv2xyzN, planeID 1
    const double o0 = a.rinv[0] * n0 + a.rinv[1] * n1 + a.rinv[2] * n2;                          // :146
    const double o1 = a.rinv[3] * n0 + a.rinv[4] * n1 + a.rinv[5] * n2;
    const double o2 = a.rinv[6] * n0 + a.rinv[7] * n1 + a.rinv[8] * n2;
    double nxy = sqrt(o0 * o0 + o1 * o1);                                                        // xyz2uvN :57
    if (nxy < 0.000001) nxy = 0.000001;
    const double nxyz = sqrt(o0 * o0 + o1 * o1 + o2 * o2);
    const double v = asin(o2 / nxyz);                                                            // :60
    double u = asin(o0 / nxy);                                                                   // :61
    if (o1 < 0.0 && u >= 0.0) u = PI_R - u;                                                      // :62-63
    else if (o1 < 0.0 && u <= 0.0) u = -PI_R - u;                                                // :64-65
    if (u != u) u = 0.0;                                                                         // :67
    const double px = (u + PI_R) / (2.0 * PI_R) * (double)W + 0.5;                               // :149
    const double py = (-v + PI_R / 2.0) / PI_R * (double)H + 0.5;                                // :150
    // warpImageFast: coordinates (py, px) are 0-based indices into the padded image
    const double fy = floor(py), fx = floor(px);
    const int r0 = (int)fy, c0 = (int)fx;
    const double ty = py - fy, tx = px - fx;
    const bool inside = (py >= 0.0) && (py <= (double)(H + 1)) && (px >= 0.0) && (px <= (double)(W + 1));
    const double w[4] = {(1.0 - ty) * (1.0 - tx), (1.0 - ty) * tx, ty * (1.0 - tx), ty * tx};
    int off[4];
    bool ok[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int r = r0 + (t >> 1), c = c0 + (t & 1);
        ok[t] = inside && r >= 0 && r <= H + 1 && c >= 0 && c <= W + 1;
        off[t] = ok[t] ? pad_index(r, c, H, W) * C : 0;
    }
    for (int ch = 0; ch < C; ++ch) {
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (ok[t]) acc = fma(w[t], (double)src[off[t] + ch], acc);
        dst[ch] = acc;
    }
}

}  // namespace

// img: n images [H][W][C] float32 (in_f64 = 0) or float64 (in_f64 = 1) on the device; out: n float64 images [H][W][C];
// rinv: 9 doubles (host), the inverse of the reference's R (pano_lsd_align.py:143-146), shared by all n images.
int rotate_panorama_device(const void* img, int in_f64, double* out, int n, int H, int W, int C, const double* rinv,
                           cudaStream_t st) {
    HN_CHECK(n >= 0 && H >= 2 && W >= 2 && C >= 1, "rotate_panorama: bad geometry");
    if (n == 0) return 0;
    HN_CHECK((long long)(H + 2) * (W + 2) * C < (1ll << 31), "rotate_panorama: image too large");
    HN_CHECK(n <= 65535 && H <= 65535, "rotate_panorama: at most 65535 images / rows per call");
    HN_CHECK(W % 2 == 0, "rotate_panorama: the reference's padding rule needs an even width (pano_lsd_align.py:160-163)");
    RotArgs a;
    for (int i = 0; i < 9; ++i) a.rinv[i] = rinv[i];
    a.H = H; a.W = W; a.C = C;
    dim3 g((W + 127) / 128, H, n);
    if (in_f64) rotate_kernel<double><<<g, 128, 0, st>>>(reinterpret_cast<const double*>(img), out, a);
    else rotate_kernel<float><<<g, 128, 0, st>>>(reinterpret_cast<const float*>(img), out, a);
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
