// tcgen05 weight gradient of a convolution (training step; reference train.py:278 `loss.backward()`
// through every nn.Conv2d of model.py:73-81, 129):
//     dW[co][dy][dx][ci] = sum over output pixels (b, ho, wo) of  dz[b][ho][wo][co] * in[b][ho*sh - ph + dy][wo*sw - pw + dx][ci]
// i.e. a GEMM  D[Cout x Cin] = dz^T * in_shifted  per filter tap whose reduction dimension is the PIXEL axis.  Both
// operands are halo-NHWC split planes (fp16 hi / lo of the 2^-4-scaled value, conv_tc.cuh; dz additionally scaled by a
// power of two, split_planes_pow2), so a TMA box of {64 channels, pixels} lands in shared memory as 128-byte rows, one
// per pixel: exactly the forward kernel's activation tile.  Read as an **MN-major** UMMA operand (instruction-descriptor
// bits 15/16; canonical 128B-swizzle layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 128 B of channels
// contiguous, 8 pixel rows per swizzle atom, SBO = 1024 B between 8-pixel groups, LBO = the distance between two
// 64-channel atoms) the same tile is the [channels x pixels] operand the weight gradient needs -- no transposed copy
// of any activation is ever made.
//
// One CTA = one work item (128 output channels x <=128 input channels x one tap x one slice of the pixel axis):
//   warp 0     TMA producer: per 64-pixel K tile the dz atoms (hi, lo) and the tap-shifted input atoms (hi, lo); the
//              pixel tile is {tw columns, rows, images} like the forward kernel's row boxes (zero H padding = TMA
//              out-of-bounds fill, circular W padding = the halo column, stride 2 along H = TMA traversal stride,
//              stride 2 along W = the parity view); 3-stage mbarrier ring of 64 KB stages
//   warp 1     MMA issuer: 4 K-steps of 16 pixels x 3 products (hi*hi -> main accumulator; hi*lo, lo*hi -> cross
//              accumulator) per tile, tcgen05.commit frees the stage
//   warp 2     TMEM allocator (2 x 128 columns)
//   warps 4-11 epilogue: tcgen05.ld, main + cross, power-of-two rescale, vector fp32 reductions (red.global.add.v4)
//              into dW[Cout][taps][Cin] (several pixel slices and the CTAs of different images add into the same tile)
#include <cuda.h>
#include <cstdlib>
#include <cstring>

#include "hn_common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"
#include "bwd_kernels.cuh"

namespace hn {

namespace {

using namespace tc;

constexpr int KT = 64;                    // pixels per pipeline stage (the GEMM's K dimension)
constexpr int ATOM = KT * 128;            // one operand atom: 64 pixels x 64 fp16 channels = 8 KB
constexpr int PLANE = 2 * ATOM;           // up to two 64-channel atoms per operand plane
constexpr int STAGE = 4 * PLANE;          // dz hi, dz lo, in hi, in lo = 64 KB
constexpr int NST = 3;
constexpr int BAR_OFF = NST * STAGE;
constexpr int SMEM_TOTAL = BAR_OFF + 256 + 1024;       // barriers + alignment slack
constexpr int ACC_COLS = 128;             // columns per accumulator
constexpr int TMEM_COLS = 2 * ACC_COLS;   // hi*hi + cross products
constexpr int EPI_WARP0 = 4;

struct WgArgs {
    float* dw;                // [Cout][taps][Cin], accumulated
    const float* absmax;      // of dz: its planes hold dz * pow2_factor(*absmax) * 2^-4
    int Cout, Cin, taps, kw;
    int n_tiles;              // 128-channel tiles along Cin (the 128-channel tiles along Cout come from the grid size)
    int tw, rpt, wsegs, Ho;   // K tile = rpt output rows x tw columns (rows of one image, or rpt / Ho whole images)
    int out_halo;
    int sh, ph, woff, parity; // woff = in_halo - pw; parity = 1: stride 2 along W through the [Wp/2][2] view
    int num_kt, kt_per_slice, items_per_slice;
    uint32_t lbo, sbo;        // operand descriptor offsets: ATOM between 64-channel atoms, 1024 between 8-pixel groups
};

// MN-major, 128-byte swizzle shared-memory operand descriptor: rows of 128 B (64 channels of one pixel), 8-pixel
// groups 1024 B apart (stride byte offset), 64-channel atoms ATOM bytes apart (leading byte offset)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address        bits [0,14)
    d |= (uint64_t)(lbo >> 4) << 16;                      // leading byte offset  bits [16,30)
    d |= (uint64_t)(sbo >> 4) << 32;                      // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor, fp16 x fp16 -> fp32, A and B both MN-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_mn(int n) {
    return umma_idesc(n, 0, 0, BM) | (1u << 15) | (1u << 16);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,
                const __grid_constant__ CUtensorMap tmIh, const __grid_constant__ CUtensorMap tmIl, const WgArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* done_bar = empty_bar + NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- which work item is this CTA's?  Items of one pixel slice are adjacent in the grid: they share their operand
    // tiles through L2.
    const int slice = (int)blockIdx.x / a.items_per_slice;
    int rem = (int)blockIdx.x - slice * a.items_per_slice;
    const int tap = rem % a.taps;
    rem /= a.taps;
    const int nt = rem % a.n_tiles, mt = rem / a.n_tiles;
    const int co0 = mt * 128, ci0 = nt * 128;
    const int m_atoms = min(2, (a.Cout - co0) / 64), n_atoms = min(2, (a.Cin - ci0) / 64);
    const int kt0 = slice * a.kt_per_slice, kt1 = min(a.num_kt, kt0 + a.kt_per_slice);
    const int dy = tap / a.kw, dx = tap - dy * a.kw;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmZh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmZl)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmIh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmIl)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(2 * (m_atoms + n_atoms) * ATOM);
            for (int kt = kt0; kt < kt1; ++kt) {
                const int rg = kt / a.wsegs;
                const int wo0 = (kt - rg * a.wsegs) * a.tw;
                const int row0 = rg * a.rpt;
                const int b = row0 / a.Ho, ho = row0 - b * a.Ho;
                const int hin = ho * a.sh + dy - a.ph;
                mbar_wait(empty_bar + stage, phase ^ 1);
                uint8_t* sZ = smem + stage * STAGE;
                uint8_t* sI = sZ + 2 * PLANE;
                mbar_expect_tx(full_bar + stage, bytes);
                for (int j = 0; j < m_atoms; ++j) {
                    tma_load_4d(sZ + j * ATOM, &tmZh, full_bar + stage, co0 + 64 * j, wo0 + a.out_halo, ho, b);
                    tma_load_4d(sZ + PLANE + j * ATOM, &tmZl, full_bar + stage, co0 + 64 * j, wo0 + a.out_halo, ho, b);
                }
                for (int j = 0; j < n_atoms; ++j) {
                    if (a.parity) {
                        const int p = dx + a.woff;
                        tma_load_5d(sI + j * ATOM, &tmIh, full_bar + stage, ci0 + 64 * j, p & 1, wo0 + (p >> 1), hin, b);
                        tma_load_5d(sI + PLANE + j * ATOM, &tmIl, full_bar + stage, ci0 + 64 * j, p & 1, wo0 + (p >> 1), hin, b);
                    } else {
                        tma_load_4d(sI + j * ATOM, &tmIh, full_bar + stage, ci0 + 64 * j, wo0 + dx + a.woff, hin, b);
                        tma_load_4d(sI + PLANE + j * ATOM, &tmIl, full_bar + stage, ci0 + 64 * j, wo0 + dx + a.woff, hin, b);
                    }
                }
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // hi*hi into the main accumulator, the two small cross products (2^-11 of the result) into their own: tcgen05
        // accumulates with truncation (conv_tc.cu), and a slice is at most 64 tiles = 256 accumulation steps of the main sum
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_mn(64 * n_atoms);
            const uint32_t d_main = tmem_base, d_cross = tmem_base + ACC_COLS;
            int stage = 0;
            uint32_t phase = 0;
            for (int kt = kt0; kt < kt1; ++kt) {
                mbar_wait(full_bar + stage, phase);
                tc_fence_after();
                const uint32_t sZ = smem_u32(smem + stage * STAGE);
                const uint32_t sI = sZ + 2 * PLANE;
                const uint64_t z_hi = umma_desc_mn_sw128(sZ, a.lbo, a.sbo), z_lo = umma_desc_mn_sw128(sZ + PLANE, a.lbo, a.sbo);
                const uint64_t i_hi = umma_desc_mn_sw128(sI, a.lbo, a.sbo), i_lo = umma_desc_mn_sw128(sI + PLANE, a.lbo, a.sbo);
#pragma unroll
                for (int k = 0; k < KT / 16; ++k) {
                    const uint64_t ko = (uint64_t)((k * 16 * 128) >> 4);     // 16 pixel rows of 128 B further on
                    const uint32_t acc = (kt != kt0 || k != 0) ? 1u : 0u;      // the slice's first K-step overwrites
                    umma_f16(d_main, z_hi + ko, i_hi + ko, idesc, acc);
                    umma_f16(d_cross, z_hi + ko, i_lo + ko, idesc, acc);
                    umma_f16(d_cross, z_lo + ko, i_hi + ko, idesc, 1);
                }
                umma_commit(empty_bar + stage);               // frees the smem stage when the MMAs retire
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);                            // both accumulators complete
        }
    } else if (warp >= EPI_WARP0) {
        // =============================== epilogue ===============================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - EPI_WARP0) >> 2;        // column half
        const int row = q * 32 + lane;                   // accumulator row = output channel of the tile
        const bool valid = row < 64 * m_atoms;
        const int cols = 32 * n_atoms;                   // columns per warp (the two warps of a quarter split N)
        const float factor = 256.f / pow2_factor(__ldg(a.absmax));     // both operands carry 2^-4; dz also the power of two
        float* drow = a.dw + ((size_t)(co0 + (valid ? row : 0)) * a.taps + tap) * a.Cin + ci0 + half * cols;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        if (kt1 > kt0) {
            for (int c0 = 0; c0 < cols; c0 += 32) {
                uint32_t vm[32], vc[32];
                tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(half * cols + c0), vm);
                tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(ACC_COLS + half * cols + c0), vc);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(drow + c0 + j, (__uint_as_float(vm[j]) + __uint_as_float(vc[j])) * factor,
                                   (__uint_as_float(vm[j + 1]) + __uint_as_float(vc[j + 1])) * factor,
                                   (__uint_as_float(vm[j + 2]) + __uint_as_float(vc[j + 2])) * factor,
                                   (__uint_as_float(vm[j + 3]) + __uint_as_float(vc[j + 3])) * factor);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

struct Geometry {
    int tw, rpt, wsegs, box_rows, imgs;
    long long num_kt;
};

// pixel tiling of the output (= dz) geometry: 64 pixels per tile as {tw columns, box_rows rows, imgs images}
bool tile_geometry(const Act& dz, Geometry* g) {
    const int tw = dz.W < KT ? dz.W : KT;
    if (tw < 8 || (tw & (tw - 1)) != 0 || dz.W % tw != 0) return false;
    const int rpt = KT / tw;
    g->tw = tw; g->rpt = rpt; g->wsegs = dz.W / tw;
    if (dz.H % rpt == 0) { g->box_rows = rpt; g->imgs = 1; }
    else if (rpt % dz.H == 0) { g->box_rows = dz.H; g->imgs = rpt / dz.H; }
    else return false;
    const long long rows = (long long)dz.B * dz.H;
    g->num_kt = ((rows + rpt - 1) / rpt) * g->wsegs;
    return true;
}

// pixel slices: at most 64 tiles (4096 pixels, 256 accumulation steps) and at least 8 tiles per CTA; among those the
// slice count whose CTA total fills whole waves of the SMs best (one CTA per SM), fewer slices on a tie (fewer reductions).
// Returns the tiles per slice, *slices_out = the slice count that follows from it.
int choose_slices(long long num_kt, long long items, int sms, long long* slices_out) {
    const long long smin = (num_kt + 63) / 64;
    long long smax = num_kt / 8;
    if (smax < smin) smax = smin;
    long long slices = smin;
    double best = 0.0;
    for (long long s = smin; s <= smax; ++s) {
        const long long total = items * s, waves = (total + sms - 1) / sms;
        const double fill = (double)total / (double)(waves * sms);
        if (fill > best + 0.02) { best = fill; slices = s; }
        if (total >= 16ll * sms) break;
    }
    const int kt_per_slice = (int)((num_kt + slices - 1) / slices);
    *slices_out = (num_kt + kt_per_slice - 1) / kt_per_slice;
    return kt_per_slice;
}

}  // namespace

bool wgrad_tc_on() {
    // HN_WGRAD_TC=0/1 selects the fp32 CUDA-core / the tcgen05 weight-gradient kernel for the convolutions it supports
    const char* e = getenv("HN_WGRAD_TC");
    return e ? atoi(e) != 0 : HN_WGRAD_TC_DEFAULT != 0;
}

bool conv_wgrad_tc_supported(const ConvDesc& d, const Act& in, const Act& dz) {
    if (d.Cin % 64 != 0 || d.Cout % 64 != 0) return false;
    if (d.pw > in.halo || in.halo > 1 || dz.halo > 1) return false;
    if (in.C != d.Cin || dz.C != d.Cout || in.B != dz.B) return false;
    if (d.sh != 1 && d.sh != 2) return false;
    if (d.sw != 1 && d.sw != 2) return false;
    if (d.sw == 2 && (in.Wp() % 2) != 0) return false;
    if ((in.H + 2 * d.ph - d.kh) / d.sh + 1 != dz.H || (in.W + 2 * d.pw - d.kw) / d.sw + 1 != dz.W) return false;
    Geometry g;
    if (!tile_geometry(dz, &g)) return false;
    if (g.box_rows * d.sh > 256 || g.imgs > 256) return false;
    return g.num_kt > 0 && g.num_kt < (1ll << 24);
}

// The host-side plan of conv_wgrad_tc for a shape (no device work; `sms` = SM count to plan for):
// plan[0..9] = tw, rows per tile, tiles per row, box rows, images per tile, pixel tiles, tiles per slice, slices,
// work items per slice (Cout tiles x Cin tiles x taps), CTAs.
int conv_wgrad_tc_plan(const ConvDesc& d, const Act& in, const Act& dz, int sms, int plan[10]) {
    HN_CHECK(conv_wgrad_tc_supported(d, in, dz), "conv_wgrad_tc: unsupported shape");
    HN_CHECK(sms >= 1, "conv_wgrad_tc_plan: sms must be positive");
    Geometry g;
    tile_geometry(dz, &g);
    const long long items = (long long)((d.Cout + 127) / 128) * ((d.Cin + 127) / 128) * d.kh * d.kw;
    long long slices = 0;
    const int kps = choose_slices(g.num_kt, items, sms, &slices);
    HN_CHECK(items * slices < (1ll << 31), "conv_wgrad_tc: too many work items");
    const int v[10] = {g.tw, g.rpt, g.wsegs, g.box_rows, g.imgs, (int)g.num_kt, kps, (int)slices, (int)items, (int)(items * slices)};
    for (int i = 0; i < 10; ++i) plan[i] = v[i];
    return 0;
}

int conv_wgrad_tc(const ConvDesc& d, const Act& in, const unsigned short* in_planes, const Act& dz,
                  const unsigned short* dz_planes, const float* dz_absmax, float* dw_ohwi, cudaStream_t st) {
    HN_CHECK(conv_wgrad_tc_supported(d, in, dz), "conv_wgrad_tc: unsupported shape");
    Geometry g;
    tile_geometry(dz, &g);
    WgArgs a;
    memset(&a, 0, sizeof(a));
    a.dw = dw_ohwi; a.absmax = dz_absmax;
    a.Cout = d.Cout; a.Cin = d.Cin; a.taps = d.kh * d.kw; a.kw = d.kw;
    a.n_tiles = (d.Cin + 127) / 128;
    const int m_tiles = (d.Cout + 127) / 128;
    a.tw = g.tw; a.rpt = g.rpt; a.wsegs = g.wsegs; a.Ho = dz.H; a.out_halo = dz.halo;
    a.sh = d.sh; a.ph = d.ph; a.woff = in.halo - d.pw; a.parity = (d.sw == 2);
    a.num_kt = (int)g.num_kt;
    // Descriptor convention for MN-major 128B-swizzle operands as CUTLASS documents it (cute/atom/mma_traits_sm100.hpp,
    // make_umma_desc<Major::MN>): leading byte offset = stride between 64-element atoms along M/N, stride byte offset =
    // stride between 8-row groups along K.  HN_WGRAD_TC_DESC=1 swaps the two (bring-up switch).
    {
        const char* e = getenv("HN_WGRAD_TC_DESC");
        const bool swapped = e && atoi(e) == 1;
        a.lbo = swapped ? 1024u : (uint32_t)ATOM;
        a.sbo = swapped ? (uint32_t)ATOM : 1024u;
    }
    HN_CUDA_OK(cudaMemsetAsync(dw_ohwi, 0, (size_t)d.Cout * a.taps * d.Cin * sizeof(float), st));

    int dev = 0, sms = 0;
    HN_CUDA_OK(cudaGetDevice(&dev));
    HN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long items = (long long)m_tiles * a.n_tiles * a.taps;
    long long slices = 0;
    a.kt_per_slice = choose_slices(g.num_kt, items, sms, &slices);
    a.items_per_slice = (int)items;
    HN_CHECK(items * slices < (1ll << 31), "conv_wgrad_tc: too many work items");

    CUtensorMap tmZ[2], tmI[2];
    const size_t in_plane = in.numel(), dz_plane = dz.numel();
    for (int p = 0; p < 2; ++p) {
        {   // dz planes: {Cout, Wop, Ho, B}, interior columns only (the box starts at out_halo)
            const cuuint64_t C2 = (cuuint64_t)d.Cout * 2, Wp = dz.Wp();
            cuuint64_t dims[4] = {(cuuint64_t)d.Cout, Wp, (cuuint64_t)dz.H, (cuuint64_t)dz.B};
            cuuint64_t str[3] = {C2, C2 * Wp, C2 * Wp * dz.H};
            cuuint32_t box[4] = {64, (cuuint32_t)g.tw, (cuuint32_t)g.box_rows, (cuuint32_t)g.imgs};
            if (make_map(&tmZ[p], dz_planes + p * dz_plane, 4, dims, str, box)) return -1;
        }
        // input planes: the rows a tile's output rows read under tap dy are box_rows rows sh apart (traversal stride)
        const bool strided_rows = d.sh == 2 && g.box_rows * g.imgs > 1;
        const cuuint32_t boxrows = strided_rows ? (cuuint32_t)(g.box_rows * 2) : (cuuint32_t)g.box_rows;
        const cuuint64_t C2 = (cuuint64_t)d.Cin * 2, Wp = in.Wp();
        if (!a.parity) {
            cuuint64_t dims[4] = {(cuuint64_t)d.Cin, Wp, (cuuint64_t)in.H, (cuuint64_t)in.B};
            cuuint64_t str[3] = {C2, C2 * Wp, C2 * Wp * in.H};
            cuuint32_t box[4] = {64, (cuuint32_t)g.tw, boxrows, (cuuint32_t)g.imgs};
            if (make_map(&tmI[p], in_planes + p * in_plane, 4, dims, str, box, strided_rows ? 2 : -1, 2)) return -1;
        } else {
            cuuint64_t dims[5] = {(cuuint64_t)d.Cin, 2, Wp / 2, (cuuint64_t)in.H, (cuuint64_t)in.B};
            cuuint64_t str[4] = {C2, 2 * C2, C2 * Wp, C2 * Wp * in.H};
            cuuint32_t box[5] = {64, 1, (cuuint32_t)g.tw, boxrows, (cuuint32_t)g.imgs};
            if (make_map(&tmI[p], in_planes + p * in_plane, 5, dims, str, box, strided_rows ? 3 : -1, 2)) return -1;
        }
    }
    HN_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    wgrad_tc_kernel<<<(unsigned)(items * slices), NTHREADS, SMEM_TOTAL, st>>>(tmZ[0], tmZ[1], tmI[0], tmI[1], a);
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
