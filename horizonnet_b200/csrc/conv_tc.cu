// tcgen05 tensor-core convolution / GEMM for sm_100a: split precision, three products, fp32 accumulate.
//
// Why split precision: the contract is <= 1e-4 max-abs against the fp32 reference on random-init
// weights; single-pass bf16 (1.4e-2) or tf32 (1.5e-3) miss it (SURVEY hard-part 1), and a
// bf16+bf16 split with three products measured 1.1e-4 on the B200.  Every fp32 value
// is carried as two fp16 planes of its 2^-4-scaled value:  hi = fp16(s), lo = fp16(s - hi)
// (11 + 11 significant bits, same 4 B/element as fp32, see conv_tc.cuh), and the product is
//     Ahi*Bhi + Ahi*Blo + Alo*Bhi            (the dropped Alo*Blo term is 2^-22 relative)
// accumulated in fp32 tensor memory (segmented, see below): 2.4e-6 .. 3.9e-6 max-abs end to end.  The planes are
// written by the producing kernel's epilogue, so every operand tile is MMA-ready when TMA drops
// it into shared memory (no in-kernel conversion pass).
//
// Kernels in this file (all persistent, warp-specialised, one CTA per SM, 384 threads, launched with programmatic
// stream serialisation: griddepcontrol.launch_dependents at entry, griddepcontrol.wait after the prologue):
//   conv_tc_kernel<BN>  3x3 / strided / large-K 1x1 convs and the LSTM input projections
//   gemm_tc_kernel      1x1 stride-1 convs with K <= 256 (TMA-prefetched residual, in-place epilogue, TMA store)
//   bott_tc_kernel      layer1: conv2 (3x3, 64->64) + conv3 (1x1, 64->256, + identity) fused, intermediate in shared memory
//   stem_tc_kernel      the 7x7 stride-2 stem over packed pixel pairs (no-swizzle UMMA operand with overlapping rows)
//
// conv_tc_kernel:
//   warp 0     TMA producer: per 64-channel K chunk it loads A (hi, lo: 128 pixels x 64 ch, 128B swizzle) from the
//              halo-NHWC planes with ONE box per plane -- {64 ch, tw pixels, R rows} for multi-row tiles (traversal
//              stride 2 along H for the stride-(2,1) convs, whole images when an image has fewer rows than a tile);
//              the 3x3 taps are shifted box coordinates, zero H padding is TMA out-of-bounds fill, circular W padding
//              is the halo column, stride-2 in W reads a parity view -- and B (hi, lo: BN x 64 weights, K-major);
//              3-stage mbarrier ring of 64 KB stages (4 stages of 48 KB for BN <= 64).
//              "dxr" mode (3x3, one 128-pixel row per tile): one 130-pixel box per (dy, chunk) feeds all three dx taps
//              through UMMA descriptors whose start address is shifted by dx rows (the swizzle is a function of the
//              absolute shared-memory address) -- the large-K convs were bound by L2 -> SM fill (12.5 TB/s).
//   warp 1     MMA issuer: one thread issues 12 tcgen05.mma (3 products x 4 K-steps of 16) per chunk: hi*hi into a
//              segment accumulator that is restarted every `seg` chunks, hi*lo and lo*hi into a cross accumulator
//              (tcgen05 accumulates with truncation, see the comment at the issue loop); tcgen05.commit frees the
//              smem stage and publishes segments / tiles
//   warp 2     TMEM allocator (2 segment + 2 cross accumulators = 4*BN columns)
//   warps 4-11 epilogue: tcgen05.ld (lane quarter = warp%4, column half = (warp-4)/4), segment promotion in fp32
//              registers, folded BN scale/shift, residual add, ReLU, re-split into hi/lo planes (or fp32), coalesced
//              stores through a warp-private staging buffer incl. the circular halo columns; overlaps the next tile's MMAs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "hn_common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

namespace hn {

namespace {

using namespace tc;                // BM = 128 pixels per tile (UMMA M), NTHREADS, PTX wrappers, make_map, launch_tc
constexpr int BKC = 64;            // fp16 channels per K chunk = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int EPI_WARP0 = 4;
constexpr int STAGE_PITCH = 80;    // bytes per staged row: 64 B payload + 16 B pad (conflict-free 16-byte accesses)

struct TcArgs {
    int mode;                 // 0: GEMM rows (1x1 stride-1 conv over all halo-NHWC pixels, or plain GEMM); 1: conv rows
    int M;                    // mode 0: rows; mode 1: B*Ho output rows
    int Ho, Wo, Wop;          // mode 1 output geometry (Wop = Wo + 2*out_halo)
    int out_halo;
    int tw, rows_per_tile, wsegs;
    int sh, ph, parity, woff; // parity = 1: stride 2 along W through the [Wp/2][2] view; woff = in_halo - pw
    int kw, kc_per_tap, num_kc;
    int Bimg;                 // images per plane (plane p of image b sits at index p*Bimg + b of the outer TMA dim)
    int Cout, n_tiles, num_tiles;
    const float* scale;
    const float* shift;
    const unsigned short* res; // residual planes in the OUTPUT geometry (hi at res, lo at res + out_plane), or null
    unsigned short* out;       // split output planes (hi at out, lo at out + out_plane)
    float* out_f32;           // fp32 output instead of planes (LSTM projections)
    size_t out_plane;         // elements per plane
    int relu;
    int seg;                  // K chunks per hi*hi accumulation segment
    int dxr;                  // 1: 3x3 single-row tiles load each (dy, channel chunk) input row ONCE (130 pixels) and the three
                              //    dx taps read it through row-shifted UMMA descriptors (L2->SM traffic per tap: 64 -> 43 KB)
    int rowbox;               // 1: the rows_per_tile output rows of a tile come from consecutive input rows of one image -> one TMA box
    int dbg;                  // HN_TC_DBG experiment bits: 1 skip residual reads, 2 skip output stores, 4 skip epilogue math
    int pf;                   // gemm_tc_kernel: L2 prefetch distance (in tiles) of the residual / activation boxes, 0 = off
};

template <int BN>
struct Smem {
    static constexpr int NST = (BN <= 64) ? 4 : STAGES;   // 48 KB stages (BN <= 64) leave room for a 4th
    static constexpr int A_PLANE = BM * BKC * 2;          // 16 KB
    static constexpr int B_PLANE = BN * BKC * 2;
    static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
    static constexpr int EPI_OFF = NST * STAGE;                       // 8 warp-private epilogue staging buffers
    static constexpr int BAR_OFF = EPI_OFF + 8 * 32 * STAGE_PITCH;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;    // barriers + alignment slack
    static constexpr int TMEM_COLS = 4 * BN;        // 2 hi*hi segment accumulators + 2 cross accumulators (128..512)
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    using S = Smem<BN>;
    constexpr int NST = S::NST;
    pdl_trigger();
    const int tile0 = (int)blockIdx.x;
    const int tstride = (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tfull_bar = empty_bar + NST;       // [2] hi*hi segment accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;           // [2] hi*hi segment accumulator drained
    uint64_t* cempty_bar = tempty_bar + 2;          // [2] cross-product accumulator drained
    uint64_t* afull_bar = cempty_bar + 2;           // [2] dxr mode: input-row slot filled
    uint64_t* aempty_bar = afull_bar + 2;           // [2] dxr mode: the three taps have read the input-row slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 2);
    // dxr mode re-partitions the operand ring: two input-row slots (hi + lo plane, 136 rows of 128 B each, 130 used)
    // followed by NST weight-tile slots
    constexpr int AX_PLANE = 17 * 1024;
    constexpr int AX_SLOT = 2 * AX_PLANE;
    constexpr int BX_OFF = 2 * AX_SLOT;
    static_assert(BX_OFF + NST * 2 * S::B_PLANE <= NST * S::STAGE, "dxr layout does not fit the operand ring");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(afull_bar + i, 1); mbar_init(aempty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) {               // accumulator hand-back: the 8 epilogue warps arrive
            mbar_init(tfull_bar + i, 1);
            mbar_init(tempty_bar + i, 8);
            mbar_init(cempty_bar + i, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                              // everything below reads / writes activations

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0 && a.dxr) {
            // 3x3 conv, one output row of 128 pixels per tile: per (dy, channel chunk) ONE box of 130 input pixels,
            // then the three weight tiles of that (dy, chunk)
            int ast = 0, bst = 0;
            uint32_t aph = 0, bph = 0;
            const int ndyc = a.num_kc / a.kw;
            for (int tile = tile0; tile < a.num_tiles; tile += tstride) {
                const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
                const int rg = mt / a.wsegs;
                const int wo0 = (mt - rg * a.wsegs) * a.tw;
                const int b = rg / a.Ho, ho = rg - b * a.Ho;
                for (int dyc = 0; dyc < ndyc; ++dyc) {
                    const int dy = dyc / a.kc_per_tap, cc = dyc - dy * a.kc_per_tap;
                    const int hin = ho * a.sh + dy - a.ph;
                    mbar_wait(aempty_bar + ast, aph ^ 1);
                    uint8_t* sAx = smem + ast * AX_SLOT;
                    mbar_expect_tx(afull_bar + ast, 2u * 130u * 128u);
                    tma_load_4d(sAx, &tmA, afull_bar + ast, cc * BKC, wo0 + a.woff, hin, b);
                    tma_load_4d(sAx + AX_PLANE, &tmA, afull_bar + ast, cc * BKC, wo0 + a.woff, hin, a.Bimg + b);
                    if (++ast == 2) { ast = 0; aph ^= 1; }
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(empty_bar + bst, bph ^ 1);
                        uint8_t* sB = smem + BX_OFF + bst * 2 * S::B_PLANE;
                        mbar_expect_tx(full_bar + bst, 2u * S::B_PLANE);
                        const int kb = ((dy * a.kw + dx) * a.kc_per_tap + cc) * BKC;
                        tma_load_3d(sB, &tmB, full_bar + bst, kb, nt * BN, 0);
                        tma_load_3d(sB + S::B_PLANE, &tmB, full_bar + bst, kb, nt * BN, 1);
                        if (++bst == NST) { bst = 0; bph ^= 1; }
                    }
                }
            }
        } else if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < a.num_tiles; tile += tstride) {
                const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
                int valid_rows = 1, row0 = 0, wo0 = 0;
                if (a.mode == 1) {
                    const int rg = mt / a.wsegs;
                    wo0 = (mt - rg * a.wsegs) * a.tw;
                    row0 = rg * a.rows_per_tile;
                    valid_rows = min(a.rows_per_tile, a.M - row0);
                }
                // a row box always delivers the full tile (rows past the end are out-of-bounds zero fill, full byte count)
                const uint32_t a_bytes = (a.mode == 0) ? 2u * S::A_PLANE
                                                       : (uint32_t)(2 * (a.rowbox ? a.rows_per_tile : valid_rows) * a.tw * BKC * 2);
                for (int kc = 0; kc < a.num_kc; ++kc) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sA = smem + stage * S::STAGE;
                    uint8_t* sB = sA + 2 * S::A_PLANE;
                    const int tap = kc / a.kc_per_tap;
                    const int c0 = (kc - tap * a.kc_per_tap) * BKC;
                    mbar_expect_tx(full_bar + stage, a_bytes + 2u * S::B_PLANE);
                    if (a.mode == 0) {
                        tma_load_3d(sA, &tmA, full_bar + stage, c0, mt * BM, 0);
                        tma_load_3d(sA + S::A_PLANE, &tmA, full_bar + stage, c0, mt * BM, 1);
                    } else {
                        const int dy = tap / a.kw, dx = tap - dy * a.kw;
                        // rowbox: the tile's output rows map to consecutive input rows of one image: one box per plane
                        // (small per-row boxes cost tensor-pipe time: 31 % on the W=32 layers vs 55-60 % on W>=128)
                        const int nbox = a.rowbox ? 1 : valid_rows;
                        for (int rr = 0; rr < nbox; ++rr) {
                            const int R = row0 + rr;
                            const int b = R / a.Ho, ho = R - b * a.Ho;
                            const int hin = ho * a.sh + dy - a.ph;
                            uint8_t* dst = sA + rr * a.tw * (BKC * 2);
                            if (a.parity) {
                                const int p = dx + a.woff;
                                tma_load_5d(dst, &tmA, full_bar + stage, c0, p & 1, wo0 + (p >> 1), hin, b);
                                tma_load_5d(dst + S::A_PLANE, &tmA, full_bar + stage, c0, p & 1, wo0 + (p >> 1), hin,
                                            a.Bimg + b);
                            } else {
                                tma_load_4d(dst, &tmA, full_bar + stage, c0, wo0 + dx + a.woff, hin, b);
                                tma_load_4d(dst + S::A_PLANE, &tmA, full_bar + stage, c0, wo0 + dx + a.woff, hin,
                                            a.Bimg + b);
                            }
                        }
                    }
                    tma_load_3d(sB, &tmB, full_bar + stage, kc * BKC, nt * BN, 0);
                    tma_load_3d(sB + S::B_PLANE, &tmB, full_bar + stage, kc * BKC, nt * BN, 1);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // tcgen05 accumulates into fp32 TMEM with truncation (measured: -3.4e-6 mean relative bias after
        // 256 accumulation steps, tools/probe_tc_accum.py), which over K up to 18432 costs ~1e-4 end to
        // end.  So (1) the small cross products (hi*lo, lo*hi; 2^-11 of the result) get their own
        // accumulator, and (2) the hi*hi accumulator is restarted every `seg` K-chunks: the epilogue
        // warps drain each segment and add it to a register-resident fp32 sum with round-to-nearest.
        if (lane == 0 && a.dxr) {
            constexpr uint32_t idesc = umma_idesc(BN, 0, 0, BM);
            int ast = 0, bst = 0;
            uint32_t aph = 0, bph = 0;
            int it = 0, g = 0;
            const int ndyc = a.num_kc / a.kw;
            for (int tile = tile0; tile < a.num_tiles; tile += tstride, ++it) {
                const int cbuf = it & 1;
                mbar_wait(cempty_bar + cbuf, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_cross = tmem_base + (2 + cbuf) * BN;
                uint32_t d_main = tmem_base;
                int seg_pos = 0, mbuf = 0, kc = 0;
                for (int dyc = 0; dyc < ndyc; ++dyc) {
                    mbar_wait(afull_bar + ast, aph);
                    tc_fence_after();
                    const uint32_t sAx = smem_u32(smem + ast * AX_SLOT);
                    for (int dx = 0; dx < 3; ++dx, ++kc) {
                        if (seg_pos == 0) {
                            mbuf = g & 1;
                            mbar_wait(tempty_bar + mbuf, ((g >> 1) & 1) ^ 1);
                            tc_fence_after();
                            d_main = tmem_base + mbuf * BN;
                        }
                        mbar_wait(full_bar + bst, bph);
                        tc_fence_after();
                        const uint32_t sB = smem_u32(smem + BX_OFF + bst * 2 * S::B_PLANE);
                        // tap dx = the same 130-pixel row read from pixel dx on: start address + dx rows of 128 B.
                        // Measured: the 128-byte swizzle is applied to the absolute shared-memory address bits, so the
                        // shifted start needs NO descriptor base offset (setting it to (start >> 7) & 7 gives wrong results).
                        const uint64_t a_hi = umma_desc_sw128(sAx + dx * 128);
                        const uint64_t a_lo = umma_desc_sw128(sAx + AX_PLANE + dx * 128);
                        const uint64_t b_hi = umma_desc_sw128(sB), b_lo = umma_desc_sw128(sB + S::B_PLANE);
#pragma unroll
                        for (int k = 0; k < BKC / 16; ++k) {
                            const uint64_t ko = (uint64_t)((k * 16 * 2) >> 4);
                            umma_f16(d_main, a_hi + ko, b_hi + ko, idesc, (seg_pos | k) != 0);
                            umma_f16(d_cross, a_hi + ko, b_lo + ko, idesc, (kc | k) != 0);
                            umma_f16(d_cross, a_lo + ko, b_hi + ko, idesc, 1);
                        }
                        umma_commit(empty_bar + bst);
                        if (++seg_pos == a.seg || kc == a.num_kc - 1) {
                            umma_commit(tfull_bar + mbuf);
                            seg_pos = 0;
                            ++g;
                        }
                        if (++bst == NST) { bst = 0; bph ^= 1; }
                    }
                    umma_commit(aempty_bar + ast);           // the input-row slot is free once the three taps retire
                    if (++ast == 2) { ast = 0; aph ^= 1; }
                }
            }
        } else if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(BN, 0, 0, BM);       // fp16 x fp16 -> fp32
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, g = 0;
            for (int tile = tile0; tile < a.num_tiles; tile += tstride, ++it) {
                const int cbuf = it & 1;
                mbar_wait(cempty_bar + cbuf, ((it >> 1) & 1) ^ 1);   // epilogue drained this cross accumulator
                tc_fence_after();
                const uint32_t d_cross = tmem_base + (2 + cbuf) * BN;
                uint32_t d_main = tmem_base;
                int seg_pos = 0, mbuf = 0;
                for (int kc = 0; kc < a.num_kc; ++kc) {
                    if (seg_pos == 0) {
                        mbuf = g & 1;
                        mbar_wait(tempty_bar + mbuf, ((g >> 1) & 1) ^ 1);   // segment accumulator drained
                        tc_fence_after();
                        d_main = tmem_base + mbuf * BN;
                    }
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * S::STAGE);
                    const uint32_t sB = sA + 2 * S::A_PLANE;
                    const uint64_t a_hi = umma_desc_sw128(sA), a_lo = umma_desc_sw128(sA + S::A_PLANE);
                    const uint64_t b_hi = umma_desc_sw128(sB), b_lo = umma_desc_sw128(sB + S::B_PLANE);
#pragma unroll
                    for (int k = 0; k < BKC / 16; ++k) {
                        const uint64_t ko = (uint64_t)((k * 16 * 2) >> 4);     // advance 32 B inside the swizzle row
                        umma_f16(d_main, a_hi + ko, b_hi + ko, idesc, (seg_pos | k) != 0);
                        umma_f16(d_cross, a_hi + ko, b_lo + ko, idesc, (kc | k) != 0);
                        umma_f16(d_cross, a_lo + ko, b_hi + ko, idesc, 1);
                    }
                    umma_commit(empty_bar + stage);               // frees the smem stage when the MMAs retire
                    if (++seg_pos == a.seg || kc == a.num_kc - 1) {
                        umma_commit(tfull_bar + mbuf);            // segment (and, at the end, the tile) complete
                        seg_pos = 0;
                        ++g;
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // =============================== epilogue ===============================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - EPI_WARP0) >> 2;        // column half
        // BN >= 64: the two warps of a lane quarter split the columns; BN == 32: warp of half 0 takes all 32
        constexpr int COLS_PER_WARP = (BN >= 64) ? BN / 2 : BN;
        constexpr int NCHUNK = COLS_PER_WARP / 32;
        constexpr int CW = 32;                                           // columns per tcgen05.ld chunk
        const bool works = (BN >= 64) || half == 0;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint8_t* stage = smem + S::EPI_OFF + (warp - EPI_WARP0) * 32 * STAGE_PITCH;
        const int nseg = (a.num_kc + a.seg - 1) / a.seg;
        int it = 0, g = 0;
        for (int tile = tile0; tile < a.num_tiles; tile += tstride, ++it) {
            const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
            const int cbuf = it & 1;
            const int r = q * 32 + lane;                 // accumulator row = pixel of the tile
            // ---- where does this row live in the output?
            bool valid;
            size_t pix;                                  // pixel index in the output tensor (incl. halo columns)
            size_t halo_pix = 0;
            bool has_halo = false;
            if (a.mode == 0) {
                const long long m = (long long)mt * BM + r;
                valid = m < a.M;
                pix = (size_t)m;
            } else {
                const int rg = mt / a.wsegs;
                const int wo = (mt - rg * a.wsegs) * a.tw + (r % a.tw);
                const int R = rg * a.rows_per_tile + r / a.tw;
                valid = R < a.M;
                pix = (size_t)R * a.Wop + wo + a.out_halo;
                if (a.out_halo) {
                    if (wo == 0) { has_halo = true; halo_pix = (size_t)R * a.Wop + a.Wo + 1; }
                    else if (wo == a.Wo - 1) { has_halo = true; halo_pix = (size_t)R * a.Wop; }
                }
            }
            // ---- drain the hi*hi segments into registers (fp32 adds, round to nearest)
            float sum[COLS_PER_WARP];
#pragma unroll
            for (int j = 0; j < COLS_PER_WARP; ++j) sum[j] = 0.f;
            for (int sgi = 0; sgi < nseg; ++sgi, ++g) {
                const int mbuf = g & 1;
                mbar_wait(tfull_bar + mbuf, (g >> 1) & 1);
                tc_fence_after();
                if (works) {
#pragma unroll
                    for (int ch = 0; ch < NCHUNK; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + lane_base + (uint32_t)(mbuf * BN + half * COLS_PER_WARP + ch * 32), v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum[ch * 32 + j] += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + mbuf);
            }
            // ---- the last segment's commit also covers the cross products of the whole tile.
            // Global traffic of the epilogue is routed through a warp-private staging buffer so that
            // every load/store instruction touches whole 64-byte row segments (4 lanes x 16 B per pixel,
            // 8 pixels per instruction) instead of 32 scattered 16-byte pieces.
            unsigned long long rowpix[4], rowhalo[4];           // pixel / halo pixel of row it*8 + lane/4
            {
                const unsigned long long mypix = valid ? (unsigned long long)pix : ~0ull;
                const unsigned long long myhalo = (valid && has_halo) ? (unsigned long long)halo_pix : ~0ull;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    rowpix[i4] = __shfl_sync(0xffffffffu, mypix, i4 * 8 + (lane >> 2));
                    rowhalo[i4] = __shfl_sync(0xffffffffu, myhalo, i4 * 8 + (lane >> 2));
                }
            }
            const int seg16 = (lane & 3) * 16;
            uint8_t* my_row = stage + lane * STAGE_PITCH;
            // 64 bytes per row, this lane's row -> global, coalesced
            auto put64 = [&](const uint4& d0, const uint4& d1, const uint4& d2, const uint4& d3, uint8_t* base,
                             size_t pitch, size_t colb) {
                uint4* w = reinterpret_cast<uint4*>(my_row);
                w[0] = d0; w[1] = d1; w[2] = d2; w[3] = d3;
                __syncwarp();
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const uint4 v = *reinterpret_cast<const uint4*>(stage + (i4 * 8 + (lane >> 2)) * STAGE_PITCH + seg16);
                    if (rowpix[i4] != ~0ull) *reinterpret_cast<uint4*>(base + rowpix[i4] * pitch + colb + seg16) = v;
                    if (rowhalo[i4] != ~0ull) *reinterpret_cast<uint4*>(base + rowhalo[i4] * pitch + colb + seg16) = v;
                }
                __syncwarp();
            };
            // global -> 64 bytes of this lane's row, coalesced
            auto get64 = [&](const uint8_t* base, size_t pitch, size_t colb, uint4 (&d)[4]) {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (rowpix[i4] != ~0ull) v = __ldg(reinterpret_cast<const uint4*>(base + rowpix[i4] * pitch + colb + seg16));
                    *reinterpret_cast<uint4*>(stage + (i4 * 8 + (lane >> 2)) * STAGE_PITCH + seg16) = v;
                }
                __syncwarp();
                const uint4* rd = reinterpret_cast<const uint4*>(my_row);
                d[0] = rd[0]; d[1] = rd[1]; d[2] = rd[2]; d[3] = rd[3];
                __syncwarp();
            };
#pragma unroll
            for (int ch = 0; ch < (works ? NCHUNK : 0); ++ch) {
                const int col0 = half * COLS_PER_WARP + ch * 32;          // column inside the tile
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)((2 + cbuf) * BN + col0), v);
                const int n0 = nt * BN + col0;                            // output channel of v[0]
                float y[32];
#pragma unroll
                for (int j = 0; j < CW; j += 4) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + n0 + j));
                    const float4 sf = __ldg(reinterpret_cast<const float4*>(a.shift + n0 + j));
                    y[j + 0] = fmaf(sum[ch * 32 + j + 0] + __uint_as_float(v[j + 0]), sc.x, sf.x);
                    y[j + 1] = fmaf(sum[ch * 32 + j + 1] + __uint_as_float(v[j + 1]), sc.y, sf.y);
                    y[j + 2] = fmaf(sum[ch * 32 + j + 2] + __uint_as_float(v[j + 2]), sc.z, sf.z);
                    y[j + 3] = fmaf(sum[ch * 32 + j + 3] + __uint_as_float(v[j + 3]), sc.w, sf.w);
                }
                if (a.res && !(a.dbg & 1)) {
                    uint4 rh[4], rl[4];
                    const size_t pitch = (size_t)a.Cout * 2, colb = (size_t)n0 * 2;
                    get64(reinterpret_cast<const uint8_t*>(a.res), pitch, colb, rh);
                    get64(reinterpret_cast<const uint8_t*>(a.res + a.out_plane), pitch, colb, rl);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t hw[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w}, lw[4] = {rl[j].x, rl[j].y, rl[j].z, rl[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {       // y is in plane units here: no rescaling needed
                            const float2 fh = unpack_half2(hw[e]), fl = unpack_half2(lw[e]);
                            y[j * 8 + 2 * e + 0] += fh.x + fl.x;
                            y[j * 8 + 2 * e + 1] += fh.y + fl.y;
                        }
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                if (a.dbg & 2) continue;
                if (a.out_f32) {
                    const size_t pitch = (size_t)a.Cout * 4;
                    uint8_t* base = reinterpret_cast<uint8_t*>(a.out_f32);
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const float* yy = y + hq * 16;
                        put64(make_uint4(__float_as_uint(yy[0]), __float_as_uint(yy[1]), __float_as_uint(yy[2]), __float_as_uint(yy[3])),
                              make_uint4(__float_as_uint(yy[4]), __float_as_uint(yy[5]), __float_as_uint(yy[6]), __float_as_uint(yy[7])),
                              make_uint4(__float_as_uint(yy[8]), __float_as_uint(yy[9]), __float_as_uint(yy[10]), __float_as_uint(yy[11])),
                              make_uint4(__float_as_uint(yy[12]), __float_as_uint(yy[13]), __float_as_uint(yy[14]), __float_as_uint(yy[15])),
                              base, pitch, (size_t)(n0 + hq * 16) * 4);
                    }
                } else {
                    uint32_t ph[16], pl[16];
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) split2_scaled(y[2 * j], y[2 * j + 1], ph[j], pl[j]);
                    const size_t pitch = (size_t)a.Cout * 2, colb = (size_t)n0 * 2;
                    put64(make_uint4(ph[0], ph[1], ph[2], ph[3]), make_uint4(ph[4], ph[5], ph[6], ph[7]),
                          make_uint4(ph[8], ph[9], ph[10], ph[11]), make_uint4(ph[12], ph[13], ph[14], ph[15]),
                          reinterpret_cast<uint8_t*>(a.out), pitch, colb);
                    put64(make_uint4(pl[0], pl[1], pl[2], pl[3]), make_uint4(pl[4], pl[5], pl[6], pl[7]),
                          make_uint4(pl[8], pl[9], pl[10], pl[11]), make_uint4(pl[12], pl[13], pl[14], pl[15]),
                          reinterpret_cast<uint8_t*>(a.out + a.out_plane), pitch, colb);
                }
            }
            // cross accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cempty_bar + cbuf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
    }
}

// ================================================================================================
// gemm_tc_kernel: 1x1 stride-1 convolutions in GEMM mode with K <= 256 (conv3 / conv1 of layer1-3: the
// bottlenecks).  These layers are bound by the epilogue, not by the MMAs: in conv_tc_kernel the residual
// loads and the output stores of a tile sit in the epilogue warps' dependency chain (measured: 4.7 of
// 13.3 ms of the encoder convs disappear when those memory instructions are removed, HN_TC_DBG).  Here
// all epilogue traffic is asynchronous: the producer warp TMA-loads the residual tile (hi, lo planes,
// 128 rows x 64 channels, 128-byte swizzle) into one of two epilogue buffers a whole tile ahead, the
// epilogue warps update it IN PLACE from TMEM (BN scale/shift, + residual, ReLU, re-split) and one thread
// hands the buffer to the TMA engine for the store.  Tiles are 128 x 64 so that a 3-stage operand ring
// (48 KB/stage) and the two 32 KB epilogue buffers fit into shared memory together.
constexpr int GBN = 64;

struct GSmem {
    static constexpr int A_PLANE = BM * BKC * 2;          // 16 KB
    static constexpr int B_PLANE = GBN * BKC * 2;         // 8 KB
    static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;            // 48 KB
    static constexpr int E_PLANE = BM * GBN * 2;          // 16 KB: 128 rows x 64 channels fp16
    static constexpr int EBUF = 2 * E_PLANE;              // hi + lo
    static constexpr int EPI_OFF = STAGES * STAGE;        // 144 KB
    static constexpr int BAR_OFF = EPI_OFF + 2 * EBUF;    // 208 KB
    static constexpr int CST_OFF = BAR_OFF + 256;         // epilogue constants scale[Cout] shift[Cout], Cout <= 1024
    static constexpr int MAX_COUT = 1024;
    static constexpr int TOTAL = CST_OFF + 2 * MAX_COUT * 4 + 1024;
    static constexpr int TMEM_COLS = 4 * GBN;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const TcArgs a) {
    using S = GSmem;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;       // [2] hi*hi segment accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;           // [2] hi*hi segment accumulator drained
    uint64_t* cempty_bar = tempty_bar + 2;          // [2] cross-product accumulator drained
    uint64_t* rfull_bar = cempty_bar + 2;           // [2] epilogue buffer holds the residual tile (or is simply free)
    uint64_t* efree_bar = rfull_bar + 2;            // [2] the TMA store has finished reading the epilogue buffer
    uint64_t* oready_bar = efree_bar + 2;           // [2] all 8 epilogue warps have written their part of the output tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oready_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_res = a.res != nullptr;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 8); mbar_init(cempty_bar + i, 8);
            mbar_init(rfull_bar + i, 1); mbar_init(efree_bar + i, 1); mbar_init(oready_bar + i, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // epilogue constants -> shared memory (the per-tile __ldg of scale / shift miss L1 after every fence.proxy.async: ncu showed
    // them as the top stall of the conv3 epilogue; weights-side data, safe to read before pdl_wait)
    float* cst = reinterpret_cast<float*>(smem + S::CST_OFF);
    for (int i = threadIdx.x; i < 2 * a.Cout; i += NTHREADS) cst[i] = (i < a.Cout) ? a.scale[i] : a.shift[i - a.Cout];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
                if (a.pf > 0) {                  // residual and activation boxes of the tile a.pf rounds ahead -> L2
                    const int ta = tile + a.pf * (int)gridDim.x;
                    if (ta < a.num_tiles) {
                        const int mta = ta / a.n_tiles, nta = ta - mta * a.n_tiles;
                        if (has_res) {
                            tma_prefetch_3d(&tmR, nta * GBN, mta * BM, 0);
                            tma_prefetch_3d(&tmR, nta * GBN, mta * BM, 1);
                        }
                        for (int kc = 0; kc < a.num_kc; ++kc) {
                            tma_prefetch_3d(&tmA, kc * BKC, mta * BM, 0);
                            tma_prefetch_3d(&tmA, kc * BKC, mta * BM, 1);
                        }
                    }
                }
                // operands first: the MMAs of this tile overlap the epilogue of the previous one
                for (int kc = 0; kc < a.num_kc; ++kc) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sA = smem + stage * S::STAGE;
                    uint8_t* sB = sA + 2 * S::A_PLANE;
                    mbar_expect_tx(full_bar + stage, 2u * S::A_PLANE + 2u * S::B_PLANE);
                    tma_load_3d(sA, &tmA, full_bar + stage, kc * BKC, mt * BM, 0);
                    tma_load_3d(sA + S::A_PLANE, &tmA, full_bar + stage, kc * BKC, mt * BM, 1);
                    tma_load_3d(sB, &tmB, full_bar + stage, kc * BKC, nt * GBN, 0);
                    tma_load_3d(sB + S::B_PLANE, &tmB, full_bar + stage, kc * BKC, nt * GBN, 1);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // epilogue buffer of this tile: wait until its previous store is done, then load the residual
                const int eb = it & 1;
                mbar_wait(efree_bar + eb, ((it >> 1) & 1) ^ 1);
                uint8_t* ebuf = smem + S::EPI_OFF + eb * S::EBUF;
                if (has_res) {
                    mbar_expect_tx(rfull_bar + eb, 2u * S::E_PLANE);
                    tma_load_3d(ebuf, &tmR, rfull_bar + eb, nt * GBN, mt * BM, 0);
                    tma_load_3d(ebuf + S::E_PLANE, &tmR, rfull_bar + eb, nt * GBN, mt * BM, 1);
                } else {
                    mbar_arrive(rfull_bar + eb);
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (same scheme as conv_tc_kernel) ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(GBN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, g = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int cbuf = it & 1;
                mbar_wait(cempty_bar + cbuf, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_cross = tmem_base + (2 + cbuf) * GBN;
                uint32_t d_main = tmem_base;
                int seg_pos = 0, mbuf = 0;
                for (int kc = 0; kc < a.num_kc; ++kc) {
                    if (seg_pos == 0) {
                        mbuf = g & 1;
                        mbar_wait(tempty_bar + mbuf, ((g >> 1) & 1) ^ 1);
                        tc_fence_after();
                        d_main = tmem_base + mbuf * GBN;
                    }
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * S::STAGE);
                    const uint32_t sB = sA + 2 * S::A_PLANE;
                    const uint64_t a_hi = umma_desc_sw128(sA), a_lo = umma_desc_sw128(sA + S::A_PLANE);
                    const uint64_t b_hi = umma_desc_sw128(sB), b_lo = umma_desc_sw128(sB + S::B_PLANE);
#pragma unroll
                    for (int k = 0; k < BKC / 16; ++k) {
                        const uint64_t ko = (uint64_t)((k * 16 * 2) >> 4);
                        umma_f16(d_main, a_hi + ko, b_hi + ko, idesc, (seg_pos | k) != 0);
                        umma_f16(d_cross, a_hi + ko, b_lo + ko, idesc, (kc | k) != 0);
                        umma_f16(d_cross, a_lo + ko, b_hi + ko, idesc, 1);
                    }
                    umma_commit(empty_bar + stage);
                    if (++seg_pos == a.seg || kc == a.num_kc - 1) {
                        umma_commit(tfull_bar + mbuf);
                        seg_pos = 0;
                        ++g;
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // =============================== epilogue ===============================
        const int q = warp & 3;                          // TMEM lane quarter
        const int half = (warp - EPI_WARP0) >> 2;        // column half: 32 of the 64 columns
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nseg = (a.num_kc + a.seg - 1) / a.seg;
        const int r = q * 32 + lane;                     // row of the tile owned by this thread
        int it = 0, g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
            const int cbuf = it & 1, eb = it & 1;
            float sum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = 0.f;
            for (int sgi = 0; sgi < nseg; ++sgi, ++g) {
                const int mbuf = g & 1;
                mbar_wait(tfull_bar + mbuf, (g >> 1) & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(mbuf * GBN + half * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(v[j]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + mbuf);
            }
            {   // cross products, then give both accumulators back before the math
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)((2 + cbuf) * GBN + half * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(v[j]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(cempty_bar + cbuf);
            }
            const int n0 = nt * GBN + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(cst + n0 + j);
                const float4 sf = *reinterpret_cast<const float4*>(cst + a.Cout + n0 + j);
                sum[j + 0] = fmaf(sum[j + 0], sc.x, sf.x);
                sum[j + 1] = fmaf(sum[j + 1], sc.y, sf.y);
                sum[j + 2] = fmaf(sum[j + 2], sc.z, sf.z);
                sum[j + 3] = fmaf(sum[j + 3], sc.w, sf.w);
            }
            // epilogue buffer: row r, 16-byte chunk c lives at r*128 + ((c ^ (r & 7)) << 4)  (128-byte swizzle)
            mbar_wait(rfull_bar + eb, (it >> 1) & 1);
            const uint32_t e_hi = smem_u32(smem + S::EPI_OFF + eb * S::EBUF) + r * 128;
            const uint32_t e_lo = e_hi + S::E_PLANE;
            uint4 rh[4], rl[4];                          // all residual loads in flight before the first use
            if (has_res) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                    rh[c] = ld_shared_v4(e_hi + off);
                    rl[c] = ld_shared_v4(e_lo + off);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t off = (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                float* y = sum + c * 8;
                if (has_res) {
                    const uint32_t hw[4] = {rh[c].x, rh[c].y, rh[c].z, rh[c].w}, lw[4] = {rl[c].x, rl[c].y, rl[c].z, rl[c].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 fh = unpack_half2(hw[e]), fl = unpack_half2(lw[e]);
                        y[2 * e + 0] += fh.x + fl.x;
                        y[2 * e + 1] += fh.y + fl.y;
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
                }
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2_scaled(y[2 * e], y[2 * e + 1], ph[e], pl[e]);
                st_shared_v4(e_hi + off, make_uint4(ph[0], ph[1], ph[2], ph[3]));
                st_shared_v4(e_lo + off, make_uint4(pl[0], pl[1], pl[2], pl[3]));
            }
            // make the generic-proxy writes visible to the TMA engine and tell the store warp
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(oready_bar + eb);
        }
    } else if (warp == 3) {
        // =============================== store warp ===============================
        // one thread hands finished output tiles to the TMA engine and returns the buffer to the producer
        // as soon as the engine has read it, so the residual of tile it+2 is prefetched a whole tile ahead
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
                const int eb = it & 1;
                mbar_wait(oready_bar + eb, (it >> 1) & 1);
                const uint8_t* ebuf = smem + S::EPI_OFF + eb * S::EBUF;
                tma_store_3d(&tmO, ebuf, nt * GBN, mt * BM, 0);
                tma_store_3d(&tmO, ebuf + S::E_PLANE, nt * GBN, mt * BM, 1);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(efree_bar + eb);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // all stores have landed
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
    }
}


// ================================================================================================
// stem_tc_kernel: the 7x7 stride-2 stem (3 -> 64 channels, model.py:73-75) as an implicit GEMM on tcgen05.
//
// Cin = 3 does not give a K-major operand by itself, so a pre-pass (stem_pack_kernel) writes the normalised
// input as split planes of PIXEL PAIRS:  P[plane][b][h][j][8] fp16,  element e of pair j = channel e&3 (channel 3
// is zero) of input column 2j + (e>>2) - 3 (circular in W), h = input row + 3 (rows outside the image are zero:
// the zero H padding of the normalised input).  With K ordered (dy, pair p, e) -- i.e. an 8-wide dx window whose
// 8th tap has zero weight -- the A row of output pixel xo for (dy, p) is pair xo + p of row 2*yo + dy: for a tile
// of 128 consecutive output pixels all four p are the SAME shared-memory bytes shifted by p*16 B.  That is exactly
// the no-swizzle K-major UMMA layout (core matrix = 8 rows x 16 B contiguous, SBO = 128 B) with a leading-dimension
// offset of one row (LBO = 16 B): a tile needs ONE TMA box per plane (7 rows x 136 pairs, 15 KB) and the MMAs read
// overlapping operands straight out of it.  K = 7 x 32 = 224 (147 useful), 14 K-steps x 3 products per tile.
// The weights ([2][64][224] fp16 planes, 128-byte swizzle) stay resident in shared memory for the whole kernel.
// Epilogue: TMEM -> BN scale/shift -> ReLU -> fp32 [128 px][64 ch] tile in shared memory (2 x 16 KB, swizzled)
// -> two TMA stores into the halo-NHWC fp32 stem output (the max-pool does not read the halo columns).
constexpr int SP_ROWS = 517;                 // input rows -3 .. 513
constexpr int SP_PAIRS = 520;                // pixel pairs per row (516 used, padded to a multiple of 8)
constexpr int SP_K = 224;                    // 7 dy x 4 pairs x 8
constexpr float STEM_IN_SCALE = 16.f;        // planes hold 16 * normalised input: keeps the lo plane a normal fp16

struct StSmem {
    static constexpr int B_CHUNK = 64 * BKC * 2;              // 8 KB: 64 couts x 64 k
    static constexpr int B_BYTES = 2 * 4 * B_CHUNK;           // hi/lo x 4 chunks = 64 KB
    static constexpr int A_ROW = 136 * 16;                    // one dy: 136 pairs x 16 B
    static constexpr int A_PLANE = 7 * A_ROW;                 // 15232 B
    static constexpr int A_STAGE = 30 * 1024;                 // hi + lo (30464 B) rounded up
    static constexpr int NSTAGE = 3;
    static constexpr int A_OFF = B_BYTES;
    static constexpr int E_HALF = 128 * 128;                  // 128 px x 32 ch fp32
    static constexpr int EBUF = 2 * E_HALF;
    static constexpr int EPI_OFF = A_OFF + NSTAGE * A_STAGE;  // 154 KB
    static constexpr int BAR_OFF = EPI_OFF + 2 * EBUF;        // 218 KB
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 128;                     // two 64-column accumulators
};

struct StemArgs {
    int B, num_tiles;
    const float* scale;       // accumulator -> true units for an input scaled by 2^-4 (tc_aux[0..64))
    const float* shift;       // folded BN shift
    float floor;              // 0: ReLU (model.py:75); -inf: none
};

// K-major operand without swizzle: 8-row x 16-byte core matrices, `lbo` bytes between the two core matrices of a
// K=16 step, `sbo` bytes between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_interleave(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100); layout type 0 = no swizzle
    return d;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}

__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x, int Cx,
                                                        unsigned short* __restrict__ out, int B) {
    const size_t total = (size_t)B * SP_ROWS * SP_PAIRS;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int j = (int)(i % SP_PAIRS);
    const size_t t = i / SP_PAIRS;
    const int h = (int)(t % SP_ROWS);
    const int b = (int)(t / SP_ROWS);
    const int row = h - 3;
    uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
    if (row >= 0 && row < 512 && j < 516) {
        const float mean[3] = {0.485f, 0.456f, 0.406f};       // reference model.py:186
        const float stdv[3] = {0.229f, 0.224f, 0.225f};       // reference model.py:187
        unsigned short hs[8], ls[8];
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
            int px = 2 * j + e2 - 3;
            px = px < 0 ? px + 1024 : (px >= 1024 ? px - 1024 : px);        // circular W (model.py:27-29)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = 0.f;
                if (c < 3) v = (__ldg(x + (((size_t)b * Cx + c) * 512 + row) * 1024 + px) - mean[c]) / stdv[c] * STEM_IN_SCALE;
                const __half hh = __float2half_rn(v);
                hs[e2 * 4 + c] = __half_as_ushort(hh);
                ls[e2 * 4 + c] = __half_as_ushort(__float2half_rn(v - __half2float(hh)));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { hi[q] = pack2(hs[2 * q], hs[2 * q + 1]); lo[q] = pack2(ls[2 * q], ls[2 * q + 1]); }
    }
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + (total + i) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// OIHW [64][3][7][7] -> OIHW [64][4][7][8] with zero 4th channel / 8th column (so that pack_weight_tc's K order
// (dy, dx, c) becomes the kernel's (dy, pair, e) order)
__global__ void stem_weight_pad_kernel(const float* __restrict__ w, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 4 * 7 * 8) return;
    const int dx = i % 8, dy = (i / 8) % 7, c = (i / 56) % 4, n = i / 224;
    out[i] = (c < 3 && dx < 7) ? w[((n * 3 + c) * 7 + dy) * 7 + dx] : 0.f;
}

__global__ void __launch_bounds__(NTHREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const StemArgs a) {
    using S = StSmem;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::NSTAGE;
    uint64_t* bfull_bar = empty_bar + S::NSTAGE;   // weights resident
    uint64_t* tfull_bar = bfull_bar + 1;           // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained
    uint64_t* oready_bar = tempty_bar + 2;         // [2] output tile staged by all 8 epilogue warps
    uint64_t* efree_bar = oready_bar + 2;          // [2] the TMA store has read the staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(efree_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S::NSTAGE; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(bfull_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 8);
            mbar_init(oready_bar + i, 8); mbar_init(efree_bar + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // tile -> (image b, output row yo, 128-pixel segment sg): 4 segments per row, 256 rows per image
    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bfull_bar, (uint32_t)S::B_BYTES);
            for (int pl = 0; pl < 2; ++pl)
                for (int c = 0; c < 4; ++c)
                    tma_load_3d(smem + (pl * 4 + c) * S::B_CHUNK, &tmB, bfull_bar, c * BKC, 0, pl);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int b = tile >> 10, yo = (tile >> 2) & 255, sg = tile & 3;
                mbar_wait(empty_bar + stage, phase ^ 1);
                uint8_t* sA = smem + S::A_OFF + stage * S::A_STAGE;
                mbar_expect_tx(full_bar + stage, 2u * S::A_PLANE);
                tma_load_4d(sA, &tmA, full_bar + stage, 0, sg * 16, 2 * yo, b);
                tma_load_4d(sA + S::A_PLANE, &tmA, full_bar + stage, 0, sg * 16, 2 * yo, a.B + b);
                if (++stage == S::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(64, 0, 0);
            mbar_wait(bfull_bar, 0);
            tc_fence_after();
            const uint32_t sB = smem_u32(smem);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty_bar + acc, ((it >> 1) & 1) ^ 1);
                mbar_wait(full_bar + stage, phase);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * 64);
                const uint32_t sA = smem_u32(smem + S::A_OFF + stage * S::A_STAGE);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {            // hi*hi, hi*lo, lo*hi
                    const uint32_t pa = sA + (prod == 2 ? S::A_PLANE : 0);
                    const uint32_t pb = sB + (prod == 1 ? 4 * S::B_CHUNK : 0);
#pragma unroll
                    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const uint64_t ad = umma_desc_interleave(pa + dy * S::A_ROW + h2 * 32, 16, 128);
                            const uint64_t bd = umma_desc_sw128(pb + (dy >> 1) * S::B_CHUNK) +
                                                (uint64_t)((((dy & 1) * 32 + h2 * 16) * 2) >> 4);
                            umma_f16(d, ad, bd, idesc, (prod | dy | h2) != 0);
                        }
                }
                umma_commit(empty_bar + stage);
                umma_commit(tfull_bar + acc);
                if (++stage == S::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= EPI_WARP0) {
        const int q = warp & 3;
        const int half = (warp - EPI_WARP0) >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int r = q * 32 + lane;
        float sc[32], sf[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            sc[j] = __ldg(a.scale + half * 32 + j) * (ACT_SCALE / STEM_IN_SCALE);   // tc_aux assumes 2^-4-scaled inputs
            sf[j] = __ldg(a.shift + half * 32 + j);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, eb = it & 1;
            mbar_wait(tfull_bar + acc, (it >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(acc * 64 + half * 32), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + acc);
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = fmaxf(fmaf(__uint_as_float(v[j]), sc[j], sf[j]), a.floor);
            mbar_wait(efree_bar + eb, ((it >> 1) & 1) ^ 1);
            const uint32_t e = smem_u32(smem + S::EPI_OFF + eb * S::EBUF + half * S::E_HALF) + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st_shared_v4(e + (uint32_t)((c ^ (r & 7)) << 4),
                             make_uint4(__float_as_uint(y[4 * c]), __float_as_uint(y[4 * c + 1]),
                                        __float_as_uint(y[4 * c + 2]), __float_as_uint(y[4 * c + 3])));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(oready_bar + eb);
        }
    } else if (warp == 3) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int b = tile >> 10, yo = (tile >> 2) & 255, sg = tile & 3;
                const int eb = it & 1;
                mbar_wait(oready_bar + eb, (it >> 1) & 1);
                const uint8_t* ebuf = smem + S::EPI_OFF + eb * S::EBUF;
                const int pix = (b * 256 + yo) * 514 + sg * 128 + 1;
                tma_store_2d(&tmO, ebuf, 0, pix);
                tma_store_2d(&tmO, ebuf + S::E_HALF, 32, pix);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(efree_bar + eb);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
    }
}

// ================================================================================================
// bott_tc_kernel: conv2 (3x3, stride 1, 64 -> 64) + BN + ReLU  FUSED WITH  conv3 (1x1, 64 -> 256) + BN + identity + ReLU
// of a layer1 bottleneck (torchvision Bottleneck.forward; reference model.py:78).
//
// Unfused, conv2 is bound by shared-memory operand reads (N = 64 tiles: 30 % tensor pipe, 16 % DRAM) and conv3 by HBM
// (its 256-channel residual read + output write: 72 % DRAM, 12 % tensor pipe), and they run one after the other.  Here one
// CTA computes, per tile of 128 pixels of one output row:
//   stage 1   the conv2 tile exactly as conv_tc_kernel<64> in "dxr" mode does (same MMA order, segments, epilogue maths),
//             but its epilogue writes the BN+ReLU'd result as hi/lo planes into SHARED memory (t2, [128 px][64 ch] rows of
//             128 B, 128-byte swizzle: exactly a K-major UMMA A operand) instead of HBM;
//   stage 2   conv3 as four 128 x 64 GEMM tiles with A = t2, B = conv3 weights, the residual tile TMA-prefetched into an
//             epilogue buffer, updated in place and TMA-stored (as gemm_tc_kernel does), halo columns by direct stores.
// The 64-channel intermediate never visits HBM (-0.54 GB per block) and, more importantly, conv2's MMAs of tile i+1 overlap
// conv3's HBM traffic of tile i.  The MMA warp issues  c2(0), c2(1), c3(0), c2(2), c3(1), ...  so that the tensor pipe
// never waits for the conv2 epilogue.  Results are bit-identical to the unfused kernels (same products, same
// accumulator structure, same epilogue arithmetic); HN_TC_FUSE=0 selects the unfused path.
//   warp 0      TMA producer (input rows, conv2 / conv3 weight tiles through one ring)
//   warp 1      MMA issuer          warp 2   TMEM allocator (512 columns), then residual-tile loader          warp 3   TMA store
//   warps 4-11  epilogue, conv2 and conv3 (lane quarter = warp % 4, column half = (warp - 4) / 4)
struct BottArgs {
    int Ho, Wo, Wop, wsegs, Bimg;
    int num_tiles;
    int n3;                              // conv3 n-tiles of 64 channels (Cout3 / 64)
    int C3;                              // conv3 output channels
    int seg;
    const float* scale2; const float* shift2;     // conv2: accumulator -> plane units, shift in plane units
    const float* scale3; const float* shift3;     // conv3
    unsigned short* out; size_t out_plane;        // output planes (halo-column stores)
};

struct BtSmem {
    static constexpr int AX_PLANE = 17 * 1024;            // 130 input pixels x 128 B (136 rows reserved)
    static constexpr int AX_SLOT = 2 * AX_PLANE;          // hi + lo
    static constexpr int NB = 3;                          // weight-tile ring
    static constexpr int B_PLANE = 64 * BKC * 2;          // 8 KB
    static constexpr int B_STAGE = 2 * B_PLANE;           // 16 KB
    static constexpr int B_OFF = 2 * AX_SLOT;             // 68 KB
    static constexpr int T2_OFF = B_OFF + NB * B_STAGE;   // 116 KB (1024-aligned)
    static constexpr int T_PLANE = BM * BKC * 2;          // 16 KB
    static constexpr int E_OFF = T2_OFF + 2 * T_PLANE;    // 148 KB
    static constexpr int E_PLANE = BM * 64 * 2;           // 16 KB
    static constexpr int EBUF = 2 * E_PLANE;
    static constexpr int BAR_OFF = E_OFF + 2 * EBUF;      // 212 KB
    static constexpr int CST_OFF = BAR_OFF + 512;         // epilogue constants: scale2[64] shift2[64] scale3[C3] shift3[C3] (C3 <= 512)
    static constexpr int TOTAL = CST_OFF + (128 + 2 * 512) * 4 + 1024;
    static constexpr int TMEM_COLS = 512;                 // c2: 2 main + 2 cross, c3: 2 main + 2 cross accumulators of 64 columns
};
static_assert(BtSmem::T2_OFF % 1024 == 0 && BtSmem::E_OFF % 1024 == 0 && BtSmem::B_OFF % 1024 == 0, "swizzle atoms need 1024-byte alignment");

__global__ void __launch_bounds__(NTHREADS, 1)
bott_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmO, const BottArgs a) {
    using S = BtSmem;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // [NB] weight tile landed
    uint64_t* empty_bar = full_bar + S::NB;         // [NB] weight tile consumed
    uint64_t* afull_bar = empty_bar + S::NB;        // [2] input-row slot filled
    uint64_t* aempty_bar = afull_bar + 2;           // [2] input-row slot consumed
    uint64_t* tfull_bar = aempty_bar + 2;           // [2] conv2 hi*hi segment ready
    uint64_t* tempty_bar = tfull_bar + 2;           // [2] conv2 segment drained
    uint64_t* cempty_bar = tempty_bar + 2;          // [2] conv2 cross accumulator drained
    uint64_t* t2ready_bar = cempty_bar + 2;         // [1] conv2 epilogue has written t2
    uint64_t* t2free_bar = t2ready_bar + 1;         // [1] (unused: the epilogue order itself guarantees that t2 is free)
    uint64_t* dfull_bar = t2free_bar + 1;           // [2] conv3 accumulators ready
    uint64_t* dempty_bar = dfull_bar + 2;           // [2] conv3 accumulators drained
    uint64_t* rfull_bar = dempty_bar + 2;           // [2] residual tile landed in the epilogue buffer
    uint64_t* oready_bar = rfull_bar + 2;           // [2] output tile complete in the epilogue buffer
    uint64_t* efree_bar = oready_bar + 2;           // [2] TMA store has read the epilogue buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(efree_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB3)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S::NB; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(afull_bar + i, 1); mbar_init(aempty_bar + i, 1);
            mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 8); mbar_init(cempty_bar + i, 8);
            mbar_init(dfull_bar + i, 1); mbar_init(dempty_bar + i, 8);
            mbar_init(rfull_bar + i, 1); mbar_init(oready_bar + i, 8); mbar_init(efree_bar + i, 1);
        }
        mbar_init(t2ready_bar, 8); mbar_init(t2free_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // epilogue constants -> shared memory (ncu: the per-tile __ldg of scale / shift cost the conv3 epilogue 20 % of its time --
    // they miss L1 after every fence.proxy.async; weights-side data, safe to read before pdl_wait)
    float* cst = reinterpret_cast<float*>(smem + S::CST_OFF);
    for (int i = threadIdx.x; i < 128 + 2 * a.C3; i += NTHREADS)
        cst[i] = (i < 64) ? a.scale2[i] : (i < 128) ? a.shift2[i - 64] : (i < 128 + a.C3) ? a.scale3[i - 128] : a.shift3[i - 128 - a.C3];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    // TMEM columns: conv2 main [0,64) [64,128), conv2 cross [128,192) [192,256); conv3 main [256,320) [320,384), cross [384,448) [448,512)
    const int n_local = (a.num_tiles > (int)blockIdx.x) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n3 = a.n3;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int ast = 0, bst = 0, g3 = 0;
            uint32_t aph = 0, bph = 0;
            auto load_c3 = [&](int) {                 // conv3 weight tiles of one pixel tile (same for every tile: L2 hits)
                for (int j = 0; j < n3; ++j, ++g3) {
                    mbar_wait(empty_bar + bst, bph ^ 1);
                    uint8_t* sB = smem + S::B_OFF + bst * S::B_STAGE;
                    mbar_expect_tx(full_bar + bst, (uint32_t)S::B_STAGE);
                    tma_load_3d(sB, &tmB3, full_bar + bst, 0, j * 64, 0);
                    tma_load_3d(sB + S::B_PLANE, &tmB3, full_bar + bst, 0, j * 64, 1);
                    if (++bst == S::NB) { bst = 0; bph ^= 1; }
                }
            };
            int prev = -1;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int rg = tile / a.wsegs;
                const int wo0 = (tile - rg * a.wsegs) * BM;
                const int b = rg / a.Ho, ho = rg - b * a.Ho;
                for (int dy = 0; dy < 3; ++dy) {
                    const int hin = ho + dy - 1;
                    mbar_wait(aempty_bar + ast, aph ^ 1);
                    uint8_t* sAx = smem + ast * S::AX_SLOT;
                    mbar_expect_tx(afull_bar + ast, 2u * 130u * 128u);
                    tma_load_4d(sAx, &tmA, afull_bar + ast, 0, wo0, hin, b);
                    tma_load_4d(sAx + S::AX_PLANE, &tmA, afull_bar + ast, 0, wo0, hin, a.Bimg + b);
                    if (++ast == 2) { ast = 0; aph ^= 1; }
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(empty_bar + bst, bph ^ 1);
                        uint8_t* sB = smem + S::B_OFF + bst * S::B_STAGE;
                        mbar_expect_tx(full_bar + bst, (uint32_t)S::B_STAGE);
                        const int kb = (dy * 3 + dx) * BKC;
                        tma_load_3d(sB, &tmB2, full_bar + bst, kb, 0, 0);
                        tma_load_3d(sB + S::B_PLANE, &tmB2, full_bar + bst, kb, 0, 1);
                        if (++bst == S::NB) { bst = 0; bph ^= 1; }
                    }
                }
                if (prev >= 0) load_c3(prev);
                prev = tile;
            }
            if (prev >= 0) load_c3(prev);
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(64, 0, 0, BM);
            int ast = 0, bst = 0, g = 0, g3 = 0;
            uint32_t aph = 0, bph = 0;
            const uint32_t t2 = smem_u32(smem + S::T2_OFF);
            auto conv3 = [&](int k) {                // conv3 of local tile k: A = t2 (hi, lo), B = weight ring
                mbar_wait(t2ready_bar, (uint32_t)(k & 1));
                tc_fence_after();
                const uint64_t a_hi = umma_desc_sw128(t2), a_lo = umma_desc_sw128(t2 + S::T_PLANE);
                for (int j = 0; j < n3; ++j, ++g3) {
                    const int db = g3 & 1;
                    mbar_wait(dempty_bar + db, ((g3 >> 1) & 1) ^ 1);
                    mbar_wait(full_bar + bst, bph);
                    tc_fence_after();
                    const uint32_t sB = smem_u32(smem + S::B_OFF + bst * S::B_STAGE);
                    const uint64_t b_hi = umma_desc_sw128(sB), b_lo = umma_desc_sw128(sB + S::B_PLANE);
                    const uint32_t d_main = tmem_base + 256 + db * 64, d_cross = tmem_base + 384 + db * 64;
#pragma unroll
                    for (int kk = 0; kk < BKC / 16; ++kk) {
                        const uint64_t ko = (uint64_t)((kk * 16 * 2) >> 4);
                        umma_f16(d_main, a_hi + ko, b_hi + ko, idesc, kk != 0);
                        umma_f16(d_cross, a_hi + ko, b_lo + ko, idesc, kk != 0);
                        umma_f16(d_cross, a_lo + ko, b_hi + ko, idesc, 1);
                    }
                    umma_commit(empty_bar + bst);
                    umma_commit(dfull_bar + db);
                    if (++bst == S::NB) { bst = 0; bph ^= 1; }
                }
            };
            for (int it = 0; it < n_local; ++it) {
                // ---- conv2 of local tile `it` (identical to conv_tc_kernel<64> in dxr mode)
                const int cbuf = it & 1;
                mbar_wait(cempty_bar + cbuf, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_cross = tmem_base + 128 + cbuf * 64;
                uint32_t d_main = tmem_base;
                int seg_pos = 0, mbuf = 0, kc = 0;
                for (int dy = 0; dy < 3; ++dy) {
                    mbar_wait(afull_bar + ast, aph);
                    tc_fence_after();
                    const uint32_t sAx = smem_u32(smem + ast * S::AX_SLOT);
                    for (int dx = 0; dx < 3; ++dx, ++kc) {
                        if (seg_pos == 0) {
                            mbuf = g & 1;
                            mbar_wait(tempty_bar + mbuf, ((g >> 1) & 1) ^ 1);
                            tc_fence_after();
                            d_main = tmem_base + mbuf * 64;
                        }
                        mbar_wait(full_bar + bst, bph);
                        tc_fence_after();
                        const uint32_t sB = smem_u32(smem + S::B_OFF + bst * S::B_STAGE);
                        const uint64_t a_hi = umma_desc_sw128(sAx + dx * 128);
                        const uint64_t a_lo = umma_desc_sw128(sAx + S::AX_PLANE + dx * 128);
                        const uint64_t b_hi = umma_desc_sw128(sB), b_lo = umma_desc_sw128(sB + S::B_PLANE);
#pragma unroll
                        for (int kk = 0; kk < BKC / 16; ++kk) {
                            const uint64_t ko = (uint64_t)((kk * 16 * 2) >> 4);
                            umma_f16(d_main, a_hi + ko, b_hi + ko, idesc, (seg_pos | kk) != 0);
                            umma_f16(d_cross, a_hi + ko, b_lo + ko, idesc, (kc | kk) != 0);
                            umma_f16(d_cross, a_lo + ko, b_hi + ko, idesc, 1);
                        }
                        umma_commit(empty_bar + bst);
                        if (++seg_pos == a.seg || kc == 8) {
                            umma_commit(tfull_bar + mbuf);
                            seg_pos = 0;
                            ++g;
                        }
                        if (++bst == S::NB) { bst = 0; bph ^= 1; }
                    }
                    umma_commit(aempty_bar + ast);
                    if (++ast == 2) { ast = 0; aph ^= 1; }
                }
                if (it > 0) conv3(it - 1);
            }
            if (n_local > 0) conv3(n_local - 1);
        }
    } else if (warp == 2) {
        // =============================== residual loader ===============================
        // its own thread, so that the identity tile of n-tile j is requested the moment its epilogue buffer is released
        // (two n-tiles ahead of its use): the residual comes from HBM, and behind the weight ring (first version) it was
        // requested only when the conv3 MMAs of that n-tile were about to issue -- the epilogue then waited a DRAM latency
        // per n-tile and the fused kernel was no faster than conv2 + conv3.
        if (lane == 0) {
            int g3 = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int rg = tile / a.wsegs;
                const int pix0 = rg * a.Wop + 1 + (tile - rg * a.wsegs) * BM;
                for (int j = 0; j < n3; ++j, ++g3) {
                    const int eb = g3 & 1;
                    mbar_wait(efree_bar + eb, ((g3 >> 1) & 1) ^ 1);
                    uint8_t* ebuf = smem + S::E_OFF + eb * S::EBUF;
                    mbar_expect_tx(rfull_bar + eb, 2u * S::E_PLANE);
                    tma_load_3d(ebuf, &tmR, rfull_bar + eb, j * 64, pix0, 0);
                    tma_load_3d(ebuf + S::E_PLANE, &tmR, rfull_bar + eb, j * 64, pix0, 1);
                }
            }
        }
    } else if (warp == 3) {
        // =============================== store warp ===============================
        if (lane == 0) {
            int g3 = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int rg = tile / a.wsegs;
                const int pix0 = rg * a.Wop + 1 + (tile - rg * a.wsegs) * BM;
                for (int j = 0; j < n3; ++j, ++g3) {
                    const int eb = g3 & 1;
                    mbar_wait(oready_bar + eb, (g3 >> 1) & 1);
                    const uint8_t* ebuf = smem + S::E_OFF + eb * S::EBUF;
                    tma_store_3d(&tmO, ebuf, j * 64, pix0, 0);
                    tma_store_3d(&tmO, ebuf + S::E_PLANE, j * 64, pix0, 1);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(efree_bar + eb);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp >= EPI_WARP0) {
        // =============================== epilogue warps (conv2 and conv3) ===============================
        // All 8 warps do both jobs (lane quarter = warp % 4, 32 of the 64 columns each).  Per tile `it`:
        //   (1) drain the conv2 segments + cross accumulator of tile it into 32 registers (BN + ReLU applied),
        //   (2) run the four conv3 n-tile epilogues of tile it-1 (their MMAs were issued after conv2(it)),
        //   (3) write the conv2 result as the t2 rows -- safe now: every conv3 MMA that reads t2(it-1) has completed,
        //       this warp has just consumed their accumulators -- and release conv3(it).
        // (First version: 4 warps for conv2 and 4 for conv3, 64 columns each; ncu showed the conv3 warps busy 70 % of
        // the time and everything else waiting for them: 19.5 k cycles per tile.)
        const int q = warp & 3;
        const int half = (warp - EPI_WARP0) >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int r = q * 32 + lane;                     // pixel of the tile = accumulator lane = row of t2 / epilogue buffer
        const int nseg = (9 + a.seg - 1) / a.seg;
        const uint32_t t_hi = smem_u32(smem + S::T2_OFF) + r * 128, t_lo = t_hi + S::T_PLANE;
        int g = 0, g3 = 0;
        auto conv3_epilogue = [&](int tile) {
            const int rg = tile / a.wsegs;
            const int wo0 = (tile - rg * a.wsegs) * BM;
            // circular halo columns of the output row (hn_common.cuh): wp = Wo + 1 copies wo = 0, wp = 0 copies wo = Wo - 1
            long long halo_pix = -1;
            if (wo0 + r == 0) halo_pix = (long long)rg * a.Wop + a.Wo + 1;
            else if (wo0 + r == a.Wo - 1) halo_pix = (long long)rg * a.Wop;
            for (int j = 0; j < n3; ++j, ++g3) {
                const int db = g3 & 1, eb = g3 & 1;
                mbar_wait(dfull_bar + db, (g3 >> 1) & 1);
                tc_fence_after();
                float sum[32];
                {
                    uint32_t v[32], w[32];
                    tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(256 + db * 64 + half * 32), v);
                    tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(384 + db * 64 + half * 32), w);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) sum[jj] = (0.f + __uint_as_float(v[jj])) + __uint_as_float(w[jj]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dempty_bar + db);
                const int n0 = j * 64 + half * 32;
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    const float4 sc = *reinterpret_cast<const float4*>(cst + 128 + n0 + jj);
                    const float4 sf = *reinterpret_cast<const float4*>(cst + 128 + a.C3 + n0 + jj);
                    sum[jj + 0] = fmaf(sum[jj + 0], sc.x, sf.x);
                    sum[jj + 1] = fmaf(sum[jj + 1], sc.y, sf.y);
                    sum[jj + 2] = fmaf(sum[jj + 2], sc.z, sf.z);
                    sum[jj + 3] = fmaf(sum[jj + 3], sc.w, sf.w);
                }
                mbar_wait(rfull_bar + eb, (g3 >> 1) & 1);
                const uint32_t e_hi = smem_u32(smem + S::E_OFF + eb * S::EBUF) + r * 128;
                const uint32_t e_lo = e_hi + S::E_PLANE;
                unsigned short* hrow = (halo_pix >= 0) ? a.out + (size_t)halo_pix * a.C3 + n0 : nullptr;
                uint4 rh[4], rl[4];                      // all eight 16-byte residual loads in flight before the first use
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                    rh[c] = ld_shared_v4(e_hi + off);
                    rl[c] = ld_shared_v4(e_lo + off);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                    float* y = sum + c * 8;
                    const uint32_t hw[4] = {rh[c].x, rh[c].y, rh[c].z, rh[c].w}, lw[4] = {rl[c].x, rl[c].y, rl[c].z, rl[c].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 fh = unpack_half2(hw[e]), fl = unpack_half2(lw[e]);
                        y[2 * e + 0] += fh.x + fl.x;
                        y[2 * e + 1] += fh.y + fl.y;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
                    uint32_t ph[4], pl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split2_scaled(y[2 * e], y[2 * e + 1], ph[e], pl[e]);
                    const uint4 oh = make_uint4(ph[0], ph[1], ph[2], ph[3]), ol = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    st_shared_v4(e_hi + off, oh);
                    st_shared_v4(e_lo + off, ol);
                    if (hrow) {
                        *reinterpret_cast<uint4*>(hrow + c * 8) = oh;
                        *reinterpret_cast<uint4*>(hrow + a.out_plane + c * 8) = ol;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(oready_bar + eb);
            }
        };
        int prev = -1, it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int cbuf = it & 1;
            float y2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y2[j] = 0.f;
            for (int sgi = 0; sgi < nseg; ++sgi, ++g) {
                const int mbuf = g & 1;
                mbar_wait(tfull_bar + mbuf, (g >> 1) & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(mbuf * 64 + half * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) y2[j] += __uint_as_float(v[j]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + mbuf);
            }
            {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(128 + cbuf * 64 + half * 32), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(cempty_bar + cbuf);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 sc = *reinterpret_cast<const float4*>(cst + half * 32 + j);
                    const float4 sf = *reinterpret_cast<const float4*>(cst + 64 + half * 32 + j);
                    y2[j + 0] = fmaxf(fmaf(y2[j + 0] + __uint_as_float(v[j + 0]), sc.x, sf.x), 0.f);
                    y2[j + 1] = fmaxf(fmaf(y2[j + 1] + __uint_as_float(v[j + 1]), sc.y, sf.y), 0.f);
                    y2[j + 2] = fmaxf(fmaf(y2[j + 2] + __uint_as_float(v[j + 2]), sc.z, sf.z), 0.f);
                    y2[j + 3] = fmaxf(fmaf(y2[j + 3] + __uint_as_float(v[j + 3]), sc.w, sf.w), 0.f);
                }
            }
            if (prev >= 0) conv3_epilogue(prev);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2_scaled(y2[c * 8 + 2 * e], y2[c * 8 + 2 * e + 1], ph[e], pl[e]);
                const uint32_t off = (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                st_shared_v4(t_hi + off, make_uint4(ph[0], ph[1], ph[2], ph[3]));
                st_shared_v4(t_lo + off, make_uint4(pl[0], pl[1], pl[2], pl[3]));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(t2ready_bar);
            prev = tile;
        }
        if (prev >= 0) conv3_epilogue(prev);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
    }
}

// -------------------------------------------------------------------------------- host side
int tc_segment_chunks() {
    // K chunks (of 64) accumulated in TMEM before the sum is promoted to registers; HN_TC_SEG overrides (tuning)
    static int seg = [] {
        const char* e = getenv("HN_TC_SEG");
        int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : v;
    }();
    return seg;
}

bool tc_prefetch_on() {
    // gemm_tc_kernel prefetches the activation / residual boxes of its K = 64 layers into L2 two tile rounds ahead; HN_TC_PF=0 disables
    static const bool on = [] { const char* e = getenv("HN_TC_PF"); return !(e && atoi(e) == 0); }();
    return on;
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, cudaStream_t st) {
    using S = Smem<BN>;
    HN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    int dev = 0, sms = 0;
    HN_CUDA_OK(cudaGetDevice(&dev));
    HN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    HN_CUDA_OK(launch_tc(conv_tc_kernel<BN>, grid, S::TOTAL, st, tmA, tmB, a));
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace

bool conv_tc_supported(const ConvDesc& d, const Act& in, const Act& out) {
    if (d.Cin % 64 != 0 || d.Cout % 32 != 0) return false;
    if (d.Cout > 64 && d.Cout % 128 != 0) return false;
    if (d.Cout < 64 && d.Cout != 32) return false;
    if (d.pw > in.halo) return false;
    const bool gemm = (d.kh == 1 && d.kw == 1 && d.sh == 1 && d.sw == 1 && in.halo == out.halo);
    if (gemm) return true;
    if (d.sw != 1 && d.sw != 2) return false;
    if (d.sw == 2 && (in.Wp() % 2) != 0) return false;
    const int tw = out.W < 128 ? out.W : 128;
    if (tw < 8 || (tw & (tw - 1)) != 0 || out.W % tw != 0) return false;       // 8-row swizzle atoms, exact W tiling
    return true;
}

// in / out / residual are split plane pairs in halo-NHWC geometry (hi plane, then lo plane).
// wq: [2][Cout][K] weight planes, K = (dy*kw+dx)*Cin + c; tc_scale folds BN scale and the plane scales.
int conv_tc_planes(const ConvDesc& d, const unsigned short* wq, const float* tc_aux, const Act& in,
                   const unsigned short* in_planes, const Act& out, unsigned short* out_planes, float* out_f32,
                   const unsigned short* res_planes, cudaStream_t st) {
    HN_CHECK(conv_tc_supported(d, in, out), "conv_tc: unsupported shape");
    const int K = d.kh * d.kw * d.Cin;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    const bool gemm = (d.kh == 1 && d.kw == 1 && d.sh == 1 && d.sw == 1 && in.halo == out.halo);
    int BN = d.Cout >= 128 ? 128 : d.Cout;
    {   // layers with too few 128x128 tiles to fill the SMs (the tail of the height-reduction convs): halve the tile width
        const long long rows = gemm ? (long long)in.B * in.H * in.Wp() : (long long)out.B * out.H * out.W;
        const long long tiles128 = ((rows + BM - 1) / BM) * (d.Cout / 128 > 0 ? d.Cout / 128 : 1);
        if (BN == 128 && tiles128 < 110) BN = 64;
    }
    const size_t in_plane = in.numel(), out_plane = out.numel();
    CUtensorMap tmA, tmB;
    {   // weights: {K, Cout, 2}
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)d.Cout, 2};
        cuuint64_t str[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * d.Cout * 2};
        cuuint32_t box[3] = {BKC, (cuuint32_t)BN, 1};
        if (make_map(&tmB, wq, 3, dims, str, box)) return -1;
    }
    a.Cout = d.Cout;
    a.n_tiles = d.Cout / BN;
    a.kw = d.kw;
    a.kc_per_tap = d.Cin / BKC;
    a.num_kc = d.kh * d.kw * a.kc_per_tap;
    HN_CHECK(!(out_f32 && res_planes), "conv_tc: residual with fp32 output is not supported");
    a.scale = out_f32 ? tc_aux : tc_aux + d.Cout;            // fp32 outputs in true units, planes in plane units
    a.shift = out_f32 ? d.shift : tc_aux + 2 * d.Cout;
    a.res = res_planes; a.out = out_planes; a.out_f32 = out_f32; a.out_plane = out_plane; a.relu = d.relu;
    a.Bimg = in.B;
    a.seg = tc_segment_chunks();
    { static int dbg = [] { const char* e = getenv("HN_TC_DBG"); return e ? atoi(e) : 0; }(); a.dbg = dbg; }
    long long m_tiles;
    // asynchronous-epilogue GEMM kernel: plane output, K <= 512, Cout a multiple of 64
    static const bool gemm_kernel_on = [] { const char* e = getenv("HN_TC_GEMM"); return !(e && atoi(e) == 0); }();
    const bool use_gemm_kernel = gemm && gemm_kernel_on && !out_f32 && d.Cin <= 256 && d.Cout % GBN == 0 && d.Cout <= GSmem::MAX_COUT;
    if (use_gemm_kernel) {
        const long long Mtot = (long long)in.B * in.H * in.Wp();
        a.mode = 0;
        a.M = (int)Mtot;
        a.n_tiles = d.Cout / GBN;
        // L2 prefetch two tile rounds ahead, K = 64 only: measured -9..-11 % on layer1's conv1 / downsample (HBM-bound, one
        // 32 KB activation box per tile); +38 % on the K = 256 conv1 (8 boxes per tile compete with the demand loads) and
        // +2..4 % on conv3 of layer2/3 and on bott_tc_kernel, where it is therefore not used (profiles/r02_experiments.md)
        a.pf = (tc_prefetch_on() && a.num_kc == 1) ? 2 : 0;
        CUtensorMap tmR, tmO;
        {
            cuuint64_t dims[3] = {(cuuint64_t)d.Cin, (cuuint64_t)Mtot, 2};
            cuuint64_t str[2] = {(cuuint64_t)d.Cin * 2, (cuuint64_t)in_plane * 2};
            cuuint32_t box[3] = {BKC, BM, 1};
            if (make_map(&tmA, in_planes, 3, dims, str, box)) return -1;
        }
        {
            cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)d.Cout, 2};
            cuuint64_t str[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * d.Cout * 2};
            cuuint32_t box[3] = {BKC, GBN, 1};
            if (make_map(&tmB, wq, 3, dims, str, box)) return -1;
        }
        {
            cuuint64_t dims[3] = {(cuuint64_t)d.Cout, (cuuint64_t)Mtot, 2};
            cuuint64_t str[2] = {(cuuint64_t)d.Cout * 2, (cuuint64_t)out_plane * 2};
            cuuint32_t box[3] = {GBN, BM, 1};
            if (make_map(&tmO, out_planes, 3, dims, str, box)) return -1;
            if (make_map(&tmR, res_planes ? res_planes : out_planes, 3, dims, str, box)) return -1;
        }
        const long long mt = (Mtot + BM - 1) / BM;
        HN_CHECK(mt * a.n_tiles < (1ll << 31), "conv_tc: too many tiles");
        a.num_tiles = (int)(mt * a.n_tiles);
        int dev = 0, sms = 0;
        HN_CUDA_OK(cudaGetDevice(&dev));
        HN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        // (Round 2 negative result: a variant that keeps the 128 x K activation tile resident across the n-tiles of an
        // m-tile -- the tile is otherwise re-fetched from L2 per n-tile, 8-16x on conv3 of layer2/3 -- measured 10-17 %
        // SLOWER per launch (layer3 conv3 183 -> 214 us, layer2 conv3 224 -> 247 us): the resident tile leaves room for
        // only 2 weight stages at K = 256 and serialises the A fill at every unit start.  Removed; see profiles/README.md.)
        HN_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GSmem::TOTAL));
        HN_CUDA_OK(launch_tc(gemm_tc_kernel, a.num_tiles < sms ? a.num_tiles : sms, GSmem::TOTAL, st, tmA, tmB, tmR, tmO, a));
        HN_LAUNCH_OK();
        return 0;
    }
    if (gemm) {
        const long long Mtot = (long long)in.B * in.H * in.Wp();
        a.mode = 0;
        a.M = (int)Mtot;
        cuuint64_t dims[3] = {(cuuint64_t)d.Cin, (cuuint64_t)Mtot, 2};
        cuuint64_t str[2] = {(cuuint64_t)d.Cin * 2, (cuuint64_t)in_plane * 2};
        cuuint32_t box[3] = {BKC, BM, 1};
        if (make_map(&tmA, in_planes, 3, dims, str, box)) return -1;
        m_tiles = (Mtot + BM - 1) / BM;
    } else {
        a.mode = 1;
        a.Ho = out.H; a.Wo = out.W; a.Wop = out.Wp(); a.out_halo = out.halo;
        a.M = out.B * out.H;
        a.tw = out.W < 128 ? out.W : 128;
        a.rows_per_tile = BM / a.tw;
        a.wsegs = out.W / a.tw;
        a.sh = d.sh; a.ph = d.ph; a.woff = in.halo - d.pw;
        a.parity = (d.sw == 2);
        // one box for all rows of a tile when they are consecutive input rows of one image (stride 1 along H)
        static const bool rowbox_on = [] { const char* e = getenv("HN_TC_ROWBOX"); return !(e && atoi(e) == 0); }();
        static const bool rowbox2_on = [] { const char* e = getenv("HN_TC_ROWBOX2"); return !(e && atoi(e) == 0); }();
        // ... or whole images when an image has fewer output rows than a tile (the last height-reduction convs):
        // the box then spans R / Ho consecutive images of the plane
        int box_rows = 1, box_imgs = 1;
        a.rowbox = 0;
        if (rowbox_on && a.rows_per_tile > 1 && (d.sh == 1 || (d.sh == 2 && rowbox2_on))) {
            if (out.H % a.rows_per_tile == 0) { a.rowbox = 1; box_rows = a.rows_per_tile; }
            else if (a.rows_per_tile % out.H == 0) { a.rowbox = 1; box_rows = out.H; box_imgs = a.rows_per_tile / out.H; }
        }
        // stride 2 along H (height-reduction convs): the box spans 2*rows input rows and takes every second one
        const cuuint32_t boxrows = a.rowbox ? (cuuint32_t)(box_rows * d.sh) : 1u;
        const int hs = (a.rowbox && d.sh == 2) ? 2 : 1;
        const cuuint64_t C2 = (cuuint64_t)d.Cin * 2, Wp = in.Wp();
        // dx reuse (HN_TC_DXR=0 disables)
        static const int dxr_mode = [] { const char* e = getenv("HN_TC_DXR"); return e ? atoi(e) : 1; }();
        a.dxr = (dxr_mode != 0 && !a.parity && d.kw == 3 && a.rows_per_tile == 1 && a.tw == BM &&
                 (long long)a.wsegs * BM + a.woff + 2 <= (long long)Wp) ? 1 : 0;
        if (!a.parity) {
            cuuint64_t dims[4] = {(cuuint64_t)d.Cin, Wp, (cuuint64_t)in.H, (cuuint64_t)2 * in.B};
            cuuint64_t str[3] = {C2, C2 * Wp, C2 * Wp * in.H};
            cuuint32_t box[4] = {BKC, a.dxr ? 130u : (cuuint32_t)a.tw, boxrows, (cuuint32_t)box_imgs};
            if (make_map(&tmA, in_planes, 4, dims, str, box, hs == 2 ? 2 : -1, hs)) return -1;
        } else {
            cuuint64_t dims[5] = {(cuuint64_t)d.Cin, 2, Wp / 2, (cuuint64_t)in.H, (cuuint64_t)2 * in.B};
            cuuint64_t str[4] = {C2, 2 * C2, C2 * Wp, C2 * Wp * in.H};
            cuuint32_t box[5] = {BKC, 1, (cuuint32_t)a.tw, boxrows, (cuuint32_t)box_imgs};
            if (make_map(&tmA, in_planes, 5, dims, str, box, hs == 2 ? 3 : -1, hs)) return -1;
        }
        m_tiles = (long long)((a.M + a.rows_per_tile - 1) / a.rows_per_tile) * a.wsegs;
    }
    HN_CHECK(m_tiles * a.n_tiles < (1ll << 31), "conv_tc: too many tiles");
    a.num_tiles = (int)(m_tiles * a.n_tiles);
    if (a.num_tiles == 0) return 0;
    // (CTA pairs -- cta_group::2, M = 256 per pair, B operand fetched once per pair -- were implemented, verified and measured
    // 5-25 % SLOWER than single-CTA tiles on every layer of this net in round 1 (each SM still serves its B half to both tensor
    // cores; only the L2 fill is halved); the variant was deleted in round 2)
    switch (BN) {
        case 128: return launch<128>(tmA, tmB, a, st);
        case 64: return launch<64>(tmA, tmB, a, st);
        default: return launch<32>(tmA, tmB, a, st);
    }
}


// ---------------------------------------------------------------------- fused conv2 + conv3 of a bottleneck (host side)
bool bott_tc_supported(const ConvDesc& d2, const ConvDesc& d3, const Act& in, const Act& out) {
    static const bool on = [] { const char* e = getenv("HN_TC_FUSE"); return !(e && atoi(e) == 0); }();
    if (!on) return false;
    return d2.kh == 3 && d2.kw == 3 && d2.sh == 1 && d2.sw == 1 && d2.ph == 1 && d2.pw == 1 && d2.Cin == 64 && d2.Cout == 64 &&
           d2.relu && d3.kh == 1 && d3.kw == 1 && d3.sh == 1 && d3.sw == 1 && d3.Cin == 64 && d3.Cout % 64 == 0 && d3.relu &&
           in.halo == 1 && out.halo == 1 && in.H == out.H && in.W == out.W && in.B == out.B && out.W % BM == 0 &&
           in.C == 64 && out.C == d3.Cout;
}

// in: conv2 input planes [B][H][W+2][64]; res / out: [B][H][W+2][C3] planes (res = the block's identity); wq2 / aux2, wq3 / aux3
// as produced by pack_weight_tc for the two convolutions.
int bott_tc_planes(const ConvDesc& d2, const unsigned short* wq2, const float* aux2, const ConvDesc& d3,
                   const unsigned short* wq3, const float* aux3, const Act& in, const unsigned short* in_planes, const Act& out,
                   unsigned short* out_planes, const unsigned short* res_planes, cudaStream_t st) {
    HN_CHECK(bott_tc_supported(d2, d3, in, out), "bott_tc: unsupported shape");
    HN_CHECK(res_planes != nullptr, "bott_tc: the fused kernel expects the block's identity");
    const size_t in_plane = in.numel(), out_plane = out.numel();
    const cuuint64_t Wp = in.Wp();
    CUtensorMap tmA, tmB2, tmB3, tmR, tmO;
    {
        cuuint64_t dims[4] = {64, Wp, (cuuint64_t)in.H, (cuuint64_t)2 * in.B};
        cuuint64_t str[3] = {128, 128 * Wp, 128 * Wp * in.H};
        cuuint32_t box[4] = {BKC, 130, 1, 1};
        if (make_map(&tmA, in_planes, 4, dims, str, box)) return -1;
        (void)in_plane;
    }
    {
        cuuint64_t dims[3] = {576, 64, 2};
        cuuint64_t str[2] = {576 * 2, 576 * 64 * 2};
        cuuint32_t box[3] = {BKC, 64, 1};
        if (make_map(&tmB2, wq2, 3, dims, str, box)) return -1;
    }
    {
        cuuint64_t dims[3] = {64, (cuuint64_t)d3.Cout, 2};
        cuuint64_t str[2] = {64 * 2, (cuuint64_t)64 * d3.Cout * 2};
        cuuint32_t box[3] = {BKC, 64, 1};
        if (make_map(&tmB3, wq3, 3, dims, str, box)) return -1;
    }
    {
        const cuuint64_t Mtot = (cuuint64_t)out.B * out.H * out.Wp();
        cuuint64_t dims[3] = {(cuuint64_t)d3.Cout, Mtot, 2};
        cuuint64_t str[2] = {(cuuint64_t)d3.Cout * 2, (cuuint64_t)out_plane * 2};
        cuuint32_t box[3] = {64, BM, 1};
        if (make_map(&tmO, out_planes, 3, dims, str, box)) return -1;
        if (make_map(&tmR, res_planes, 3, dims, str, box)) return -1;
    }
    BottArgs a;
    memset(&a, 0, sizeof(a));
    a.Ho = out.H; a.Wo = out.W; a.Wop = out.Wp(); a.wsegs = out.W / BM; a.Bimg = in.B;
    const long long tiles = (long long)out.B * out.H * a.wsegs;
    HN_CHECK(tiles < (1ll << 31) && (long long)out.B * out.H * out.Wp() < (1ll << 31), "bott_tc: too many tiles");
    a.num_tiles = (int)tiles;
    a.n3 = d3.Cout / 64; a.C3 = d3.Cout;
    a.seg = tc_segment_chunks();
    a.scale2 = aux2 + 64; a.shift2 = aux2 + 2 * 64;                        // plane units (conv_tc.cuh: tc_aux)
    a.scale3 = aux3 + d3.Cout; a.shift3 = aux3 + 2 * d3.Cout;
    a.out = out_planes; a.out_plane = out_plane;
    if (a.num_tiles == 0) return 0;
    HN_CUDA_OK(cudaFuncSetAttribute(bott_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BtSmem::TOTAL));
    int dev = 0, sms = 0;
    HN_CUDA_OK(cudaGetDevice(&dev));
    HN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HN_CUDA_OK(launch_tc(bott_tc_kernel, a.num_tiles < sms ? a.num_tiles : sms, BtSmem::TOTAL, st, tmA, tmB2, tmB3, tmR, tmO, a));
    HN_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------- stem on tensor cores (host side)
size_t stem_tc_scratch_bytes(int B) { return (size_t)2 * B * SP_ROWS * SP_PAIRS * 16; }

int stem_tc_pack_weights(const float* w_oihw, float* pad_scratch, unsigned short* wq, const float* scale,
                         const float* shift, float* tc_aux, cudaStream_t st) {
    stem_weight_pad_kernel<<<(64 * SP_K + 255) / 256, 256, 0, st>>>(w_oihw, pad_scratch);
    HN_LAUNCH_OK();
    return pack_weight_tc(pad_scratch, wq, scale, shift, tc_aux, tc_aux + 3 * 64, 64, 4, 7, 8, st);
}

int stem_tc(const float* x_nchw, int B, int in_channels, const unsigned short* wq, const float* tc_aux,
            const float* shift, unsigned short* scratch, const Act& out, cudaStream_t st, bool relu) {
    HN_CHECK(in_channels >= 3, "stem: input needs >= 3 channels (reference model.py:252 uses x[:, :3])");
    HN_CHECK(out.B == B && out.H == 256 && out.W == 512 && out.C == 64 && out.halo == 1, "stem: bad output tensor");
    HN_CHECK(B >= 1 && B <= 2048, "stem: bad batch");
    const size_t total = (size_t)B * SP_ROWS * SP_PAIRS;
    stem_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_nchw, in_channels, scratch, B);
    HN_LAUNCH_OK();
    CUtensorMap tmA, tmB, tmO;
    {   // packed input: {64 = 8 pairs x 8 elements, 65 pair groups, 517 rows, 2 planes x B}
        cuuint64_t dims[4] = {64, SP_PAIRS / 8, SP_ROWS, (cuuint64_t)2 * B};
        cuuint64_t str[3] = {128, (cuuint64_t)SP_PAIRS * 16, (cuuint64_t)SP_ROWS * SP_PAIRS * 16};
        cuuint32_t box[4] = {64, 17, 7, 1};
        if (make_map(&tmA, scratch, 4, dims, str, box, -1, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE))
            return -1;
    }
    {   // weights: {K = 224, 64, 2}; the last 64-wide chunk is half out of bounds (zero filled, never multiplied)
        cuuint64_t dims[3] = {SP_K, 64, 2};
        cuuint64_t str[2] = {SP_K * 2, SP_K * 64 * 2};
        cuuint32_t box[3] = {BKC, 64, 1};
        if (make_map(&tmB, wq, 3, dims, str, box)) return -1;
    }
    {   // fp32 halo-NHWC output as {64 channels, B*256*514 pixels}
        cuuint64_t dims[2] = {64, (cuuint64_t)B * 256 * 514};
        cuuint64_t str[1] = {256};
        cuuint32_t box[2] = {32, 128};
        if (make_map(&tmO, out.p, 2, dims, str, box, -1, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return -1;
    }
    StemArgs a;
    a.B = B;
    a.num_tiles = B * 1024;
    a.scale = tc_aux;
    a.shift = shift;
    a.floor = relu ? 0.f : -INFINITY;       // no ReLU: the raw output the train-mode batch statistics are taken from
    HN_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, StSmem::TOTAL));
    int dev = 0, sms = 0;
    HN_CUDA_OK(cudaGetDevice(&dev));
    HN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HN_CUDA_OK(launch_tc(stem_tc_kernel, a.num_tiles < sms ? a.num_tiles : sms, StSmem::TOTAL, st, tmA, tmB, tmO, a));
    HN_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------- format conversion kernels
namespace {

__global__ void split_kernel(const float* __restrict__ in, unsigned short* __restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    unsigned short h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_scaled(f[j], h[j], l[j]);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    *reinterpret_cast<uint2*>(out + n + i) = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
}

__global__ void merge_kernel(const unsigned short* __restrict__ in, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = merge_scaled(in[i], in[n + i]);
}

__global__ void absmax_kernel(const float* __restrict__ w, size_t n, float* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0: int order = float order
}

// weight scale exponent t: max|w| * 2^t lands in (2^13, 2^14]: inside fp16 range, lo plane normal
__device__ __forceinline__ float weight_scale(float absmax) {
    if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
    int e;
    frexpf(16384.f / absmax, &e);
    return ldexpf(1.f, e - 1);
}

// fp32 weights -> [2][Cout][K] planes of w * 2^t (fp16 hi, fp16 lo); source index from (n, k)
template <bool OIHW>
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, unsigned short* __restrict__ out,
                                      const float* __restrict__ absmax, int Cout, int Cin, int kh, int kw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t K = (size_t)Cin * kh * kw;
    const size_t total = (size_t)Cout * K;
    if (i >= total) return;
    const int n = (int)(i / K);
    const size_t k = i - (size_t)n * K;
    float v;
    if (OIHW) {
        const int c = (int)(k % Cin);
        const int tap = (int)(k / Cin);
        const int dy = tap / kw, dx = tap % kw;
        v = w[(((size_t)n * Cin + c) * kh + dy) * kw + dx];
    } else {
        v = w[k * Cout + n];                 // packed [K][Cout] (unit-test entry point)
    }
    const float s = v * weight_scale(*absmax);
    const __half h = __float2half_rn(s);
    out[i] = __half_as_ushort(h);
    out[total + i] = __half_as_ushort(__float2half_rn(s - __half2float(h)));
}

__global__ void tc_scale_kernel(const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ absmax, float* __restrict__ tc_aux, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    const float s = (scale ? scale[i] : 1.f) * ACT_UNSCALE / weight_scale(*absmax);   // accumulator -> true units
    tc_aux[i] = s;
    tc_aux[C + i] = s * ACT_SCALE;                                                    // accumulator -> plane units
    tc_aux[2 * C + i] = (shift ? shift[i] : 0.f) * ACT_SCALE;
}

template <bool OIHW>
int pack_weight_impl(const float* w, unsigned short* wq, const float* scale, const float* shift, float* tc_aux,
                     float* scratch, int Cout, int Cin, int kh, int kw, cudaStream_t st) {
    const size_t total = (size_t)Cout * Cin * kh * kw;
    HN_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float), st));
    // 4 elements per thread (the first version's 64 serial loads per thread cost 32 us per layer: 4.4 ms per training step)
    absmax_kernel<<<(unsigned)((total + 256 * 4 - 1) / (256 * 4)), 256, 0, st>>>(w, total, scratch);
    HN_LAUNCH_OK();
    pack_weight_tc_kernel<OIHW><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, wq, scratch, Cout, Cin, kh, kw);
    HN_LAUNCH_OK();
    tc_scale_kernel<<<(Cout + 255) / 256, 256, 0, st>>>(scale, shift, scratch, tc_aux, Cout);
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace

int split_planes(const float* in, unsigned short* out, size_t n, cudaStream_t st) {
    HN_CHECK(n % 4 == 0, "split_planes: element count must be a multiple of 4");
    if (n == 0) return 0;
    split_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(in, out, n);
    HN_LAUNCH_OK();
    return 0;
}

int merge_planes(const unsigned short* in, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
    HN_LAUNCH_OK();
    return 0;
}

int pack_weight_tc(const float* w_oihw, unsigned short* wq, const float* scale, const float* shift, float* tc_aux,
                   float* scratch, int Cout, int Cin, int kh, int kw, cudaStream_t st) {
    return pack_weight_impl<true>(w_oihw, wq, scale, shift, tc_aux, scratch, Cout, Cin, kh, kw, st);
}

int tc_aux_update(const float* scale, const float* shift, const float* absmax, float* tc_aux, int Cout, cudaStream_t st) {
    tc_scale_kernel<<<(Cout + 255) / 256, 256, 0, st>>>(scale, shift, absmax, tc_aux, Cout);
    HN_LAUNCH_OK();
    return 0;
}

// Unit-test convenience: fp32 halo-NHWC in and out, planes built on the fly.
int conv_tc(const ConvDesc& d, const Act& in, const Act& out, const float* residual, cudaStream_t st) {
    HN_CHECK(conv_tc_supported(d, in, out), "conv_tc: unsupported shape");
    const size_t K = (size_t)d.kh * d.kw * d.Cin;
    const size_t n_in = in.numel(), n_out = out.numel(), n_w = K * d.Cout;
    unsigned short *pin = nullptr, *pout = nullptr, *pres = nullptr, *pw = nullptr;
    float* aux = nullptr;          // [3*Cout] tc_aux + 1 scratch float
    HN_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&pin), n_in * 4, st));
    HN_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&pout), n_out * 4, st));
    HN_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&pw), n_w * 4, st));
    HN_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&aux), (3 * d.Cout + 1) * sizeof(float), st));
    if (residual) HN_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&pres), n_out * 4, st));
    int rc = split_planes(in.p, pin, n_in, st);
    if (!rc && residual) rc = split_planes(residual, pres, n_out, st);
    if (!rc) rc = pack_weight_impl<false>(d.w, pw, d.scale, d.shift, aux, aux + 3 * d.Cout, d.Cout, d.Cin, d.kh, d.kw, st);
    if (!rc) rc = conv_tc_planes(d, pw, aux, in, pin, out, pout, nullptr, pres, st);
    if (!rc) rc = merge_planes(pout, out.p, n_out, st);
    cudaFreeAsync(pin, st); cudaFreeAsync(pout, st); cudaFreeAsync(pw, st); cudaFreeAsync(aux, st);
    if (pres) cudaFreeAsync(pres, st);
    return rc;
}

}  // namespace hn
