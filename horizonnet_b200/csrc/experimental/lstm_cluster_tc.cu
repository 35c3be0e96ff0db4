// ROUND-2 CANDIDATE -- compiles for sm_100a, NOT yet run on hardware, NOT part of the library build
// (horizonnet_b200/build.py SOURCES does not list it).  Validate with tests/test_gpu_parity.py::test_lstm_* after adding
// it to SOURCES and routing lstm_layer_cluster() to lstm_layer_cluster_tc().
//
// Why: lstm_cluster.cu is bound by the mma.sync issue rate (ncu: HMMA pipe 98 % busy, ~4.5 clk per m16n8k16 per SM,
// 768 of them per CTA and group-step = 1.98 us).  Here two of the three products move to tcgen05:
//     W_hi*h_hi + W_hi*h_lo   one tcgen05.mma chain per group-step: A = W_hi (128 rows x 512, fp16, shared memory,
//                             K-major 128-byte swizzle), B = the RECEIVE BUFFER ITSELF laid out as the no-swizzle K-major
//                             operand [N = 16 rows: 8 columns of h_hi, 8 columns of h_lo][K = 512]
//                             (core matrix (k-block kb, plane g) at kb*256 + g*128, LBO = 256 B, SBO = 128 B), so the
//                             1024-byte block a peer bulk-copies in IS four K-blocks of the operand; D[128 x 16] fp32
//                             in tensor memory, 32 MMAs (M128 N16 K16) per group-step, issued by one thread
//     W_lo*h_hi               stays on mma.sync (256 instead of 768 per CTA and group-step), W_lo fragments in registers:
//                             warp w owns rows 16w..16w+15 for the whole K, so no cross-warp reduction
// Everything else (cluster of 16 CTAs per (direction, column group pair), DSMEM exchange protocol, double buffering by
// step parity, cell math) is lstm_cluster.cu's.  Expected: ~1.0 us per group-step instead of 1.98 us.
#include <cuda_fp16.h>
#include <cstdlib>
#include "../hn_common.cuh"
#include "../ptx.cuh"

namespace hn {

namespace {

constexpr int HID = 512;
constexpr int CL = 16;
constexpr int UNITS = HID / CL;         // 32 hidden units per CTA
constexpr int NCOL = 8;                 // batch columns per group
constexpr int NT = 256;
constexpr int MAXG = 2;
constexpr float H_SCALE = 256.f;
constexpr int BLK = 1024;               // one CTA's h of one step: [4 k-blocks][2 planes][8 columns][8 units] fp16
constexpr int OPB = CL * BLK;           // 16 KB: the B operand [16][512] of one (group, parity)
constexpr int PCOL = 132;

constexpr int SM_WHI = 0;                                  // [8 chunks][128 rows][128 B], 128-byte swizzle
constexpr int SM_RECV = SM_WHI + 8 * 128 * 128;            // [MAXG][2 parity][OPB]
constexpr int SM_STAGE = SM_RECV + MAXG * 2 * OPB;         // [MAXG][2 parity][BLK]
constexpr int SM_LO = SM_STAGE + MAXG * 2 * BLK;           // [8 columns][PCOL] fp32: W_lo*h_hi of the current group
constexpr int SM_GATE = SM_LO + NCOL * PCOL * 4;           // [4 gates][8 columns][32 units] fp32: (W_hi*h_hi + W_hi*h_lo)
constexpr int SM_ROWS = SM_GATE + 4 * NCOL * UNITS * 4;    // [128] 1 / (row scale * H_SCALE)
constexpr int SM_BAR = SM_ROWS + 128 * 4;                  // full[MAXG][2], dfull[MAXG], TMEM slot
constexpr int SM_TOTAL = SM_BAR + 64 + 1024;               // + slack to align the base to 1024 B
constexpr int TMEM_COLS = 32;                              // D[group]: 16 columns each

struct ClArgs {
    const float* xproj;
    const float* w_hh[2];
    float* out;
    int T, B, ngroups, gpc;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo_elem, float hi_elem) {
    const __half2 h = __floats2half2_rn(lo_elem, hi_elem);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float row_scale(float absmax) {
    if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
    int e;
    frexpf(16384.f / absmax, &e);
    return ldexpf(1.f, e - 1);
}
// K-major 128-byte-swizzle operand (rows of 128 B, 8-row atoms 1024 B apart) -- as in conv_tc.cu
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major operand without swizzle: 8-row x 16-byte core matrices, lbo between the two K core matrices of a K=16 step,
// sbo between 8-row groups -- as stem_tc_kernel's A operand in conv_tc.cu
__device__ __forceinline__ uint64_t desc_interleave(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__global__ void __launch_bounds__(NT, 1) lstm_cluster_tc_kernel(const ClArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);   // [MAXG][2]
    uint64_t* dfull_bar = full_bar + MAXG * 2;                          // [MAXG] MMAs of the group-step retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull_bar + MAXG);
    float* lo_buf = reinterpret_cast<float*>(smem + SM_LO);
    float* gate_buf = reinterpret_cast<float*>(smem + SM_GATE);
    float* row_unscale = reinterpret_cast<float*>(smem + SM_ROWS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const uint32_t rank = cl_rank();
    const int cid = blockIdx.x / CL;
    const int dir = cid % 2;
    const int g_first = (cid / 2) * a.gpc;
    const int ng = min(a.gpc, a.ngroups - g_first);
    const int tstep = dir ? -1 : 1;
    const int t_first = dir ? a.T - 1 : 0;

    if (tid == 0) {
        for (int i = 0; i < MAXG * 2; ++i) mbar_init(full_bar + i, 1);
        for (int i = 0; i < MAXG; ++i) mbar_init(dfull_bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- W_hh slice: warp w owns local rows 16w .. 16w+15 (local row lr = gate*32 + unit) for all 512 k.
    // hi -> shared memory (tcgen05 A operand), lo -> mma.sync A fragments in registers.
    // fragment register (rs + 2hf) of k-tile kt: row 16w + 8rs + gid, k = 16kt + 2tig + 8hf (+1)
    uint32_t a_lo[32][4];
    {
        float sc[2];
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
            const int lr = 16 * warp + 8 * rs + gid;
            const float* wrow = a.w_hh[dir] + (size_t)((lr >> 5) * HID + rank * UNITS + (lr & 31)) * HID;
            float mx = 0.f;
            for (int kt = 0; kt < 32; ++kt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(wrow + 16 * kt + tig * 2 + 8 * hf));
                    mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            sc[rs] = row_scale(mx);
            if (tig == 0) row_unscale[lr] = 1.f / (sc[rs] * H_SCALE);
        }
#pragma unroll
        for (int kt = 0; kt < 32; ++kt)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
                const int lr = 16 * warp + 8 * rs + gid;
                const float* wrow = a.w_hh[dir] + (size_t)((lr >> 5) * HID + rank * UNITS + (lr & 31)) * HID;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int k = 16 * kt + tig * 2 + 8 * hf;
                    const float2 v = __ldg(reinterpret_cast<const float2*>(wrow + k));
                    const float s0 = v.x * sc[rs], s1 = v.y * sc[rs];
                    const __half2 h = __floats2half2_rn(s0, s1);
                    const float2 b = __half22float2(h);
                    a_lo[kt][rs + 2 * hf] = pack_h2(s0 - b.x, s1 - b.y);
                    // hi plane: chunk k/64, row lr, 16-byte piece ((k%64)/8) ^ (lr%8), element k%8
                    uint8_t* p = smem + SM_WHI + (k >> 6) * 16384 + lr * 128 + ((((k & 63) >> 3) ^ (lr & 7)) << 4) + (k & 7) * 2;
                    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&h);
                }
            }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // W_hi: generic-proxy stores -> tensor-core reads
    if (tid == 0 && a.T >= 2)
        for (int g2 = 0; g2 < ng; ++g2) mbar_expect_tx(full_bar + g2 * 2 + 0, (uint32_t)OPB);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cl_sync();

    const int col = warp;                              // this thread's cell: (column `warp`, unit `lane`) of each group
    float unscale[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) unscale[g] = row_unscale[g * 32 + lane];
    float c_state[MAXG] = {0.f, 0.f};
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, M128 N16

    for (int step = 0; step < a.T; ++step) {
        const int t = t_first + step * tstep;
#pragma unroll
        for (int g2 = 0; g2 < MAXG; ++g2) {
            if (g2 >= ng) break;
            const int col0 = (g_first + g2) * NCOL;
            const int nb = min(NCOL, a.B - col0);
            float xp[4];
            {
                const float* xb = a.xproj + ((size_t)t * a.B + col0 + min(col, nb - 1)) * 4096 + dir * 2048 + rank * UNITS + lane;
#pragma unroll
                for (int g = 0; g < 4; ++g) xp[g] = __ldg(xb + g * HID);
            }
            if (step > 0) {
                const int par = (step - 1) & 1;
                mbar_wait(full_bar + g2 * 2 + par, ((step - 1) >> 1) & 1);
                const uint8_t* rb = smem + SM_RECV + (g2 * 2 + par) * OPB;
                if (tid == 0) {
                    // D[128 x 16] = W_hi[128 x 512] * [h_hi | h_lo][512 x 16]
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sA = smem_u32(smem + SM_WHI), sB = smem_u32(rb);
                    const uint32_t d = tmem_base + (uint32_t)(g2 * 16);
#pragma unroll 1
                    for (int j = 0; j < 32; ++j)
                        umma_f16(d, desc_sw128(sA + (j >> 2) * 16384) + (uint64_t)(((j & 3) * 32) >> 4),
                                 desc_interleave(sB + j * 512, 256, 128), IDESC, j != 0);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(dfull_bar + g2))
                                 : "memory");
                }
                // W_lo * h_hi on mma.sync: rows 16w .. 16w+15, four independent accumulators
                float d4[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d4[i][e] = 0.f;
#pragma unroll
                for (int kt = 0; kt < 32; ++kt) {
                    // h_hi of column gid: k = 16kt + 2tig (k-block 2kt), k + 8 (k-block 2kt + 1)
                    const uint8_t* hb = rb + (2 * kt) * 256 + gid * 16 + tig * 4;
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(hb);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(hb + 256);
                    mma16816(d4[kt & 3], a_lo[kt], b0, b1);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) d4[0][e] = (d4[0][e] + d4[1][e]) + (d4[2][e] + d4[3][e]);
                lo_buf[(tig * 2) * PCOL + 16 * warp + gid] = d4[0][0];
                lo_buf[(tig * 2 + 1) * PCOL + 16 * warp + gid] = d4[0][1];
                lo_buf[(tig * 2) * PCOL + 16 * warp + gid + 8] = d4[0][2];
                lo_buf[(tig * 2 + 1) * PCOL + 16 * warp + gid + 8] = d4[0][3];
                // tensor-core part: this warp may read TMEM lanes 32*(warp%4) .. +31 = gate warp%4, unit = lane
                mbar_wait(dfull_bar + g2, (step - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(g2 * 16), v);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int c = (warp >> 2) * 4 + c4;            // warps w and w+4 split the 8 columns
                    gate_buf[((warp & 3) * NCOL + c) * UNITS + lane] = __uint_as_float(v[c]) + __uint_as_float(v[8 + c]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
            }
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float sum = 0.f;
                if (step > 0) sum = gate_buf[(g * NCOL + col) * UNITS + lane] + lo_buf[col * PCOL + g * 32 + lane];
                pre[g] = fmaf(sum, unscale[g], xp[g]);
            }
            const float c_new = sigmoidf_(pre[1]) * c_state[g2] + sigmoidf_(pre[0]) * tanhf(pre[2]);
            c_state[g2] = c_new;
            const float h_new = sigmoidf_(pre[3]) * tanhf(c_new);
            if (col < nb) a.out[((size_t)t * a.B + col0 + col) * 1024 + dir * HID + rank * UNITS + lane] = h_new;
            {
                // staging block = four K-blocks of the peers' B operand: [kb = unit/8][plane][column][unit%8]
                const float hs = h_new * H_SCALE;
                const __half hh = __float2half_rn(hs);
                __half* sg = reinterpret_cast<__half*>(smem + SM_STAGE + (g2 * 2 + (step & 1)) * BLK + (lane >> 3) * 256 + col * 16) + (lane & 7);
                sg[0] = hh;
                sg[64] = __float2half_rn(hs - __half2float(hh));           // lo plane: + 128 B
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0 && step + 1 < a.T) {
                if (step + 2 < a.T) mbar_expect_tx(full_bar + g2 * 2 + ((step + 1) & 1), (uint32_t)OPB);
                const uint32_t dst = smem_u32(smem + SM_RECV + (g2 * 2 + (step & 1)) * OPB + rank * BLK);
                const uint32_t bar = smem_u32(full_bar + g2 * 2 + (step & 1));
                const uint8_t* src = smem + SM_STAGE + (g2 * 2 + (step & 1)) * BLK;
#pragma unroll 1
                for (uint32_t p = 0; p < (uint32_t)CL; ++p) cl_bulk_copy(cl_map(dst, p), src, BLK, cl_map(bar, p));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cl_sync();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
}

}  // namespace

// Same contract as lstm_layer_cluster() in lstm_cluster.cu.
int lstm_layer_cluster_tc(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, float* out, int T, int B,
                          int max_clusters, cudaStream_t st) {
    HN_CUDA_OK(cudaFuncSetAttribute(lstm_cluster_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    HN_CUDA_OK(cudaFuncSetAttribute(lstm_cluster_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    ClArgs a;
    a.xproj = xproj; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_bwd; a.out = out;
    a.T = T; a.B = B; a.ngroups = (B + NCOL - 1) / NCOL;
    a.gpc = (2 * a.ngroups <= max_clusters) ? 1 : MAXG;
    const int nclusters = 2 * ((a.ngroups + a.gpc - 1) / a.gpc);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CL);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    HN_CUDA_OK(cudaLaunchKernelEx(&cfg, lstm_cluster_tc_kernel, a));
    HN_LAUNCH_OK();
    return 0;
}

}  // namespace hn
