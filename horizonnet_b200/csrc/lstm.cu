// Persistent bidirectional LSTM recurrence (reference model.py:222-227 nn.LSTM, h0=c0=0,
// gate order i,f,g,o; c' = s(f)c + s(i)tanh(g), h' = s(o)tanh(c')).
//
// The input projections x_t W_ih^T + b_ih + b_hh for all 256 steps and both directions are one GEMM
// done beforehand (conv kernels, N = 4096 = [dir][gate][unit]); this kernel runs the 256 strictly
// sequential steps of one layer, both directions concurrently:
//   grid  = 2 directions x 64 CTAs, cooperative launch (all CTAs co-resident, 1 per SM)
//   CTA   = 8 hidden units x 4 gates = 32 rows of W_hh, held in REGISTERS for the whole sequence as
//           fp16 hi/lo mma.sync A-fragments (per-row power-of-two scale; 11+11 significant bits).
//   step  = every step needs h_{t-1} of all 512 units, i.e. an all-to-all between the 64 CTAs of a
//           direction through L2 (~2.5 us of latency), so the 32 batch columns of a launch are cut
//           into 4 sub-batches of 8 that are software-pipelined: while sub-batch j's h_t travels, the
//           CTA computes sub-batch j+1.
//   math  = per sub-step D[32 rows x 8 cols] = W[32 x 512] h[512 x 8] as mma.sync.m16n8k16 (fp16 in,
//           fp32 accumulate), three products Whi*hhi + Whi*hlo + Wlo*hhi (h travels as fp16 hi/lo planes
//           of 256*h), K split over the 8 compute warps, partials reduced through shared memory;
//           64 threads then finish one (unit, column) cell each in fp32.  (The fp32-FMA version of this
//           mat-vec was instruction-issue bound: 670 instructions per warp and sub-step, ncu.)
//   warps = 0-7 compute; 8 loads (acquire-polls the arrival counter of the next sub-step, pulls the
//           h planes with 1-D TMA bulk copies onto an mbarrier); 9 signals (fence + red.release once
//           the cells of a sub-step are stored) -- the gpu-scope fence must not sit in the compute path.
// Batches larger than 32 are processed in chunks of 32 (independent sequences).
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include "hn_common.cuh"
#include "ptx.cuh"

namespace hn {

namespace {

constexpr int HID = 512;
constexpr int NCTA_DIR = 64;               // CTAs per direction
constexpr int UNITS = HID / NCTA_DIR;      // 8 hidden units per CTA
constexpr int BCHUNK = 32;                 // batch columns per launch
constexpr int NSB = 4;                     // sub-batches pipelined through a CTA
constexpr int SBC = 8;                     // columns per sub-batch = N of the MMA
constexpr int NCOMPUTE = 256;              // threads of the 8 compute warps
constexpr int NTHREADS = NCOMPUTE + 64;    // + loader warp + signalling warp
constexpr int COLB = 2064;                 // shared bytes per column: 1024 B hi + 1024 B lo + 16 B pad (bank spread)
constexpr float H_SCALE = 256.f;           // h planes carry 256*h (|h| < 1): lo stays in fp16's normal range
constexpr long long SPIN_LIMIT_CYCLES = 4000000000ll;   // ~2 s: never hang the GPU on a logic bug

// shared memory carve-up
constexpr int SM_H = 0;                                  // [NSB][SBC][COLB]
constexpr int SM_PART = SM_H + NSB * SBC * COLB;         // [2][8 warps][32 rows][8 cols] fp32 partial sums
constexpr int SM_ROW = SM_PART + 2 * 8 * 256 * 4;        // [8 warps][32 rows] row maxima, then [32] unscale factors
constexpr int SM_BAR = SM_ROW + 8 * 32 * 4;              // mbarriers
constexpr int SM_TOTAL = SM_BAR + 3 * NSB * 8;

struct LstmArgs {
    const float* xproj;      // [T][B][4096]  (dir*2048 + gate*512 + unit), bias already added
    const float* w_hh[2];    // [2048][512] per direction (PyTorch layout, row = gate*512 + unit)
    float* out;              // [T][B][1024]  (dir*512 + unit)
    unsigned int* counters;  // [2][NSB] arrival counters, zeroed before launch
    __half* hx;              // exchange planes [2 parity][2 dir][32 cols][2 planes][512] of 256*h
    int* error_flag;
    int T, B, b0, nb;        // batch chunk [b0, b0+nb)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
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
// power-of-two scale that puts max|w| of a row into (2^13, 2^14]
__device__ __forceinline__ float row_scale(float absmax) {
    if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
    int e;
    frexpf(16384.f / absmax, &e);
    return ldexpf(1.f, e - 1);
}

__global__ void __launch_bounds__(NTHREADS, 1) lstm_layer_kernel(const LstmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);   // TMA landed
    uint64_t* empty_bar = full_bar + NSB;                               // compute warps finished reading h
    uint64_t* done_bar = empty_bar + NSB;                               // cells of the sub-step stored
    float* part = reinterpret_cast<float*>(smem + SM_PART);
    float* rowbuf = reinterpret_cast<float*>(smem + SM_ROW);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x / NCTA_DIR;
    const int cta = blockIdx.x % NCTA_DIR;
    const int nsb = (a.nb + SBC - 1) / SBC;                 // active sub-batches of this chunk
    const int tstep = dir ? -1 : 1;
    const int t_first = dir ? a.T - 1 : 0;
    unsigned int* ctr = a.counters + dir * NSB;

    if (tid == 0) {
        for (int j = 0; j < NSB; ++j) { mbar_init(full_bar + j, 1); mbar_init(empty_bar + j, 8); mbar_init(done_bar + j, 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // columns of partially filled sub-batches are never written by TMA: keep them zero
    for (int i = tid; i < NSB * SBC * COLB / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + SM_H)[i] = 0u;

    // ---- W_hh slice -> fp16 hi/lo A-fragments (compute warps).  Warp w owns k in [64w, 64w+64).
    // fragment value e of m-tile m, k-tile kt:  row = (e>>1 & 1)*8 + lane/4,  k = 64w + 16kt + (lane&3)*2 + (e&1) + 8*(e>>2)
    // local row r = 16m + row  <->  gate 2m + (e>>1 & 1), unit lane/4
    const int gid = lane >> 2, tig = lane & 3;
    uint32_t a_hi[2][4][4], a_lo[2][4][4];
    if (warp < 8) {
        // per-row absolute maximum over all 512 k (8 warps x 4 lanes hold pieces of each row)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
                float mx = 0.f;
#pragma unroll
                for (int kt = 0; kt < 4; ++kt)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int k = 64 * warp + 16 * kt + tig * 2 + 8 * hf;
                        const float2 v = __ldg(reinterpret_cast<const float2*>(
                            a.w_hh[dir] + (size_t)((2 * m + rs) * HID + cta * UNITS + gid) * HID + k));
                        mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
                    }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                if (tig == 0) rowbuf[warp * 32 + (2 * m + rs) * 8 + gid] = mx;       // row index r = gate*8 + unit
            }
    }
    __syncthreads();
    float my_unscale = 0.f;                        // threads < 32: 1 / (row scale * H_SCALE) of row tid
    if (warp < 8) {
        if (tid < 32) {
            float mx = 0.f;
            for (int w8 = 0; w8 < 8; ++w8) mx = fmaxf(mx, rowbuf[w8 * 32 + tid]);
            my_unscale = 1.f / (row_scale(mx) * H_SCALE);
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
                float mx = 0.f;
                for (int w8 = 0; w8 < 8; ++w8) mx = fmaxf(mx, rowbuf[w8 * 32 + (2 * m + rs) * 8 + gid]);
                const float sc = row_scale(mx);
#pragma unroll
                for (int kt = 0; kt < 4; ++kt)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int k = 64 * warp + 16 * kt + tig * 2 + 8 * hf;
                        const float2 v = __ldg(reinterpret_cast<const float2*>(
                            a.w_hh[dir] + (size_t)((2 * m + rs) * HID + cta * UNITS + gid) * HID + k));
                        const float s0 = v.x * sc, s1 = v.y * sc;
                        const __half2 h = __floats2half2_rn(s0, s1);
                        const float2 b = __half22float2(h);
                        // A-fragment register index: a0a1 (row gid, k lo), a2a3 (row gid+8, k lo), a4a5 (row gid, k hi), a6a7
                        a_hi[m][kt][rs + 2 * hf] = *reinterpret_cast<const uint32_t*>(&h);
                        a_lo[m][kt][rs + 2 * hf] = pack_h2(s0 - b.x, s1 - b.y);
                    }
            }
    }
    __syncthreads();
    if (tid < 32) rowbuf[tid] = my_unscale;        // rowbuf[0..31] = unscale factor of local row r
    __syncthreads();

    if (warp == 8) {
        // =============================== loader warp ===============================
        if (lane == 0) {
            for (int step = 1; step < a.T; ++step) {
                const __half* src = a.hx + ((size_t)((step - 1) & 1) * 2 + dir) * (32 * 2 * HID);
                for (int j = 0; j < nsb; ++j) {
                    mbar_wait(empty_bar + j, ((step - 1) & 1) ^ 1);            // buffer j free again
                    const unsigned int target = (unsigned int)(NCTA_DIR * step);
                    if (ld_acquire(ctr + j) < target) {
                        const long long t0 = clock64();
                        while (ld_acquire(ctr + j) < target) {
                            if (*reinterpret_cast<volatile int*>(a.error_flag) != 0) break;
                            if (clock64() - t0 > SPIN_LIMIT_CYCLES) { atomicExch(a.error_flag, 1); break; }
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");           // generic-proxy acquire -> async-proxy reads
                    const int ncol = min(SBC, a.nb - j * SBC);
                    mbar_expect_tx(full_bar + j, (uint32_t)(ncol * 2 * HID * sizeof(__half)));
                    for (int c = 0; c < ncol; ++c)
                        bulk_load_1d(smem + SM_H + (j * SBC + c) * COLB, src + (size_t)(j * SBC + c) * (2 * HID),
                                     2 * HID * sizeof(__half), full_bar + j);
                }
            }
        }
        return;
    }
    if (warp == 9) {
        // =============================== signalling warp ===============================
        if (lane == 0) {
            for (int step = 0; step < a.T; ++step)
                for (int j = 0; j < nsb; ++j) {
                    mbar_wait(done_bar + j, step & 1);          // the 64 cells of (step, j) are stored
                    __threadfence();                            // (measured: 12 % faster than red.release alone)
                    red_release_add(ctr + j, 1u);
                }
        }
        return;
    }

    // =============================== compute warps ===============================
    // finishing threads: tid < 64 owns the cell (unit fu, column fc) of every sub-batch
    const int fu = tid >> 3, fc = tid & 7;
    float c_state[NSB];
#pragma unroll
    for (int j = 0; j < NSB; ++j) c_state[j] = 0.f;
    float unscale[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid < 64) {
#pragma unroll
        for (int g = 0; g < 4; ++g) unscale[g] = rowbuf[g * 8 + fu];
    }
    int pbuf = 0;                                  // partial-sum buffer parity (one per executed MMA sub-step)

    for (int step = 0; step < a.T; ++step) {
        const int t = t_first + step * tstep;
#pragma unroll
        for (int j = 0; j < NSB; ++j) {
            if (j >= nsb) break;
            const int col = j * SBC + fc;
            float xp[4] = {0.f, 0.f, 0.f, 0.f};
            if (tid < 64) {        // x-projection of this cell (latency hides behind the MMAs)
                const float* xb = a.xproj + ((size_t)t * a.B + a.b0 + min(col, a.nb - 1)) * 4096 + dir * 2048 + cta * UNITS + fu;
#pragma unroll
                for (int g = 0; g < 4; ++g) xp[g] = __ldg(xb + g * HID);
            }
            float* pb = part + pbuf * (8 * 256);
            if (step > 0) {
                float d[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][e] = 0.f;
                mbar_wait(full_bar + j, (step - 1) & 1);
                // B fragments: column gid, k = 64w + 16kt + tig*2 (+8): one 32-bit word each from the hi / lo plane
                const uint8_t* hb = smem + SM_H + (j * SBC + gid) * COLB + (64 * warp + tig * 2) * 2;
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) {
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(hb + kt * 32);
                    const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(hb + kt * 32 + 16);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(hb + 1024 + kt * 32);
                    const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(hb + 1024 + kt * 32 + 16);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        mma16816(d[m], a_hi[m][kt], bh0, bh1);
                        mma16816(d[m], a_hi[m][kt], bl0, bl1);
                        mma16816(d[m], a_lo[m][kt], bh0, bh1);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar + j);                     // this warp is done with the h buffer
                // partial sums of this warp's k-range: rows 16m + gid (+8), columns tig*2 (+1)
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    *reinterpret_cast<float2*>(pb + warp * 256 + (16 * m + gid) * 8 + tig * 2) = make_float2(d[m][0], d[m][1]);
                    *reinterpret_cast<float2*>(pb + warp * 256 + (16 * m + gid + 8) * 8 + tig * 2) = make_float2(d[m][2], d[m][3]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory");
            }
            if (tid < 64) {
                float pre[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float s = 0.f;
                    if (step > 0) {
#pragma unroll
                        for (int w8 = 0; w8 < 8; ++w8) s += pb[w8 * 256 + (g * 8 + fu) * 8 + fc];
                    }
                    pre[g] = fmaf(s, unscale[g], xp[g]);
                }
                const float c_new = sigmoidf_(pre[1]) * c_state[j] + sigmoidf_(pre[0]) * tanhf(pre[2]);
                c_state[j] = c_new;
                const float h_new = sigmoidf_(pre[3]) * tanhf(c_new);
                if (col < a.nb) {
                    a.out[((size_t)t * a.B + a.b0 + col) * 1024 + dir * HID + cta * UNITS + fu] = h_new;
                    // exchange planes of 256*h for the next step's MMAs
                    const float hs = h_new * H_SCALE;
                    const __half hh = __float2half_rn(hs);
                    __half* hx = a.hx + (((size_t)(step & 1) * 2 + dir) * 32 + col) * (2 * HID) + cta * UNITS + fu;
                    hx[0] = hh;
                    hx[HID] = __float2half_rn(hs - __half2float(hh));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(done_bar + j);
            }
            if (step > 0) pbuf ^= 1;
        }
    }
}

}  // namespace

size_t lstm_scratch_bytes() { return 1024 + (size_t)2 * 2 * 32 * 2 * HID * sizeof(__half); }

// One LSTM layer, both directions.  xproj [T][B][4096], out [T][B][1024].
// scratch: lstm_scratch_bytes() bytes (arrival counters in the first 1 KB, then the h exchange planes).
int lstm_layer_cluster(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, float* out, int T, int B,
                       cudaStream_t st);      // lstm_cluster.cu

int lstm_layer(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, float* out, int T, int B,
               void* scratch, int* error_flag, cudaStream_t st) {
    // preferred: 16-CTA clusters exchanging h through distributed shared memory (HN_LSTM_CLUSTER=0: the L2-exchange
    // kernel below, which is also the fallback when the device cannot co-schedule such clusters)
    static const bool cluster_on = [] { const char* e = getenv("HN_LSTM_CLUSTER"); return !(e && atoi(e) == 0); }();
    if (cluster_on) {
        const int r = lstm_layer_cluster(xproj, w_hh_fwd, w_hh_bwd, out, T, B, st);
        if (r <= 0) return r;
    }
    HN_CUDA_OK(cudaFuncSetAttribute(lstm_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    unsigned int* counters = reinterpret_cast<unsigned int*>(scratch);
    for (int b0 = 0; b0 < B; b0 += BCHUNK) {
        LstmArgs a;
        a.xproj = xproj; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_bwd; a.out = out;
        a.counters = counters; a.error_flag = error_flag;
        a.hx = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(scratch) + 1024);
        a.T = T; a.B = B; a.b0 = b0; a.nb = (B - b0 < BCHUNK) ? (B - b0) : BCHUNK;
        HN_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * NSB * sizeof(unsigned int), st));
        void* args[] = {(void*)&a};
        HN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)lstm_layer_kernel, dim3(2 * NCTA_DIR), dim3(NTHREADS), args,
                                               SM_TOTAL, st));
        HN_LAUNCH_OK();
    }
    return 0;
}

}  // namespace hn
