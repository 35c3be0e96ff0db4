// Persistent bidirectional LSTM recurrence (reference model.py:222-227 nn.LSTM, h0=c0=0,
// gate order i,f,g,o; c' = s(f)c + s(i)tanh(g), h' = s(o)tanh(c')).
//
// The input projections x_t W_ih^T + b_ih + b_hh for all 256 steps and both directions are one GEMM
// done beforehand (conv kernels, N = 4096 = [dir][gate][unit]); this kernel runs the 256 strictly
// sequential steps of one layer, both directions concurrently:
//   grid  = 2 directions x 64 CTAs, cooperative launch (all CTAs co-resident, 1 per SM)
//   CTA   = 8 hidden units x 4 gates = 32 rows of W_hh, held in REGISTERS for the whole sequence
//           (compute warp = one hidden unit, lane = (gate, k-slice); 64 weights per thread), fp32 FMA:
//           bit-faithful to the fp32 reference up to summation order.
//   step  = every step needs h_{t-1} of all 512 units, i.e. an all-to-all between the 64 CTAs of a
//           direction through L2.  That exchange is latency (~2.5 us), so the 32 batch columns of a
//           launch are cut into NSB independent sub-batches that are software-pipelined: while
//           sub-batch j's h_t travels, the compute warps work on sub-batch j+1.
//           Warp roles: warps 0-7 compute (FMA from shared memory, transposing shuffle butterfly,
//           cell update in registers, h_t stored straight into the layer output [T][B][1024]) and
//           never wait for anything but their operands; warp 8 loads (acquire-polls the arrival
//           counter of the next sub-step and pulls h_{t-1} [SBC x 512] into shared memory with
//           1-D TMA bulk copies that complete on an mbarrier); warp 9 signals (waits until the 8
//           compute warps have stored h_t of a sub-step, then fence + red.release on the counter --
//           the gpu-scope fence costs ~1 us and must not sit in the compute warps' path).
// Batches larger than 32 are processed in chunks of 32 (independent sequences).
#include <cooperative_groups.h>
#include "hn_common.cuh"
#include "ptx.cuh"

namespace hn {

namespace {

constexpr int HID = 512;
constexpr int NCTA_DIR = 64;               // CTAs per direction
constexpr int UNITS = HID / NCTA_DIR;      // 8 hidden units per CTA
constexpr int BCHUNK = 32;                 // batch columns per launch
constexpr int NSB = 4;                     // sub-batches pipelined through a CTA
constexpr int SBC = BCHUNK / NSB;          // columns per sub-batch
constexpr int FC = SBC / 8;                // columns a lane finishes after the butterfly
constexpr int NCOMPUTE = 256;              // threads of the 8 compute warps
constexpr int NTHREADS = NCOMPUTE + 64;    // + loader warp + signalling warp
constexpr long long SPIN_LIMIT_CYCLES = 4000000000ll;   // ~2 s: never hang the GPU on a logic bug

struct LstmArgs {
    const float* xproj;      // [T][B][4096]  (dir*2048 + gate*512 + unit), bias already added
    const float* w_hh[2];    // [2048][512] per direction (PyTorch layout, row = gate*512 + unit)
    float* out;              // [T][B][1024]  (dir*512 + unit)
    unsigned int* counters;  // [2][NSB] arrival counters, zeroed before launch
    int* error_flag;
    int T, B, b0, nb;        // batch chunk [b0, b0+nb)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// two independent fp32 FMAs in one instruction (sm_100+): d.lo += a.lo*b.lo, d.hi += a.hi*b.hi
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) lstm_layer_kernel(const LstmArgs a) {
    extern __shared__ __align__(128) uint8_t lstm_smem[];
    float (*hs)[SBC * HID] = reinterpret_cast<float (*)[SBC * HID]>(lstm_smem);   // h_{t-1} of each sub-batch
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(lstm_smem + sizeof(float) * NSB * SBC * HID);   // TMA landed
    uint64_t* empty_bar = full_bar + NSB;                   // compute warps finished reading
    uint64_t* done_bar = empty_bar + NSB;                   // compute warps stored h_t of the sub-step
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x / NCTA_DIR;
    const int cta = blockIdx.x % NCTA_DIR;
    const int nsb = (a.nb + SBC - 1) / SBC;                 // active sub-batches of this chunk
    const int tstep = dir ? -1 : 1;
    const int t_first = dir ? a.T - 1 : 0;
    unsigned int* ctr = a.counters + dir * NSB;

    if (tid == 0) {
        for (int j = 0; j < NSB; ++j) { mbar_init(full_bar + j, 1); mbar_init(empty_bar + j, 8); mbar_init(done_bar + j, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows of partially filled sub-batches are never written by TMA: keep them zero
    for (int i = tid; i < NSB * SBC * HID; i += NTHREADS) (&hs[0][0])[i] = 0.f;
    __syncthreads();

    if (warp == 8) {
        // =============================== communication warp ===============================
        if (lane == 0) {
            for (int step = 1; step < a.T; ++step) {
                const int tprev = t_first + (step - 1) * tstep;
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
                    mbar_expect_tx(full_bar + j, (uint32_t)(ncol * HID * sizeof(float)));
                    for (int b = 0; b < ncol; ++b)
                        bulk_load_1d(&hs[j][b * HID],
                                     a.out + ((size_t)tprev * a.B + a.b0 + j * SBC + b) * 1024 + dir * HID,
                                     HID * sizeof(float), full_bar + j);
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
                    mbar_wait(done_bar + j, step & 1);          // all 8 compute warps stored h_t of (step, j)
                    __threadfence();
                    red_release_add(ctr + j, 1u);
                }
        }
        return;
    }

    // =============================== compute warps ===============================
    // warp = one hidden unit; lane = one of 32 k-slices (k = 128*i + 4*lane + e) holding the weights of all
    // 4 gates for that slice.  (128-bit shared loads are issued per quarter-warp, so a layout where the
    // 4 gate groups of a warp re-read the same h values costs 4x the shared-memory wavefronts: measured
    // 16.8 K wavefronts per step, 43 % short-scoreboard stalls.  Here every lane reads distinct data.)
    const int unit = cta * UNITS + warp;                   // hidden unit of this warp
    // weights as packed fp32 pairs for fma.rn.f32x2 (Blackwell: two fp32 FMAs per issue slot):
    // wreg[g][2*i+p] = (W_hh[g*512+unit][k], W[..][k+1]),  k = 128*i + 4*lane + 2*p
    unsigned long long wreg[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(a.w_hh[dir] + (size_t)(g * HID + unit) * HID);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const ulonglong2 v = __ldg(wp + i * 32 + lane);
            wreg[g][2 * i + 0] = v.x; wreg[g][2 * i + 1] = v.y;
        }
    }
    const int gate = lane >> 3, kc = lane & 7;             // after the butterfly: lane = (gate, column kc)
    float c_state[NSB][FC];
#pragma unroll
    for (int j = 0; j < NSB; ++j)
#pragma unroll
        for (int i = 0; i < FC; ++i) c_state[j][i] = 0.f;

    for (int step = 0; step < a.T; ++step) {
        const int t = t_first + step * tstep;
#pragma unroll
        for (int j = 0; j < NSB; ++j) {
            if (j >= nsb) break;
            // x-projection of the column this lane finishes (all 4 gates; latency hides behind the FMAs)
            float xp[FC][4];
#pragma unroll
            for (int i = 0; i < FC; ++i) {
                const int col = j * SBC + kc * FC + i;
                const float* xb = a.xproj + ((size_t)t * a.B + a.b0 + min(col, a.nb - 1)) * 4096 + dir * 2048 + unit;
#pragma unroll
                for (int g = 0; g < 4; ++g) xp[i][g] = __ldg(xb + g * HID);
            }
            // acc[g*SBC + b]: partial dot product of gate g with column b over this lane's k-slice
            float acc[4 * SBC];
            if (step > 0) {
                unsigned long long acc2[4 * SBC];          // (even-k partial sum, odd-k partial sum)
#pragma unroll
                for (int v = 0; v < 4 * SBC; ++v) acc2[v] = 0ull;
                mbar_wait(full_bar + j, (step - 1) & 1);
#pragma unroll
                for (int b = 0; b < SBC; ++b) {
                    const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(&hs[j][b * HID]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const ulonglong2 h = hp[i * 32 + lane];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            fma2(acc2[g * SBC + b], wreg[g][2 * i + 0], h.x);
                            fma2(acc2[g * SBC + b], wreg[g][2 * i + 1], h.y);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar + j);                     // this warp is done with hs[j]
#pragma unroll
                for (int v = 0; v < 4 * SBC; ++v)
                    acc[v] = __uint_as_float((unsigned)(acc2[v] & 0xffffffffull)) + __uint_as_float((unsigned)(acc2[v] >> 32));
            } else {
#pragma unroll
                for (int v = 0; v < 4 * SBC; ++v) acc[v] = 0.f;
            }
            // transposing butterfly over the 32 k-slices (lane bits 4..0): 4*SBC -> FC values per lane;
            // lane L ends with value indices L*FC .. L*FC+FC-1, i.e. gate L/8, columns (L%8)*FC + i
#pragma unroll
            for (int sh = 16, n = 2 * SBC; sh >= 1; sh >>= 1, n >>= 1) {
                const bool up = lane & sh;
#pragma unroll
                for (int i = 0; i < 2 * SBC; ++i) {
                    if (i < n) {
                        const float send = up ? acc[i] : acc[i + n];
                        const float keep = up ? acc[i + n] : acc[i];
                        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, sh);
                    }
                }
            }
            // all 4 gate lanes of a column fetch the 4 gate sums and update the cell redundantly
#pragma unroll
            for (int i = 0; i < FC; ++i) {
                const float gi = __shfl_sync(0xffffffffu, acc[i], kc) + xp[i][0];
                const float gf = __shfl_sync(0xffffffffu, acc[i], 8 + kc) + xp[i][1];
                const float gg = __shfl_sync(0xffffffffu, acc[i], 16 + kc) + xp[i][2];
                const float go = __shfl_sync(0xffffffffu, acc[i], 24 + kc) + xp[i][3];
                const float c_new = sigmoidf_(gf) * c_state[j][i] + sigmoidf_(gi) * tanhf(gg);
                c_state[j][i] = c_new;
                const float h_new = sigmoidf_(go) * tanhf(c_new);
                const int col = j * SBC + kc * FC + i;
                if (gate == 0 && col < a.nb)
                    a.out[((size_t)t * a.B + a.b0 + col) * 1024 + dir * HID + unit] = h_new;
            }
            // this warp's h_t stores are issued -> tell the signalling warp
            __syncwarp();
            if (lane == 0) mbar_arrive(done_bar + j);
        }
    }
}

}  // namespace

// One LSTM layer, both directions.  xproj [T][B][4096], out [T][B][1024].
int lstm_layer(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, float* out, int T, int B,
               unsigned int* counters /* >= 8 uints */, int* error_flag, cudaStream_t st) {
    const size_t smem = sizeof(float) * NSB * SBC * HID + 3 * NSB * sizeof(uint64_t);
    HN_CUDA_OK(cudaFuncSetAttribute(lstm_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int b0 = 0; b0 < B; b0 += BCHUNK) {
        LstmArgs a;
        a.xproj = xproj; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_bwd; a.out = out;
        a.counters = counters; a.error_flag = error_flag;
        a.T = T; a.B = B; a.b0 = b0; a.nb = (B - b0 < BCHUNK) ? (B - b0) : BCHUNK;
        HN_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * NSB * sizeof(unsigned int), st));
        void* args[] = {(void*)&a};
        HN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)lstm_layer_kernel, dim3(2 * NCTA_DIR),
                                               dim3(NTHREADS), args, smem, st));
        HN_LAUNCH_OK();
    }
    return 0;
}

}  // namespace hn
