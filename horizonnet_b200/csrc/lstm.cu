// Persistent bidirectional LSTM recurrence (reference model.py:222-227 nn.LSTM, h0=c0=0,
// gate order i,f,g,o; c' = s(f)c + s(i)tanh(g), h' = s(o)tanh(c')).
//
// The input projections x_t W_ih^T + b_ih + b_hh for all 256 steps and both directions are one GEMM
// done beforehand (conv kernels, N = 4096 = [dir][gate][unit]); this kernel runs the 256 strictly
// sequential steps of one layer, both directions concurrently:
//   grid  = 2 directions x 64 CTAs, cooperative launch (all CTAs co-resident, 1 per SM)
//   CTA   = 8 hidden units x 4 gates = 32 rows of W_hh, held in REGISTERS for the whole sequence
//           (thread (row, kc) keeps 64 weights: k = 32*i + 4*kc + {0..3}), fp32 FMA (bit-faithful
//           to the fp32 reference up to summation order)
//   step  = h_{t-1} [32 batch x 512] is pulled from L2 into shared memory, each warp (= one hidden
//           unit, lanes = 4 gates x 8 k-slices) reduces with a transposing shuffle butterfly, the
//           cell update happens in registers (c never leaves the SM) and h_t is written straight
//           into the layer output [T][B][1024] (which is also where the other CTAs read it from);
//           a per-direction arrival counter in global memory orders the steps.
// Batches larger than 32 are processed in chunks of 32 (independent sequences).
#include <cooperative_groups.h>
#include "hn_common.cuh"

namespace hn {

namespace {

constexpr int HID = 512;
constexpr int NCTA_DIR = 64;               // CTAs per direction
constexpr int UNITS = HID / NCTA_DIR;      // 8 hidden units per CTA
constexpr int BCHUNK = 32;                 // batch columns per launch
constexpr long long SPIN_LIMIT_CYCLES = 4000000000ll;   // ~2 s: never hang the GPU on a logic bug

struct LstmArgs {
    const float* xproj;      // [T][B][4096]  (dir*2048 + gate*512 + unit), bias already added
    const float* w_hh[2];    // [2048][512] per direction (PyTorch layout, row = gate*512 + unit)
    float* out;              // [T][B][1024]  (dir*512 + unit)
    unsigned int* counters;  // [2] arrival counters, zeroed before launch
    int* error_flag;
    int T, B, b0, nb;        // batch chunk [b0, b0+nb)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256, 1) lstm_layer_kernel(const LstmArgs a) {
    extern __shared__ __align__(16) float hs[];          // [BCHUNK][HID] previous hidden state
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x / NCTA_DIR;
    const int cta = blockIdx.x % NCTA_DIR;
    const int gate = lane >> 3, kc = lane & 7;
    const int unit = cta * UNITS + warp;                   // hidden unit of this warp
    const int wrow = gate * HID + unit;                    // row of W_hh

    // W_hh slice -> registers: wreg[i*4+j] = W[wrow][32*i + 4*kc + j]
    float wreg[64];
    {
        const float4* wp = reinterpret_cast<const float4*>(a.w_hh[dir] + (size_t)wrow * HID);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 v = __ldg(wp + i * 8 + kc);
            wreg[i * 4 + 0] = v.x; wreg[i * 4 + 1] = v.y; wreg[i * 4 + 2] = v.z; wreg[i * 4 + 3] = v.w;
        }
    }
    // after the butterfly, lane (gate, kc) owns the sums of batch columns 4*kc .. 4*kc+3
    const int myb = kc * 4 + gate;                         // the batch column this lane updates
    const bool active = myb < a.nb;
    float c_state = 0.f;

    // prefetch of the x-projection for the first step
    auto xaddr = [&](int t, int g) {
        return a.xproj + ((size_t)t * a.B + a.b0 + myb) * 4096 + dir * 2048 + g * HID + unit;
    };
    int t = dir ? a.T - 1 : 0;
    const int tstep = dir ? -1 : 1;
    float xp[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
        for (int g = 0; g < 4; ++g) xp[g] = __ldg(xaddr(t, g));
    }

    for (int step = 0; step < a.T; ++step, t += tstep) {
        float acc[BCHUNK];
#pragma unroll
        for (int b = 0; b < BCHUNK; ++b) acc[b] = 0.f;
        if (step > 0) {
            // wait until all 64 CTAs of this direction have published h of the previous step
            if (tid == 0) {
                const unsigned int target = (unsigned int)(NCTA_DIR * step);
                const long long t0 = clock64();
                while (ld_acquire(a.counters + dir) < target) {
                    if (*reinterpret_cast<volatile int*>(a.error_flag) != 0) { s_abort = 1; break; }
                    if (clock64() - t0 > SPIN_LIMIT_CYCLES) { atomicExch(a.error_flag, 1); s_abort = 1; break; }
                }
            }
            __syncthreads();
            if (s_abort) return;             // a peer CTA never arrived: fail loudly on the host side
            const int tprev = t - tstep;
            // h_{t-1}: rows of 512 floats inside out[tprev][b][dir*512 ...] -> shared (L2 loads, no L1)
            for (int i = tid; i < BCHUNK * (HID / 4); i += 256) {
                const int b = i >> 7, q = i & 127;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < a.nb)
                    v = __ldcg(reinterpret_cast<const float4*>(
                            a.out + ((size_t)tprev * a.B + a.b0 + b) * 1024 + dir * HID) + q);
                reinterpret_cast<float4*>(hs)[i] = v;
            }
            __syncthreads();
            // partial dot products over this lane's k-slice, all 32 batch columns
#pragma unroll
            for (int b = 0; b < BCHUNK; ++b) {
                const float4* hp = reinterpret_cast<const float4*>(hs + b * HID);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 h = hp[i * 8 + kc];
                    s = fmaf(wreg[i * 4 + 0], h.x, s);
                    s = fmaf(wreg[i * 4 + 1], h.y, s);
                    s = fmaf(wreg[i * 4 + 2], h.z, s);
                    s = fmaf(wreg[i * 4 + 3], h.w, s);
                }
                acc[b] = s;
            }
        }
        // transposing butterfly over the 8 k-slices (lane bits 0..2): 32 -> 16 -> 8 -> 4 values
        float r16[16], r8[8], r4[4];
        {
            const bool up = kc & 4;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float send = up ? acc[i] : acc[i + 16];
                const float keep = up ? acc[i + 16] : acc[i];
                r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
            const bool up = kc & 2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float send = up ? r16[i] : r16[i + 8];
                const float keep = up ? r16[i + 8] : r16[i];
                r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        {
            const bool up = kc & 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = up ? r8[i] : r8[i + 4];
                const float keep = up ? r8[i + 4] : r8[i];
                r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        // r4[q] = W_hh[gate row] . h[b = kc*4 + q].  Gather the 4 gates of column q == gate:
        // lanes (g', kc) for g' = 0..3 exchange so that every lane gets all 4 gate values of ITS column.
        float gv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            // value wanted: gate g, column 'gate' (my own index) -> held by lane (g, kc) in r4[gate]
            const int src_lane = g * 8 + kc;
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // every lane offers r4[q] in round q; only the lane whose 'gate' == q reads it
                const float got = __shfl_sync(0xffffffffu, r4[q], src_lane);
                if (gate == q) v = got;
            }
            gv[g] = v;
        }
        float h_new = 0.f;
        if (active) {
            const float ig = sigmoidf_(gv[0] + xp[0]);
            const float fg = sigmoidf_(gv[1] + xp[1]);
            const float gg = tanhf(gv[2] + xp[2]);
            const float og = sigmoidf_(gv[3] + xp[3]);
            c_state = fg * c_state + ig * gg;
            h_new = og * tanhf(c_state);
            a.out[((size_t)t * a.B + a.b0 + myb) * 1024 + dir * HID + unit] = h_new;
        }
        // prefetch next step's x-projection while the other CTAs catch up
        if (active && step + 1 < a.T) {
#pragma unroll
            for (int g = 0; g < 4; ++g) xp[g] = __ldg(xaddr(t + tstep, g));
        }
        __syncthreads();                      // all h_t stores of this CTA issued
        if (tid == 0) {
            __threadfence();
            atomicAdd(a.counters + dir, 1u);
        }
    }
}

}  // namespace

// One LSTM layer, both directions.  xproj [T][B][4096], out [T][B][1024].
int lstm_layer(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, float* out, int T, int B,
               unsigned int* counters /* >= 2 uints */, int* error_flag, cudaStream_t st) {
    const size_t smem = (size_t)BCHUNK * HID * sizeof(float);
    HN_CUDA_OK(cudaFuncSetAttribute(lstm_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int b0 = 0; b0 < B; b0 += BCHUNK) {
        LstmArgs a;
        a.xproj = xproj; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_bwd; a.out = out;
        a.counters = counters; a.error_flag = error_flag;
        a.T = T; a.B = B; a.b0 = b0; a.nb = (B - b0 < BCHUNK) ? (B - b0) : BCHUNK;
        HN_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), st));
        void* args[] = {(void*)&a};
        HN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)lstm_layer_kernel, dim3(2 * NCTA_DIR), dim3(256),
                                               args, smem, st));
        HN_LAUNCH_OK();
    }
    return 0;
}

}  // namespace hn
