// iid_ubench.cuh -- issue-rate micro-benchmarks of the pipes the pair-sum
// kernels are bound by (sm_100a): scalar FFMA, packed FFMA2 and DFMA.  bench.py
// reports the kernels' executed-instruction rates against these MEASURED peaks
// (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers).
#pragma once
#include <cuda_runtime.h>

namespace iid {

enum { UB_FFMA = 0, UB_FFMA2 = 1, UB_DFMA = 2 };
constexpr int UB_CHAINS = 8;    // independent dependency chains per thread
constexpr int UB_UNROLL = 16;   // FMAs per chain per loop trip

// Every thread runs UB_CHAINS independent FMA chains x = x * a + b with
// per-lane operands (as in the real kernels, nothing is uniform).
template <int KIND>
__global__ void __launch_bounds__(256) pipe_peak_kernel(const float *__restrict__ in, int trips,
                                                        float *__restrict__ out)
{
    const float a0 = in[threadIdx.x & 31], b0 = in[32 + (threadIdx.x & 31)];
    if constexpr (KIND == UB_FFMA) {
        float x[UB_CHAINS];
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) x[c] = a0 + (float)c;
        for (int t = 0; t < trips; ++t)
#pragma unroll
            for (int u = 0; u < UB_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < UB_CHAINS; ++c) x[c] = fmaf(x[c], a0, b0);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) s += x[c];
        if (s == 123.456f) out[threadIdx.x] = s;
    } else if constexpr (KIND == UB_FFMA2) {
        float2 x[UB_CHAINS];
        const float2 a2 = make_float2(a0, a0 + 1e-7f), b2 = make_float2(b0, b0 - 1e-7f);
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) x[c] = make_float2(a0 + (float)c, b0 + (float)c);
        for (int t = 0; t < trips; ++t)
#pragma unroll
            for (int u = 0; u < UB_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < UB_CHAINS; ++c) x[c] = __ffma2_rn(x[c], a2, b2);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) s += x[c].x + x[c].y;
        if (s == 123.456f) out[threadIdx.x] = s;
    } else {
        double x[UB_CHAINS];
        const double a = (double)a0, b = (double)b0;
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) x[c] = a + (double)c;
        for (int t = 0; t < trips; ++t)
#pragma unroll
            for (int u = 0; u < UB_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < UB_CHAINS; ++c) x[c] = fma(x[c], a, b);
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < UB_CHAINS; ++c) s += x[c];
        if (s == 123.456) out[threadIdx.x] = (float)s;
    }
}

}  // namespace iid
