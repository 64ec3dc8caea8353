// Do the FP32 (packed FFMA2) and FP64 (DFMA) pipes of an sm_100 SM run side by
// side?  One block of 8 warps per SM (2 per scheduler): in the mixed kernel the
// warps 0-3 run FFMA2 chains and the warps 4-7 DFMA chains, each sized to take
// about the same time alone.  Prints lane-FMA/clk/SM for FFMA2 alone, DFMA
// alone and both together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mix scripts/ubench_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8, UN = 16;

template <int MODE>  // 0 = all FFMA2, 1 = all DFMA, 2 = warps 0-3 FFMA2 + warps 4-7 DFMA
__global__ void __launch_bounds__(256, 1) mix(const float *in, int trips32, int trips64, float *out)
{
    const int warp = threadIdx.x >> 5;
    const float a0 = in[threadIdx.x & 31], b0 = in[32 + (threadIdx.x & 31)];
    const bool f32 = MODE == 0 || (MODE == 2 && warp < 4);
    if (f32) {
        float2 x[CH];
        const float2 a2 = make_float2(a0, a0 + 1e-7f), b2 = make_float2(b0, b0 - 1e-7f);
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = make_float2(a0 + c, b0 + c);
        for (int t = 0; t < trips32; ++t)
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int c = 0; c < CH; ++c) x[c] = __ffma2_rn(x[c], a2, b2);
        float s = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) s += x[c].x + x[c].y;
        if (s == 123.456f) out[threadIdx.x] = s;
    } else {
        double x[CH];
        const double a = a0, b = b0;
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = a + c;
        for (int t = 0; t < trips64; ++t)
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
        double s = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) s += x[c];
        if (s == 123.456) out[threadIdx.x] = (float)s;
    }
}

template <int MODE>
static float run(const float *in, int t32, int t64, float *out, int sms, int threads = 256)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix<MODE><<<sms, threads>>>(in, t32, t64, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    float h[64];
    for (int i = 0; i < 32; ++i) { h[i] = 0.999f + 1e-5f * i; h[32 + i] = 1e-3f * (i + 1); }
    float *in, *out;
    cudaMalloc(&in, sizeof(h));
    cudaMalloc(&out, 4096);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int t32 = 4096, t64 = 4096;
    const double per_trip = (double)CH * UN * 32;  // lane-FMAs per warp per trip (x2 for FFMA2)
    float ms = run<0>(in, t32, t64, out, sms);
    printf("FFMA2 alone (8 warps/SM): %.1f lane-FMA/clk/SM (%.3f ms)\n",
           8 * t32 * per_trip * 2 / (ms * 1e-3) / clk, ms);
    ms = run<1>(in, t32, t64, out, sms);
    printf("DFMA alone  (8 warps/SM): %.1f lane-FMA/clk/SM (%.3f ms)\n",
           8 * t64 * per_trip / (ms * 1e-3) / clk, ms);
    // 4 + 4 warps, each half sized to take the time of the pure run at 8 warps / 2
    for (int r64 = 1; r64 <= 2; ++r64) {
        const int m32 = t32, m64 = t64 / r64;
        ms = run<2>(in, m32, m64, out, sms);
        printf("mixed 4 FFMA2 warps x %d trips + 4 DFMA warps x %d trips: %.3f ms -> FFMA2 %.1f + DFMA %.1f "
               "lane-FMA/clk/SM\n", m32, m64, ms, 4 * m32 * per_trip * 2 / (ms * 1e-3) / clk,
               4 * m64 * per_trip / (ms * 1e-3) / clk);
    }
    // ONE warp per scheduler: can a lone warp keep the pipe busy?
    for (int w = 1; w <= 8; w *= 2) {
        ms = run<0>(in, t32, t64, out, sms, 32 * w);
        printf("FFMA2, %d warp(s)/SM: %.1f lane-FMA/clk/SM\n", w, w * t32 * per_trip * 2 / (ms * 1e-3) / clk);
    }
    for (int w = 1; w <= 8; w *= 2) {
        ms = run<1>(in, t32, t64, out, sms, 32 * w);
        printf("DFMA,  %d warp(s)/SM: %.1f lane-FMA/clk/SM\n", w, w * t64 * per_trip / (ms * 1e-3) / clk);
    }
    cudaFree(in);
    cudaFree(out);
    return 0;
}
