// Micro-benchmarks of the FP32 / FP64 / SFU issue rates on the GPU box.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench scripts/ubench.cu
// Prints thread-instructions per clock per SM for several instruction mixes.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

template <int KIND>
__global__ void __launch_bounds__(256) bench(float *out, long long *cyc, float a, float b)
{
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x * 0.001f + k;
    float x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = a + k * 0.25f; y[k] = b - k * 0.125f; }
    float2 acc2[8], a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc2[k] = make_float2(acc[2 * k], acc[2 * k + 1]);
    double dacc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dacc[k] = acc[k];
    double da = a, db = b;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (KIND == 0) {  // FFMA, two loop-invariant operands
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
        } else if (KIND == 1) {  // FFMA, three distinct registers each
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fmaf(x[k & 7], y[(k * 3 + 1) & 7], acc[k]);
        } else if (KIND == 2) {  // FFMA2 packed
#pragma unroll
            for (int k = 0; k < 8; ++k) acc2[k] = __ffma2_rn(acc2[k], a2, b2);
        } else if (KIND == 3) {  // DFMA
#pragma unroll
            for (int k = 0; k < 8; ++k) dacc[k] = fma(dacc[k], da, db);
        } else if (KIND == 4) {  // MUFU.SIN (via __sinf: FMUL + MUFU)
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = __sinf(acc[k]);
        } else if (KIND == 5) {  // FMUL + FADD alternating (no FMA)
#pragma unroll
            for (int k = 0; k < 16; k += 2) { acc[k] = acc[k] * a; acc[k + 1] = acc[k + 1] + b; }
        } else if (KIND == 6) {  // FFMA2 with three distinct packed registers
#pragma unroll
            for (int k = 0; k < 8; ++k)
                acc2[k] = __ffma2_rn(make_float2(x[k], y[k]), make_float2(y[(k + 3) & 7], x[(k + 5) & 7]), acc2[k]);
        } else if (KIND == 7) {  // mix: 1 MUFU per 8 FFMA
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
            acc[0] = __sinf(acc[0]);
            acc[8] = __cosf(acc[8]);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc2[k].x + acc2[k].y + (float)dacc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// The Debye inner step, scalar: 10 FP32 instructions per bin.
__global__ void __launch_bounds__(256) step_scalar(float *out, long long *cyc, float cth, float sth,
                                                   float kap, float r2, float dx, float dy, float dz)
{
    float F[32], X[32], Y[32], Z[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) { F[m] = 0; X[m] = 0; Y[m] = 0; Z[m] = 0; }
    float s = threadIdx.x * 1e-3f, c = 1.f - s;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; ++it) {
        float mk = kap * it;
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            F[m] = fmaf(s, r2, F[m]);
            const float a = fmaf(mk, c, -s);
            X[m] = fmaf(a, dx, X[m]);
            Y[m] = fmaf(a, dy, Y[m]);
            Z[m] = fmaf(a, dz, Z[m]);
            mk += kap;
            const float sn = fmaf(s, cth, c * sth);
            const float cn = fmaf(c, cth, -(s * sth));
            s = sn; c = cn;
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int m = 0; m < 32; ++m) acc += F[m] + X[m] + Y[m] + Z[m];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Same work, two half-chunks packed in float2: 10 packed instructions per 2 bins.
__global__ void __launch_bounds__(256) step_packed(float *out, long long *cyc, float cth, float sth,
                                                   float kap, float r2, float dx, float dy, float dz)
{
    float2 F[16], X[16], Y[16], Z[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { F[m] = make_float2(0, 0); X[m] = F[m]; Y[m] = F[m]; Z[m] = F[m]; }
    float2 s = make_float2(threadIdx.x * 1e-3f, threadIdx.x * 2e-3f);
    float2 c = make_float2(1.f - s.x, 1.f - s.y);
    const float2 cth2 = make_float2(cth, cth), sth2 = make_float2(sth, sth), nsth2 = make_float2(-sth, -sth);
    const float2 r22 = make_float2(r2, r2), dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy),
                 dz2 = make_float2(dz, dz), kap2 = make_float2(kap, kap);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; ++it) {
        float2 mk = make_float2(kap * it, kap * (it + 16));
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            F[m] = __ffma2_rn(s, r22, F[m]);
            const float2 a = __ffma2_rn(mk, c, make_float2(-s.x, -s.y));
            X[m] = __ffma2_rn(a, dx2, X[m]);
            Y[m] = __ffma2_rn(a, dy2, Y[m]);
            Z[m] = __ffma2_rn(a, dz2, Z[m]);
            mk = __fadd2_rn(mk, kap2);
            const float2 sn = __ffma2_rn(s, cth2, __fmul2_rn(c, sth2));
            const float2 cn = __ffma2_rn(c, cth2, __fmul2_rn(s, nsth2));
            s = sn; c = cn;
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int m = 0; m < 16; ++m) acc += F[m].x + F[m].y + X[m].x + X[m].y + Y[m].x + Y[m].y + Z[m].x + Z[m].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    int dev = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const int sms = prop.multiProcessorCount;
    float *out;
    long long *cyc, hc[1024];
    cudaMalloc(&out, sizeof(float) * 1024 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    const char *names[] = {"FFMA invariant operands", "FFMA 3 distinct regs", "FFMA2 packed (x2 flops)",
                           "DFMA", "MUFU.SIN (+FMUL)", "FMUL/FADD alternate", "FFMA2 3 distinct",
                           "16 FFMA + 2 MUFU"};
    const double per_iter[] = {16, 16, 8, 8, 32, 16, 8, 20};
    for (int warps_per_sm = 8; warps_per_sm <= 16; warps_per_sm += 8) {
        const int blocks = sms * (warps_per_sm / 8);
        printf("== %d warps/SM (%d blocks of 256) ==\n", warps_per_sm, blocks);
        for (int kind = 0; kind < 8; ++kind) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (kind) {
                case 0: bench<0><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 1: bench<1><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 2: bench<2><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 3: bench<3><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 4: bench<4><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 5: bench<5><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 6: bench<6><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                case 7: bench<7><<<blocks, 256>>>(out, cyc, 1.0001f, 0.5f); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(hc, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
            const double instr = per_iter[kind] * ITERS * 256.0 * (warps_per_sm / 8);
            printf("%-28s %8lld cyc  %7.1f thread-instr/clk/SM\n", names[kind], hc[0], instr / hc[0]);
        }
        for (int rep = 0; rep < 2; ++rep) step_scalar<<<blocks, 256>>>(out, cyc, 0.8f, 0.6f, 0.3f, 9.f, 1.f, 2.f, 2.f);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
        double bins = (ITERS / 8) * 32.0 * 256.0 * (warps_per_sm / 8);
        printf("%-28s %8lld cyc  %7.2f bins/clk/SM (x10 = %.1f instr)\n", "Debye step scalar", hc[0], bins / hc[0], 10 * bins / hc[0]);
        for (int rep = 0; rep < 2; ++rep) step_packed<<<blocks, 256>>>(out, cyc, 0.8f, 0.6f, 0.3f, 9.f, 1.f, 2.f, 2.f);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
        printf("%-28s %8lld cyc  %7.2f bins/clk/SM\n", "Debye step packed (f32x2)", hc[0], bins / hc[0]);
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
