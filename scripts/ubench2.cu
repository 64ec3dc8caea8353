// Micro-benchmark 2: the Debye bin step with PER-LANE pair constants (as in the
// real kernel, where every lane owns a different atom pair), scalar FFMA versus
// packed FFMA2.  Per-lane operands cannot be promoted to uniform registers, so
// this exposes the register-file operand-fetch behaviour of sm_100.
#include <cstdio>
#include <cuda_runtime.h>
#define REPS 256

__global__ void __launch_bounds__(256) step_scalar(float *out, long long *cyc, const float *in)
{
    float F[32], X[32], Y[32], Z[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) { F[m] = 0; X[m] = 0; Y[m] = 0; Z[m] = 0; }
    const float *q = in + threadIdx.x * 16;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < REPS; ++it) {
        const float cth = q[0] + it, sth = q[1], kap = q[2], r2 = q[3], dx = q[4], dy = q[5], dz = q[6];
        float s = q[7] * it, c = q[8];
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

__global__ void __launch_bounds__(256) step_packed(float *out, long long *cyc, const float *in)
{
    float2 F[16], X[16], Y[16], Z[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { F[m] = make_float2(0, 0); X[m] = F[m]; Y[m] = F[m]; Z[m] = F[m]; }
    const float *q = in + threadIdx.x * 16;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < REPS; ++it) {
        const float cth = q[0] + it, sth = q[1], kap = q[2], r2 = q[3], dx = q[4], dy = q[5], dz = q[6];
        const float2 cth2 = make_float2(cth, cth), sth2 = make_float2(sth, sth), nsth2 = make_float2(-sth, -sth);
        const float2 r22 = make_float2(r2, r2), dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy),
                     dz2 = make_float2(dz, dz), kap2 = make_float2(kap, kap);
        float2 s = make_float2(q[7] * it, q[9] * it), c = make_float2(q[8], q[10]);
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
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float *out, *in;
    long long *cyc, hc[8];
    cudaMalloc(&out, sizeof(float) * 1024 * 1024);
    cudaMalloc(&in, sizeof(float) * 256 * 16);
    float hin[256 * 16];
    for (int i = 0; i < 256 * 16; ++i) hin[i] = 0.3f + 0.001f * (i % 97);
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    const double bins = (double)REPS * 32.0 * 256.0;
    for (int rep = 0; rep < 2; ++rep) step_scalar<<<sms, 256>>>(out, cyc, in);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
    printf("scalar, per-lane constants: %lld cyc, %.2f bins/clk/SM (%.1f%% of 12.8)\n", hc[0], bins / hc[0], 100 * bins / hc[0] / 12.8);
    for (int rep = 0; rep < 2; ++rep) step_packed<<<sms, 256>>>(out, cyc, in);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
    printf("packed, per-lane constants: %lld cyc, %.2f bins/clk/SM (%.1f%% of 12.8)\n", hc[0], bins / hc[0], 100 * bins / hc[0] / 12.8);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
