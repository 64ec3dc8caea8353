// iid_small.cuh -- the small float64 stages around the pair sums (sm_100a):
// position staging, F(Q) normalisation, F(Q)->G(r), Rw / chi^2 with the
// chain-rule weights, and the batched grad F(Q) -> grad G(r) contraction.
// References: pyiid/experiments/elasticscatter/kernels/master_kernel.py
// (get_pdf_at_qmin :39-104, get_rw :206-236, get_chi_sq :239-266,
// grad_pdf :276-290, get_grad_rw :293-347, get_grad_chi_sq :350-375).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iid {

// Stage the caller's [n,3] float64 positions into the element-sorted, padded
// SoA layout.  IID_FP32 rounds through float32 first, as wrap_fq does
// (cpu_wrappers/flat_multi_cpu_wrap.py:11-12); the rounded value is then held
// exactly in float64.
// zero_a / zero_b (may be null): accumulators of the passes that follow (S,
// force), cleared here so that the fused sequence needs no memset nodes.
__global__ void prep_kernel(const double *__restrict__ pos,
                            const int *__restrict__ orig, int np,
                            int round_f32, double *__restrict__ x,
                            double *__restrict__ y, double *__restrict__ z,
                            float *__restrict__ valid,
                            double *__restrict__ zero_a = nullptr, int na = 0,
                            double *__restrict__ zero_b = nullptr, int nb = 0)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    for (int e = k; e < na; e += gridDim.x * blockDim.x) zero_a[e] = 0.0;
    for (int e = k; e < nb; e += gridDim.x * blockDim.x) zero_b[e] = 0.0;
    if (k >= np) return;
    const int o = orig[k];
    double px = 0.0, py = 0.0, pz = 0.0;
    if (o >= 0) {
        px = pos[3 * (size_t)o];
        py = pos[3 * (size_t)o + 1];
        pz = pos[3 * (size_t)o + 2];
        if (round_f32) {
            px = (double)(float)px;
            py = (double)(float)py;
            pz = (double)(float)pz;
        }
    }
    x[k] = px;
    y[k] = py;
    z[k] = pz;
    valid[k] = o >= 0 ? 1.f : 0.f;
}

// F[m] = 2 S[m] / na[m]; 0 where na == 0 (nan_to_num of
// flat_multi_cpu_wrap.py:57).  round_f32 reproduces the reference's float32
// result dtype (final.astype(float32) :51).
__global__ void finish_fq_kernel(const double *__restrict__ S,
                                 const double *__restrict__ inv_na, int nq,
                                 int round_f32, double *__restrict__ F)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nq) return;
    double f = 2.0 * S[m] * inv_na[m];
    if (round_f32) f = (double)(float)f;
    F[m] = f;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// G[r] = sum_m T[r][m] F[m]; one warp per r row, fixed summation order.
template <typename TM>
__global__ void gr_kernel(const TM *__restrict__ T,
                          const double *__restrict__ F, int64_t nr, int nq,
                          int qp, double *__restrict__ G)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nr) return;
    const TM *t = T + (size_t)row * qp;
    double acc = 0.0;
    for (int m = lane; m < nq; m += 32) acc = fma((double)t[m], F[m], acc);
    acc = warp_sum_d(acc);
    if (lane == 0) G[row] = acc;
}

// The same with F[m] = 2 S[m] / na[m] formed on the fly (finish_fq_kernel
// folded in; block 0 also stores F): one launch less in the fused sequence,
// identical arithmetic.
__global__ void gr_from_s_kernel(const double *__restrict__ T, const double *__restrict__ S,
                                 const double *__restrict__ inv_na, int64_t nr, int nq, int qp,
                                 double *__restrict__ F, double *__restrict__ G)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0)
        for (int m = threadIdx.x; m < nq; m += blockDim.x) F[m] = 2.0 * S[m] * inv_na[m];
    if (row >= nr) return;
    const double *t = T + (size_t)row * qp;
    double acc = 0.0;
    for (int m = lane; m < nq; m += 32) acc = fma(t[m], 2.0 * S[m] * inv_na[m], acc);
    acc = warp_sum_d(acc);
    if (lane == 0) G[row] = acc;
}

// Block-wide sum of up to three doubles (blockDim.x == 1024).
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c,
                                           double *sm)
{
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    c = warp_sum_d(c);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; sm[64 + w] = c; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    double xa = lane < nw ? sm[lane] : 0.0;
    double xb = lane < nw ? sm[32 + lane] : 0.0;
    double xc = lane < nw ? sm[64 + lane] : 0.0;
    a = warp_sum_d(xa);
    b = warp_sum_d(xb);
    c = warp_sum_d(xc);
}

// Rw / chi^2, scale, and c[r] such that grad[i,w] = sum_r c[r] dG[i,w,r].
// out[4] = {value*conv, scale, value, scale_true}.
// get_rw: scale = (gc.go)/(gc.gc); scale <= 0 -> (1, 1);
//   grad = -rw/(d.d) * sum_r (scale dG + gc grad_a) d,  d = go - scale gc,
//   grad_a = (-2 a (gc.dG) + (go.dG))/(gc.gc) with a the true scale.
// get_chi_sq: scale <= 0 -> scale = 1; grad = -2 sum_r (...) d.
__global__ void potential_kernel(const double *__restrict__ gc,
                                 const double *__restrict__ go, int nr,
                                 int potential, double conv,
                                 double *__restrict__ out,
                                 double *__restrict__ cr,
                                 double *__restrict__ coef)
{
    __shared__ double sm[96];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const double x = gc[r], y = go[r];
        a = fma(x, y, a);
        b = fma(x, x, b);
        c = fma(y, y, c);
    }
    block_sum3(a, b, c, sm);
    const double scale_true = b > 0.0 ? a / b : 0.0;
    const bool pos = scale_true > 0.0;
    const double scale = pos ? scale_true : 1.0;
    double dd = 0.0, gd = 0.0, zz = 0.0;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const double d = go[r] - scale * gc[r];
        dd = fma(d, d, dd);
        gd = fma(gc[r], d, gd);
    }
    block_sum3(dd, gd, zz, sm);
    double value, pref;
    if (potential == 0) {  // Rw
        value = pos ? sqrt(dd / c) : 1.0;
        pref = dd > 0.0 ? -value / dd : 0.0;
    } else {  // chi^2
        value = dd;
        pref = -2.0;
    }
    const double gdb = b > 0.0 ? gd / b : 0.0;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const double d = go[r] - scale * gc[r];
        cr[r] = pref * (scale * d + gdb * (go[r] - 2.0 * scale_true * gc[r]));
    }
    if (threadIdx.x == 0) {
        out[0] = value * conv;
        out[1] = scale;
        out[2] = value;
        out[3] = scale_true;
        if (coef) {
            // c = coef[0] * go - coef[1] * gc  (the chain-rule vector as a
            // combination of the target and the model, see wq_from_q_kernel)
            coef[0] = pref * (scale + gdb);
            coef[1] = pref * (scale * scale + 2.0 * gdb * scale_true);
            out[4] = 0.0;  // fused host path (out has 8 slots): restraint energy, summed later
        }
    }
}

// M = T^T T [qp x qp], once per transform: with it
//   wq = conv T^T c = conv (coef0 T^T go - coef1 M F)
// needs no pass over the R x Q matrix per evaluation.
__global__ void __launch_bounds__(256) ttt_kernel(const double *__restrict__ T, int nr, int nq,
                                                  int qp, double *__restrict__ M)
{
    __shared__ double sa[16][17], sb[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int r0 = 0; r0 < nr; r0 += 16) {
        // tile rows r0..r0+15: sa[r][col-of-block-y], sb[r][col-of-block-x]
        const int r = r0 + ty;
        const int ca = blockIdx.y * 16 + tx, cb = blockIdx.x * 16 + tx;
        sa[ty][tx] = (r < nr && ca < nq) ? T[(size_t)r * qp + ca] : 0.0;
        sb[ty][tx] = (r < nr && cb < nq) ? T[(size_t)r * qp + cb] : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fma(sa[k][ty], sb[k][tx], acc);
        __syncthreads();
    }
    if (m < qp && n < qp) M[(size_t)m * qp + n] = (m < nq && n < nq) ? acc : 0.0;
}

// wq[m] = conv * (coef0 * vgo[m] - coef1 * sum_n M[n][m] F[n]); M is symmetric,
// so the loads are coalesced along m.  Block = 32 bins m x 32 slices of n.
__global__ void __launch_bounds__(1024) wq_from_q_kernel(
    const double *__restrict__ M, const double *__restrict__ F, const double *__restrict__ vgo,
    const double *__restrict__ coef, int nq, int qp, double conv, double *__restrict__ wq)
{
    __shared__ double part[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (m < nq)
        for (int n = ty; n < nq; n += 32) acc = fma(M[(size_t)n * qp + m], F[n], acc);
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && m < nq) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += part[k][tx];
        wq[m] = conv * (coef[0] * vgo[m] - coef[1] * t);
    }
}

// wq[m] = conv * sum_r c[r] T[r][m].  Block = slab of WQ_ROWS rows, thread =
// one m (coalesced along m), four independent partial sums per thread; the
// slab partials meet in one double atomicAdd per (slab, m).  Few, fat slabs:
// same-address double atomics serialise at ~0.2 us each in L2.
constexpr int WQ_ROWS = 128;
__global__ void wq_kernel(const double *__restrict__ T,
                          const double *__restrict__ cr, int nr, int nq, int qp,
                          double conv, double *__restrict__ wq)
{
    const int r0 = blockIdx.x * WQ_ROWS;
    const int r1 = min(nr, r0 + WQ_ROWS);
    for (int mm = threadIdx.x; mm < nq; mm += blockDim.x) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = r0;
        for (; r + 3 < r1; r += 4) {
            a0 = fma(cr[r], T[(size_t)r * qp + mm], a0);
            a1 = fma(cr[r + 1], T[(size_t)(r + 1) * qp + mm], a1);
            a2 = fma(cr[r + 2], T[(size_t)(r + 2) * qp + mm], a2);
            a3 = fma(cr[r + 3], T[(size_t)(r + 3) * qp + mm], a3);
        }
        for (; r < r1; ++r) a0 = fma(cr[r], T[(size_t)r * qp + mm], a0);
        atomicAdd(&wq[mm], conv * ((a0 + a1) + (a2 + a3)));
    }
}

// out[row][r] = sum_m gf[row][m] T[r][m]: float64 tiled contraction, one
// (atom, direction) row of grad F(Q) per output row (master_kernel.grad_pdf
// :276-290 does one 65 536-point FFT per row instead).
constexpr int GP_BM = 64, GP_BN = 64, GP_BK = 16;
template <typename TIN>
__global__ void __launch_bounds__(256) grad_pdf_kernel(
    const TIN *__restrict__ gf, const double *__restrict__ T, int64_t rows,
    int nq, int qp, int nr, double *__restrict__ out)
{
    __shared__ double sa[GP_BK][GP_BM + 1];
    __shared__ double sb[GP_BK][GP_BN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16
    const int64_t row0 = (int64_t)blockIdx.y * GP_BM;
    const int col0 = blockIdx.x * GP_BN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < nq; k0 += GP_BK) {
        for (int e = threadIdx.x; e < GP_BM * GP_BK; e += 256) {
            const int k = e % GP_BK, i = e / GP_BK;
            const int64_t row = row0 + i;
            sa[k][i] = (row < rows && k0 + k < nq)
                           ? (double)gf[row * nq + k0 + k] : 0.0;
        }
        for (int e = threadIdx.x; e < GP_BN * GP_BK; e += 256) {
            const int k = e % GP_BK, j = e / GP_BK;
            const int col = col0 + j;
            sb[k][j] = (col < nr && k0 + k < nq)
                           ? T[(size_t)col * qp + k0 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GP_BK; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sa[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sb[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + ty * 4 + i;
        if (row >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + tx + 16 * j;
            if (col < nr) out[row * nr + col] = acc[i][j];
        }
    }
}

}  // namespace iid
