/*
 * oracle/iid_oracle.c -- TEST INFRASTRUCTURE ONLY (not the product path).
 *
 * CPU restatement, in plain C, of the reference's flattened-pair Debye-sum
 * kernels (pyIID `pyiid/experiments/elasticscatter/kernels/cpu_flat.py`,
 * `kernels/cpu_experimental.py`, `kernels/__init__.py`) and of the chunk
 * workers that drive them (`atomics/cpu_atomics.py`,
 * `cpu_wrappers/flat_serial_cpu_wrap.py`, `cpu_wrappers/flat_multi_cpu_wrap.py`).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may call this library.  The product (pyiid_b200) never does.
 *
 * The arithmetic follows the reference operation by operation: float32
 * storage of every intermediate (d, r, norm, omega, grad_omega, grad), the
 * same loop nest, the same float32 products `sv = qbin * (float)qx`,
 * `sv * r` before sinf/cosf, float64 column sum for F(Q) and a sequential
 * float32 scatter-sum for the gradient.  Like the reference it MATERIALISES
 * the K x Q and K x 3 x Q intermediates for a chunk of pairs; that memory
 * streaming is part of the reference algorithm's cost and is kept so the CPU
 * baseline timing is honest.  Compile with -ffp-contract=off (numba/LLVM does
 * not fuse mul+add by default).
 *
 * Parity pin: oracle/ref_shim.py loads the reference's own numba kernels from
 * /root/reference (build container only) and tests/golden/make_golden.py
 * stores their outputs; tests/test_oracle.py checks this file against them.
 *
 * The `real` type is float for the reference as shipped and double for the
 * "f4 -> f8" variant of the same code used by the FP64
 * parity tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* kernels/__init__.py:15-19  k_to_ij (float64 sqrt), i > j */
static inline void k_to_ij(int64_t k, int64_t *i, int64_t *j)
{
    double id = floor((1.0 + sqrt(1.0 + 8.0 * (double)k)) / 2.0);
    int64_t ii = (int64_t)id;
    /* guard the float64 sqrt at large k (the reference has none; it cannot
     * run there at all because of its int32 casts) */
    while (ii * (ii - 1) / 2 > k) --ii;
    while ((ii + 1) * ii / 2 <= k) ++ii;
    *i = ii;
    *j = k - ii * (ii - 1) / 2;
}

#define DEFINE_ORACLE(real, SUF, SIN, COS, SQRT)                               \
/* cpu_flat.py:14-32 get_d_array */                                            \
static void get_d_array_##SUF(real *d, const real *q, int64_t m,               \
                              int64_t offset)                                  \
{                                                                              \
    for (int64_t k = 0; k < m; ++k) {                                          \
        int64_t i, j;                                                          \
        k_to_ij(k + offset, &i, &j);                                           \
        for (int w = 0; w < 3; ++w) d[3 * k + w] = q[3 * i + w] - q[3 * j + w];\
    }                                                                          \
}                                                                              \
/* cpu_flat.py:35-52 get_r_array: tmp += d*d in working precision, sqrt */     \
static void get_r_array_##SUF(real *r, const real *d, int64_t m)               \
{                                                                              \
    for (int64_t k = 0; k < m; ++k) {                                          \
        real tmp = (real)0;                                                    \
        for (int w = 0; w < 3; ++w) {                                          \
            real p = d[3 * k + w] * d[3 * k + w];                              \
            tmp = tmp + p;                                                     \
        }                                                                      \
        r[k] = SQRT(tmp);                                                      \
    }                                                                          \
}                                                                              \
/* cpu_flat.py:55-74 get_normalization_array: norm[k,q] = f_i(q) f_j(q) */     \
static void get_norm_array_##SUF(real *norm, const real *scat, int64_t m,      \
                                 int64_t Q, int64_t offset)                    \
{                                                                              \
    for (int64_t k = 0; k < m; ++k) {                                          \
        int64_t i, j;                                                          \
        k_to_ij(k + offset, &i, &j);                                           \
        for (int64_t qx = 0; qx < Q; ++qx)                                     \
            norm[k * Q + qx] = scat[i * Q + qx] * scat[j * Q + qx];            \
    }                                                                          \
}                                                                              \
/* cpu_flat.py:77-96 get_omega: Q-outer, k-inner; sin(sv*r)/r */               \
static void get_omega_##SUF(real *omega, const real *r, real qbin, int64_t m,  \
                            int64_t Q)                                         \
{                                                                              \
    for (int64_t qx = 0; qx < Q; ++qx) {                                       \
        real sv = qbin * (real)qx;                                             \
        for (int64_t k = 0; k < m; ++k) {                                      \
            real rk = r[k];                                                    \
            real arg = sv * rk;                                                \
            omega[k * Q + qx] = SIN(arg) / rk;                                 \
        }                                                                      \
    }                                                                          \
}                                                                              \
/* cpu_flat.py:99-105 get_fq: fq[k,q] = norm*omega (Q-outer, k-inner) */       \
static void get_fq_##SUF(real *fq, const real *omega, const real *norm,        \
                         int64_t m, int64_t Q)                                 \
{                                                                              \
    for (int64_t qx = 0; qx < Q; ++qx)                                         \
        for (int64_t k = 0; k < m; ++k)                                        \
            fq[k * Q + qx] = norm[k * Q + qx] * omega[k * Q + qx];             \
}                                                                              \
/* cpu_flat.py:120-131 get_grad_omega */                                       \
static void get_grad_omega_##SUF(real *go, const real *omega, const real *r,   \
                                 const real *d, real qbin, int64_t m,          \
                                 int64_t Q)                                    \
{                                                                              \
    for (int64_t qx = 0; qx < Q; ++qx) {                                       \
        real sv = (real)qx * qbin;                                             \
        for (int64_t k = 0; k < m; ++k) {                                      \
            real rk = r[k];                                                    \
            real arg = sv * rk;                                                \
            real t = sv * COS(arg);                                            \
            real a = t - omega[k * Q + qx];                                    \
            real rr = rk * rk;                                                 \
            a = a / rr;                                                        \
            for (int w = 0; w < 3; ++w)                                        \
                go[(k * 3 + w) * Q + qx] = a * d[3 * k + w];                   \
        }                                                                      \
    }                                                                          \
}                                                                              \
/* cpu_flat.py:134-153 get_grad_fq: grad[k,w,q] = norm[k,q]*grad_omega */      \
static void get_grad_fq_##SUF(real *grad, const real *go, const real *norm,    \
                              int64_t m, int64_t Q)                            \
{                                                                              \
    for (int64_t k = 0; k < m; ++k)                                            \
        for (int w = 0; w < 3; ++w)                                            \
            for (int64_t qx = 0; qx < Q; ++qx)                                 \
                grad[(k * 3 + w) * Q + qx] =                                   \
                    norm[k * Q + qx] * go[(k * 3 + w) * Q + qx];               \
}                                                                              \
/* cpu_experimental.py:8-15 experimental_sum_grad_cpu */                       \
static void sum_grad_##SUF(real *new_grad, const real *grad, int64_t m,        \
                           int64_t Q, int64_t k_cov)                           \
{                                                                              \
    for (int64_t k = 0; k < m; ++k) {                                          \
        int64_t i, j;                                                          \
        k_to_ij(k + k_cov, &i, &j);                                            \
        for (int64_t qx = 0; qx < Q; ++qx)                                     \
            for (int w = 0; w < 3; ++w) {                                      \
                real g = grad[(k * 3 + w) * Q + qx];                           \
                new_grad[(i * 3 + w) * Q + qx] -= g;                           \
                new_grad[(j * 3 + w) * Q + qx] += g;                           \
            }                                                                  \
    }                                                                          \
}                                                                              \
/* atomics/cpu_atomics.py:60-78 atomic_fq: one chunk [k_cov, k_cov+m) of the   \
 * pair list -> float64 column sums (added into sum64[Q]) */                   \
static int atomic_fq_##SUF(const real *q, const real *scat, int64_t Q,         \
                           real qbin, int64_t m, int64_t k_cov, double *sum64) \
{                                                                              \
    real *d = (real *)malloc(sizeof(real) * 3 * m);                            \
    real *r = (real *)malloc(sizeof(real) * m);                                \
    real *norm = (real *)malloc(sizeof(real) * m * Q);                         \
    real *omega = (real *)malloc(sizeof(real) * m * Q);                        \
    real *fq = (real *)malloc(sizeof(real) * m * Q);                           \
    if (!d || !r || !norm || !omega || !fq) {                                  \
        free(d); free(r); free(norm); free(omega); free(fq);                   \
        return -1;                                                             \
    }                                                                          \
    get_d_array_##SUF(d, q, m, k_cov);                                         \
    get_r_array_##SUF(r, d, m);                                                \
    get_norm_array_##SUF(norm, scat, m, Q, k_cov);                             \
    get_omega_##SUF(omega, r, qbin, m, Q);                                     \
    get_fq_##SUF(fq, omega, norm, m, Q);                                       \
    /* fq.sum(axis=0, dtype=float64): row-by-row accumulation */               \
    for (int64_t k = 0; k < m; ++k)                                            \
        for (int64_t qx = 0; qx < Q; ++qx) sum64[qx] += (double)fq[k * Q + qx];\
    free(d); free(r); free(norm); free(omega); free(fq);                       \
    return 0;                                                                  \
}                                                                              \
/* atomics/cpu_atomics.py:81-102 atomic_grad_fq: one chunk, scatter-summed     \
 * into rtn[N,3,Q] (working precision, sequential in k like the reference) */  \
static int atomic_grad_fq_##SUF(const real *q, const real *scat, int64_t Q,    \
                                real qbin, int64_t m, int64_t k_cov,           \
                                real *rtn)                                     \
{                                                                              \
    real *d = (real *)malloc(sizeof(real) * 3 * m);                            \
    real *r = (real *)malloc(sizeof(real) * m);                                \
    real *norm = (real *)malloc(sizeof(real) * m * Q);                         \
    real *omega = (real *)malloc(sizeof(real) * m * Q);                        \
    real *go = (real *)malloc(sizeof(real) * 3 * m * Q);                       \
    real *grad = (real *)malloc(sizeof(real) * 3 * m * Q);                     \
    if (!d || !r || !norm || !omega || !go || !grad) {                         \
        free(d); free(r); free(norm); free(omega); free(go); free(grad);       \
        return -1;                                                             \
    }                                                                          \
    get_d_array_##SUF(d, q, m, k_cov);                                         \
    get_r_array_##SUF(r, d, m);                                                \
    get_norm_array_##SUF(norm, scat, m, Q, k_cov);                             \
    get_omega_##SUF(omega, r, qbin, m, Q);                                     \
    get_grad_omega_##SUF(go, omega, r, d, qbin, m, Q);                         \
    get_grad_fq_##SUF(grad, go, norm, m, Q);                                   \
    sum_grad_##SUF(rtn, grad, m, Q, k_cov);                                    \
    free(d); free(r); free(norm); free(omega); free(go); free(grad);           \
    return 0;                                                                  \
}                                                                              \
/* Pair-sum part of wrap_fq (flat_serial_cpu_wrap.py:13-69 /                   \
 * flat_multi_cpu_wrap.py:22-60): S[q] = sum_k norm*omega in float64 over the  \
 * pair range [k_begin, k_end), processed in chunks of `chunk` pairs.          \
 * nthreads > 1 distributes chunks over OpenMP threads the way the reference   \
 * distributes them over a multiprocessing.Pool; per-chunk float64 sums are    \
 * then added in chunk order (np.sum(ans, axis=0, dtype=float64)). */          \
int oracle_fq_pairsum_##SUF(const real *q, const real *scat, int64_t n,        \
                            int64_t Q, real qbin, int64_t k_begin,             \
                            int64_t k_end, int64_t chunk, int nthreads,        \
                            double *sum64)                                     \
{                                                                              \
    (void)n;                                                                   \
    if (chunk <= 0) chunk = 1 << 14;                                           \
    int64_t nchunk = (k_end - k_begin + chunk - 1) / chunk;                    \
    for (int64_t qx = 0; qx < Q; ++qx) sum64[qx] = 0.0;                        \
    if (nchunk <= 0) return 0;                                                 \
    double *part = (double *)calloc((size_t)(nchunk * Q), sizeof(double));     \
    if (!part) return -1;                                                      \
    int err = 0;                                                               \
    if (nthreads < 1) nthreads = 1;                                            \
    _Pragma("omp parallel for schedule(dynamic, 1) num_threads(nthreads)")     \
    for (int64_t c = 0; c < nchunk; ++c) {                                     \
        int64_t k0 = k_begin + c * chunk;                                      \
        int64_t m = (k0 + chunk <= k_end) ? chunk : (k_end - k0);              \
        if (atomic_fq_##SUF(q, scat, Q, qbin, m, k0, part + c * Q)) err = -1;  \
    }                                                                          \
    for (int64_t c = 0; c < nchunk; ++c)                                       \
        for (int64_t qx = 0; qx < Q; ++qx) sum64[qx] += part[c * Q + qx];      \
    free(part);                                                                \
    return err;                                                                \
}                                                                              \
/* Pair-sum part of wrap_fq_grad (flat_serial_cpu_wrap.py:72-133,              \
 * flat_multi_cpu_wrap.py:63-102).  nthreads == 1: chunks are processed in     \
 * pair order into ONE accumulator array, which is bit-identical to the        \
 * single-chunk flat-serial path.  nthreads > 1: one accumulator per thread,   \
 * summed at the end (the Pool variant, np.sum(ans, axis=0)). */               \
int oracle_grad_pairsum_ws_##SUF(const real *q, const real *scat, int64_t n,   \
                              int64_t Q, real qbin, int64_t k_begin,           \
                              int64_t k_end, int64_t chunk, int nthreads,      \
                              real *rtn, real *ws)                             \
{                                                                              \
    if (chunk <= 0) chunk = 1 << 12;                                           \
    int64_t nchunk = (k_end - k_begin + chunk - 1) / chunk;                    \
    size_t sz = (size_t)n * 3 * (size_t)Q;                                     \
    memset(rtn, 0, sizeof(real) * sz);                                         \
    if (nchunk <= 0) return 0;                                                 \
    if (nthreads <= 1) {                                                       \
        for (int64_t c = 0; c < nchunk; ++c) {                                 \
            int64_t k0 = k_begin + c * chunk;                                  \
            int64_t m = (k0 + chunk <= k_end) ? chunk : (k_end - k0);          \
            if (atomic_grad_fq_##SUF(q, scat, Q, qbin, m, k0, rtn)) return -1; \
        }                                                                      \
        return 0;                                                              \
    }                                                                          \
    int err = 0;                                                               \
    real *part = ws;                                                           \
    if (part) memset(part, 0, sizeof(real) * sz * (size_t)nthreads);           \
    else part = (real *)calloc(sz * (size_t)nthreads, sizeof(real));           \
    if (!part) return -1;                                                      \
    _Pragma("omp parallel num_threads(nthreads)")                              \
    {                                                                          \
        int t = 0;                                                             \
        OMP_TID(t);                                                            \
        _Pragma("omp for schedule(dynamic, 1)")                                \
        for (int64_t c = 0; c < nchunk; ++c) {                                 \
            int64_t k0 = k_begin + c * chunk;                                  \
            int64_t m = (k0 + chunk <= k_end) ? chunk : (k_end - k0);          \
            if (atomic_grad_fq_##SUF(q, scat, Q, qbin, m, k0,                  \
                                     part + (size_t)t * sz))                   \
                err = -1;                                                      \
        }                                                                      \
    }                                                                          \
    for (int t = 0; t < nthreads; ++t)                                         \
        for (size_t e = 0; e < sz; ++e) rtn[e] += part[(size_t)t * sz + e];    \
    if (!ws) free(part);                                                       \
    return err;                                                                \
}                                                                              \
/* `ws` (may be NULL) of the _ws variant: caller-provided nthreads * n*3*Q      \
 * accumulators, so that repeated timed calls (bench.py --impl reference) do   \
 * not re-allocate and page-fault 150 MB per thread on every call. */          \
int oracle_grad_pairsum_##SUF(const real *q, const real *scat, int64_t n,      \
                              int64_t Q, real qbin, int64_t k_begin,           \
                              int64_t k_end, int64_t chunk, int nthreads,      \
                              real *rtn)                                       \
{                                                                              \
    return oracle_grad_pairsum_ws_##SUF(q, scat, n, Q, qbin, k_begin, k_end,   \
                                        chunk, nthreads, rtn, NULL);           \
}                                                                              \
/* Expose the per-pair intermediates for the kernel-internals tests            \
 * (reference tests/test_scatter_internals.py:39-94). */                       \
int oracle_pair_internals_##SUF(const real *q, const real *scat, int64_t n,    \
                                int64_t Q, real qbin, real *d, real *r,        \
                                real *norm, real *omega)                       \
{                                                                              \
    int64_t K = n * (n - 1) / 2;                                               \
    get_d_array_##SUF(d, q, K, 0);                                             \
    get_r_array_##SUF(r, d, K);                                                \
    get_norm_array_##SUF(norm, scat, K, Q, 0);                                 \
    get_omega_##SUF(omega, r, qbin, K, Q);                                     \
    return 0;                                                                  \
}

#ifdef _OPENMP
#define OMP_TID(t) (t) = omp_get_thread_num()
#else
#define OMP_TID(t) (t) = 0
#endif

DEFINE_ORACLE(float, f32, sinf, cosf, sqrtf)
DEFINE_ORACLE(double, f64, sin, cos, sqrt)

/* Normaliser of wrap_fq / wrap_fq_grad: na[q] = mean_k(norm[k,q]) * n.
 * mode 0: float64 mean (the reference's commented alternative,
 *         flat_multi_cpu_wrap.py:55) -- the oracle's normaliser.
 * mode 1: the reference AS SHIPPED, np.mean(norm, axis=0, dtype=float32) *
 *         float32(n): a naive sequential float32 accumulation over the K rows
 *         (flat_multi_cpu_wrap.py:54); reported as the "as-is" deviation. */
int oracle_normaliser_f32(const float *scat, int64_t n, int64_t Q, int mode,
                          double *na)
{
    int64_t K = n * (n - 1) / 2;
    if (mode == 0) {
        for (int64_t qx = 0; qx < Q; ++qx) na[qx] = 0.0;
        for (int64_t i = 1; i < n; ++i)
            for (int64_t j = 0; j < i; ++j)
                for (int64_t qx = 0; qx < Q; ++qx) {
                    float nm = scat[i * Q + qx] * scat[j * Q + qx];
                    na[qx] += (double)nm;
                }
        for (int64_t qx = 0; qx < Q; ++qx)
            na[qx] = K > 0 ? na[qx] / (double)K * (double)n : 0.0;
    } else {
        float *acc = (float *)calloc((size_t)Q, sizeof(float));
        if (!acc) return -1;
        for (int64_t i = 1; i < n; ++i)
            for (int64_t j = 0; j < i; ++j)
                for (int64_t qx = 0; qx < Q; ++qx) {
                    float nm = scat[i * Q + qx] * scat[j * Q + qx];
                    acc[qx] = acc[qx] + nm;
                }
        for (int64_t qx = 0; qx < Q; ++qx) {
            float mean = K > 0 ? acc[qx] / (float)K : 0.0f;
            na[qx] = (double)(mean * (float)n);
        }
        free(acc);
    }
    return 0;
}

int oracle_normaliser_f64(const double *scat, int64_t n, int64_t Q, int mode,
                          double *na)
{
    (void)mode;
    int64_t K = n * (n - 1) / 2;
    for (int64_t qx = 0; qx < Q; ++qx) na[qx] = 0.0;
    for (int64_t i = 1; i < n; ++i)
        for (int64_t j = 0; j < i; ++j)
            for (int64_t qx = 0; qx < Q; ++qx)
                na[qx] += scat[i * Q + qx] * scat[j * Q + qx];
    for (int64_t qx = 0; qx < Q; ++qx)
        na[qx] = K > 0 ? na[qx] / (double)K * (double)n : 0.0;
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
