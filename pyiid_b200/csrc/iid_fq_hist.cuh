// iid_fq_hist.cuh -- the F(Q) pair sum of a LARGE structure through a radial
// pair histogram (FP32 mode, sm_100a).
//
// cpu_flat.get_omega / get_fq (:77-114) sum f_i f_j sin(Q r_ij) / r_ij over the
// pairs for every Q bin: O(N^2 Q).  For one pair of element types the sum depends
// on the pairs only through their distances, and g_Q(r) = sin(Q r) / r is band
// limited by Q_max: on a uniform r grid with Q_max h = 1/3 a 12-point Lagrange
// interpolation reproduces it to 4e-10 of its amplitude (iid_stencil.cuh).  So
//
//   S_ab[Q] = sum_pairs g_Q(r) = sum_k C_ab[k] g_Q(r_k),
//   C_ab[k] = sum_pairs L_{k - k0(r)}(u(r))          (the interpolation's adjoint:
//                                                     each pair spreads its unit
//                                                     weight over 12 grid nodes)
//
//   1. fq_hist_grid_kernel   bounding box -> stencil, grid step, nodes in use
//   2. fq_hist_kernel        O(N^2): float64 distance, P float32 weights, 2 P
//                            shared-memory integer atomics per pair
//   3. fq_hist_transform_kernel  O(K Q): S_ab from C_ab in float64, per-block
//                            partial sums for reduce_spart_kernel (fixed order)
//
// instead of ~830 instructions per pair over 330 bins.  The pass is bound by the
// shared-memory atomic unit (one warp-wide atomic per ~5.7 cycles per SM), i.e. by
// the NUMBER of stencil points: the pass takes the shortest stencil of
// iid_stencil.cuh whose grid fits the shared-memory histogram -- 6 points up to
// 75 A at Q_max = 25, 8 points up to 180 A, the 12 points of the coarse grid up to
// 382 A; all three keep the 4e-10 bound.  The histogram is FIXED POINT (units of
// 2^-28): integer addition commutes, so F(Q) is bit-reproducible whatever the
// order of the atomics; the only native shared-memory atomic add is
// 32 bits wide, so a node is two words, high (units of 2^-12) and low (16 bits),
// with the carries folded every 49 152 pairs and the block's histogram flushed
// to the 64-bit global one every 393 216 pairs.  Pair distances carry the
// float32 rounding of the positions (FP32 mode) and nothing else: this pass is
// more accurate than the float32 recurrences it replaces.  A structure whose
// diameter (twice the radius about its bounding-box centre) does not fit the
// shared-memory histogram (382 A at Q_max = 25) keeps the direct kernel (the
// gate word below).
#pragma once
#include "iid_debye.cuh"
#include "iid_stencil.cuh"

namespace iid {

constexpr int FH_THREADS = 1024;  // 32 warps per SM: the loads and weights of one warp hide behind the others' atomics
constexpr int FH_CAP = 28672;  // nodes of the shared-memory histogram (2 x 4 B each: 224 KB)
constexpr unsigned FH_FOLD_PAIRS = 49152;    // low word: 65 535 per add, 2^32 / 65 535 adds
constexpr unsigned FH_FLUSH_PAIRS = 393216;  // high word: <= 4 916 per add (+ carries)

struct HistParams {
    const double *x, *y, *z;  // [np] sorted/padded positions
    const float *valid;
    const int *tile_type;
    const WorkItem *items;  // triangle list
    const int *order;       // item indices sorted by element pair
    int n_items, rank, world;
    const double *info;  // h, 1/h, nodes in use, gate (1 = histogram pass, 0 = direct kernel)
    unsigned long long *C;  // [pairs][stride] fixed point, units of 2^-28
    int stride;
};

// One pair: its unit weight over the nodes k - P/2 + 1 .. k + P/2 of a two-word
// fixed-point histogram in shared memory (units of 2^-28; high word 2^-12, low
// word 16 bits) that stores P/2 nodes before r = 0.  u in [0, 1) past node k.
template <int P>
__device__ __forceinline__ void fq_hist_spread(int *hi_s, unsigned *lo_s, int k, float u)
{
    constexpr int LEFT = P / 2 - 1, PAD = P / 2;
    // Lagrange weights (barycentric form by prefix / suffix products; float32: 1e-7
    // of a unit weight)
    float d[P], pre[P];
#pragma unroll
    for (int i = 0; i < P; ++i) d[i] = u - (float)(i - LEFT);
    pre[0] = 1.f;
#pragma unroll
    for (int i = 1; i < P; ++i) pre[i] = pre[i - 1] * d[i - 1];
    float suf = 1.f;
    int *hp = hi_s + (k - LEFT + PAD);
    unsigned *lp = lo_s + (k - LEFT + PAD);
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
        const float w = (float)lagrange_bary<P>(i) * 268435456.f * pre[i] * suf;  // 2^28
        suf *= d[i];
        const int q = __float2int_rn(w);
        atomicAdd(hp + i, q >> 16);             // floor: q = hi 2^16 + lo
        atomicAdd(lp + i, (unsigned)(q & 0xffff));
    }
}

// info[0] = h, info[1] = 1/h, info[2] = K (nodes r = 0 .. (K-1) h), info[3] = gate,
// info[4] = stencil points: the shortest stencil / finest grid of iid_stencil.cuh the
// structure fits (6, 8 or 12 points)
__global__ void __launch_bounds__(1024) fq_hist_grid_kernel(const double *__restrict__ x,
                                                            const double *__restrict__ y,
                                                            const double *__restrict__ z,
                                                            const float *__restrict__ valid,
                                                            int np, double h12, double *info)
{
    __shared__ double smin[3][32], smax[3][32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = threadIdx.x; k < np; k += blockDim.x) {
        if (valid[k] == 0.f) continue;
        const double v[3] = {x[k], y[k], z[k]};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], v[a]); hi[a] = fmax(hi[a], v[a]); }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { smin[a][w] = lo[a]; smax[a][w] = hi[a]; }
    }
    __syncthreads();
    // every pair distance is <= twice the largest distance from the box centre
    // (a sphere's bounding-box diagonal is 1.7 diameters)
    __shared__ double ctr[3], rad[32];
    if (threadIdx.x < 3) {
        const int a = threadIdx.x, nw = blockDim.x >> 5;
        double l = 1e300, u = -1e300;
        for (int k = 0; k < nw; ++k) { l = fmin(l, smin[a][k]); u = fmax(u, smax[a][k]); }
        ctr[a] = u > l ? 0.5 * (l + u) : 0.0;
    }
    __syncthreads();
    double d2 = 0.0;
    for (int k = threadIdx.x; k < np; k += blockDim.x) {
        if (valid[k] == 0.f) continue;
        const double dx = x[k] - ctr[0], dy = y[k] - ctr[1], dz = z[k] - ctr[2];
        d2 = fmax(d2, fma(dx, dx, fma(dy, dy, dz * dz)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if (lane == 0) rad[w] = d2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, rad[k]);
        const double rmax = 2.0 * sqrt(m) * 1.0000001 + 1e-9;  // longest possible pair distance
        // (a function of the positions alone: the same structure takes the same
        // stencil, whatever was evaluated before)
        const double hs[3] = {h12 * (FH_QH_FINEST / FT_QH), h12 * (FH_QH_FINE / FT_QH), h12};
        const int pts[3] = {FH_PTS_FINEST, FH_PTS_FINE, FT_PTS};
        int t = 0;  // the finest grid the structure fits (the last one decides the gate)
        while (t < 2 && ceil(rmax / hs[t]) + (double)(pts[t] + 2) > (double)(FH_CAP - pts[t])) ++t;
        const double h = hs[t], K = ceil(rmax / h) + (double)(pts[t] + 2);
        const double cap = (double)(FH_CAP - pts[t]);
        info[0] = h;
        info[1] = 1.0 / h;
        info[2] = fmin(K, cap);
        info[3] = K <= cap ? 1.0 : 0.0;
        info[4] = (double)pts[t];
    }
}

template <int P>
__device__ __forceinline__ void fq_hist_body(const HistParams &p, unsigned char *smem_raw)
{
    const int K = (int)p.info[2], Kp = K + P;  // P/2 nodes before r = 0, P/2 past the last
    const double inv_h = p.info[1];
    int *hi_s = reinterpret_cast<int *>(smem_raw);      // [Kp] units of 2^-12
    unsigned *lo_s = reinterpret_cast<unsigned *>(hi_s + FH_CAP);  // [Kp] units of 2^-28, < 2^16 per add
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < Kp; e += blockDim.x) { hi_s[e] = 0; lo_s[e] = 0u; }
    __syncthreads();

    auto fold = [&]() {  // carries of the low words into the high words
        __syncthreads();
        for (int e = threadIdx.x; e < Kp; e += blockDim.x) {
            const unsigned l = lo_s[e];
            hi_s[e] += (int)(l >> 16);
            lo_s[e] = l & 0xffffu;
        }
        __syncthreads();
    };
    auto flush = [&](int pair) {  // the block's histogram into the global one
        __syncthreads();
        unsigned long long *C = p.C + (size_t)pair * p.stride;
        for (int e = threadIdx.x; e < Kp; e += blockDim.x) {
            const long long v = (long long)hi_s[e] * 65536ll + (long long)lo_s[e];
            if (v != 0) atomicAdd(C + e, (unsigned long long)v);
            hi_s[e] = 0;
            lo_s[e] = 0u;
        }
        __syncthreads();
    };

    int cur_pair = -1;
    unsigned since_fold = 0, since_flush = 0;  // pairs (upper bounds), block-uniform
    for (int l = blockIdx.x; l < p.n_items; l += gridDim.x) {
        const int gidx = p.order[l];
        if (gidx % p.world != p.rank) continue;
        const WorkItem it = p.items[gidx];
        const int ta = p.tile_type[it.itile], tb = it.info & 0xffff;
        const int pair = ta >= tb ? ta * (ta + 1) / 2 + tb : tb * (tb + 1) / 2 + ta;
        if (pair != cur_pair) {
            if (cur_pair >= 0) flush(cur_pair);
            cur_pair = pair;
            since_fold = since_flush = 0;
        }
        const bool diag = (it.info & ITEM_DIAG) != 0;
        const int gi = it.itile * TILE_I + lane;
        const double xi = p.x[gi], yi = p.y[gi], zi = p.z[gi];
        const bool vi = p.valid[gi] != 0.f;
        // the j atom of the NEXT trip is loaded before this trip's pair is spread: the
        // load's latency (ncu: long-scoreboard stalls at the loop head, 29 % of the
        // samples once the atomics were down to 16 per pair) hides behind the atomics
        double xj = 0.0, yj = 0.0, zj = 0.0;
        bool vj = false;
        if (it.jbegin + warp < it.jend) {
            const int g0 = it.jbegin + warp;
            xj = p.x[g0]; yj = p.y[g0]; zj = p.z[g0];
            vj = p.valid[g0] != 0.f;
        }
        for (int j0 = it.jbegin; j0 < it.jend; j0 += nw) {
            const int gj = j0 + warp, gn = gj + nw;
            double xn = 0.0, yn = 0.0, zn = 0.0;
            bool vn = false;
            if (gn < it.jend) {
                xn = p.x[gn]; yn = p.y[gn]; zn = p.z[gn];
                vn = p.valid[gn] != 0.f;
            }
            // a diagonal item holds both orders of its pairs: count i > j
            if (gj < it.jend && vi && vj && (!diag || gj < gi)) {
                const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
                const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                if (r2 > 0.0) {
                    double y = (double)rsqrtf((float)r2);
                    y = y * fma(-0.5 * r2, y * y, 1.5);
                    y = y * fma(-0.5 * r2, y * y, 1.5);  // second Newton step: 1e-15
                    const double tpos = r2 * y * inv_h;
                    const int k = (int)tpos;
                    const float u = (float)(tpos - (double)k);
                    fq_hist_spread<P>(hi_s, lo_s, k, u);
                }
            }
            xj = xn; yj = yn; zj = zn; vj = vn;
            since_fold += (unsigned)blockDim.x;
            since_flush += (unsigned)blockDim.x;
            if (since_fold >= FH_FOLD_PAIRS) {
                fold();
                since_fold = 0;
            }
            if (since_flush >= FH_FLUSH_PAIRS) {
                flush(cur_pair);
                since_flush = since_fold = 0;
            }
        }
    }
    if (cur_pair >= 0) flush(cur_pair);
}

__global__ void __launch_bounds__(FH_THREADS, 1) fq_hist_kernel(const HistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.info[3] == 0.0) return;  // the structure does not fit: the direct kernel runs
    const int pts = (int)p.info[4];  // block-uniform
    if (pts == FH_PTS_FINEST) fq_hist_body<FH_PTS_FINEST>(p, smem_raw);
    else if (pts == FH_PTS_FINE) fq_hist_body<FH_PTS_FINE>(p, smem_raw);
    else fq_hist_body<FT_PTS>(p, smem_raw);
}

// S_ab[m] from C_ab: block (chunk, pair) sums its FHT_E nodes for every Q bin in
// float64 (thread = Q bin, three-term recurrence in the node index from exact
// seeds) and stores a partial sum row for reduce_spart_kernel; the nodes it has
// read are cleared for the next evaluation.
constexpr int FHT_E = 256;
__global__ void __launch_bounds__(384) fq_hist_transform_kernel(
    unsigned long long *__restrict__ C, int stride, const double *__restrict__ info,
    const float *__restrict__ ftab, int ntypes, int nq, int qp, double qbin,
    double *__restrict__ Spart)
{
    __shared__ double cs[FHT_E];  // C[e] / |r_e| (the r = 0 node: C[e])
    const int pair = blockIdx.y;
    // pair = a (a + 1) / 2 + b, a >= b
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= pair) ++a;
    const int b = pair - a * (a + 1) / 2;
    double *out = Spart + ((size_t)pair * gridDim.x + blockIdx.x) * qp;
    const int m = threadIdx.x;
    const bool on = info[3] != 0.0;
    const int pad = (int)info[4] / 2;  // nodes stored before r = 0
    const int Kp = (int)info[2] + 2 * pad;
    const int e0 = blockIdx.x * FHT_E;
    if (!on || e0 >= Kp) {  // block-uniform
        if (m < qp) out[m] = 0.0;
        return;
    }
    const double h = info[0];
    const int ne = min(FHT_E, Kp - e0);
    bool any = false;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        unsigned long long *c = C + (size_t)pair * stride + e0 + e;
        const long long v = (long long)*c;
        *c = 0ull;
        const double r = fabs((double)(e0 + e - pad)) * h;  // the spread is even in r
        cs[e] = (double)v * (1.0 / 268435456.0) * (r > 0.0 ? 1.0 / r : 1.0);
        any |= v != 0;
    }
    const bool work = __syncthreads_or(any) != 0;
    if (m >= qp) return;
    double acc = 0.0;
    if (work && m < nq) {
        const double Q = qbin * (double)m;
        // sin(Q r_e), r_e = (e0 + e - PAD) h: angle step Q h per node
        const double turn = Q * h * 0.15915494309189533577;
        double sth, cth, s, c;
        sincospi(2.0 * (turn - rint(turn)), &sth, &cth);
        const double t0 = turn * (double)(e0 - pad);
        sincospi(2.0 * (t0 - rint(t0)), &s, &c);
        double sp = fma(s, cth, -(c * sth));  // node e0 - 1
        const double tc = cth + cth;
        const int ezero = pad - e0;  // index of the r = 0 node in this chunk (if any)
        for (int e = 0; e < ne; ++e) {
            // |r|: sin(Q |r|) = sign(r) sin(Q r)
            const double sv = (e < ezero) ? -s : s;
            acc = fma(cs[e], e == ezero ? Q : sv, acc);
            const double sn = fma(tc, s, -sp);
            sp = s;
            s = sn;
        }
        acc *= (double)ftab[(size_t)a * qp + m] * (double)ftab[(size_t)b * qp + m];
    }
    out[m] = acc;
}

}  // namespace iid
