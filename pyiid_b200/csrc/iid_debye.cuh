// iid_debye.cuh -- the pair-tiled Debye-sum kernel (sm_100a).
//
// One kernel template covers the three pair sums of the hot path:
//   MODE_FQ    S[m]       = sum_{pairs} f_i f_j sin(Q_m r)/r            (triangle)
//   MODE_GRAD  G[i,w,m]   = sum_j f_i f_j a_ij(m) (q_j - q_i)_w / na    (square, + S)
//   MODE_FORCE force[i,w] = sum_m wq[m] G[i,w,m] without forming G      (triangle)
// with Q_m = m*qbin and a = (Q cos(Q r) - sin(Q r)/r)/r^2.  They replace the
// reference's chain of element-wise kernels that materialise K x Q and
// K x 3 x Q arrays (pyiid/experiments/elasticscatter/kernels/gpu_flat.py:11-299
// driven by atomics/gpu_atomics.py:89-280; CPU twins kernels/cpu_flat.py:14-194,
// kernels/cpu_experimental.py:8-15).  Nothing O(pairs) is ever stored here.
//
// Mapping.  Atoms are sorted by element and every element run is padded to a
// multiple of 32 (ghost atoms carry valid = 0), so a 32-atom i-tile and a
// j-slab each have ONE element type and f_i f_j / na factors out of the pair
// loop.  A work item is (i-tile, j-slab).  A block takes one item; warp w of
// the block owns Q-chunk [m0, m0+C) and lane l owns atom i = 32*itile + l, so
// a thread keeps the C (x4 for the gradient) accumulators of its (atom, chunk)
// in registers for the whole item.  Positions of the j-slab are staged through
// shared memory in tiles of TJ atoms and read by broadcast.
//
// Transcendentals.  The Q grid is uniform, so sin/cos(m*theta), theta = qbin*r,
// are advanced by a complex rotation (2 FMUL + 2 FFMA per bin for both) on the
// FP32 pipe instead of 2 MUFU per bin on the 16-lane SFU.  The state is
// pre-scaled by 1/r^3 so that a_ij(m) = m*kappa*c - s (kappa = qbin*r) costs
// one FFMA.  Per (pair, chunk) the distance, the phase in turns and its
// reduction are computed in FLOAT64 (float32-rounded positions are exact in
// float64, so r and the seed phase m0*theta mod 2pi carry no float32
// cancellation error even at Q r ~ 3000 rad); the rotation constants come from
// a float32 minimax polynomial on the quadrant-reduced step angle and the chunk
// seed from MUFU.SIN/COS on the reduced phase.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iid {

constexpr int TILE_I = 32;  // atoms per i-tile: one lane each
constexpr int TJ = 128;     // j atoms staged in shared memory per tile

struct WorkItem {
    int itile;   // atoms [32*itile, 32*itile + 32) in sorted/padded order
    int jbegin;  // first j (sorted/padded index)
    int jend;    // one past the last j
    int info;    // bits 0..15 element type of the j-slab, bit 16 diagonal item
};
constexpr int ITEM_DIAG = 1 << 16;

enum Mode { MODE_FQ = 0, MODE_GRAD = 1, MODE_FORCE = 2 };

struct DebyeParams {
    const double *x, *y, *z;  // [np] sorted/padded positions
    const float *valid;       // [np] 1 = atom, 0 = ghost
    const int *orig;          // [np] caller's atom index, -1 for ghosts
    const int *tile_type;     // [np/32] element type of each i-tile
    const WorkItem *items;
    int item_begin, item_stride;
    const void *ftab;    // [n_types][qp] form factors, kernel precision
    const void *inv_na;  // [qp] 1/na (0 where na == 0), kernel precision
    const double *wq;    // [qp] chain-rule weights (MODE_FORCE)
    int nq, qp;
    double qbin, qbin_turns;  // qbin and qbin/(2 pi)
    void *G;        // [n][3][nq] kernel precision (MODE_GRAD)
    double *S;      // [qp] (MODE_FQ, MODE_GRAD; may be null in MODE_GRAD)
    double *force;  // [n][3] (MODE_FORCE)
};

// sin(2 pi f), cos(2 pi f) for |f| <= 1/8 turn: float32 minimax polynomials on
// [-pi/4, pi/4] (Cephes sinf/cosf coefficient sets), ~1 ulp.
__device__ __forceinline__ void sincos_eighth(float f, float &s, float &c)
{
    const float x = 6.28318530717958647692f * f;
    const float z = x * x;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    s = fmaf(ps * z, x, x);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-(pair, chunk) set-up: everything the bin loop needs.
template <typename T>
struct PairState {
    T s, c;      // sin, cos of m*theta scaled by 1/r^3 (0 for masked pairs)
    T sth, cth;  // sin, cos of theta = qbin*r
    T kap;       // qbin*r
    T r2;        // r^2 (s*r2 = sin/r)
    T dx, dy, dz;  // q_j - q_i
};

__device__ __forceinline__ void pair_setup(PairState<float> &o, double dxd,
                                           double dyd, double dzd, bool keep,
                                           double qbin, double qbin_turns,
                                           int m0)
{
    const double r2 = fma(dxd, dxd, fma(dyd, dyd, dzd * dzd));
    const float r2f = (float)r2;
    // 1/r: MUFU.RSQ seed, one Newton step in float64
    double y = (double)rsqrtf(r2f);
    y = y * fma(-0.5 * r2, y * y, 1.5);
    if (!(keep && r2f > 0.f)) y = 0.0;  // self pair, ghost atom or r == 0
    const double r = r2 * y;
    const double u = r * qbin_turns;  // turns per Q bin
    // rotation constants: quadrant reduction in float64, polynomial in float32
    const double q4 = rint(u * 4.0);
    const float fr = (float)fma(-0.25, q4, u);
    const int quad = ((int)q4) & 3;
    float sf, cf;
    sincos_eighth(fr, sf, cf);
    float sth = (quad & 1) ? cf : sf;
    float cth = (quad & 1) ? sf : cf;
    if (quad == 1 || quad == 2) cth = -cth;
    if (quad >= 2) sth = -sth;
    // chunk seed: phase of bin m0 reduced exactly in float64, then MUFU
    double ph = u * (double)m0;
    ph -= rint(ph);
    const float pf = 6.28318530717958647692f * (float)ph;
    const float invr = (float)y;
    const float b3 = invr * invr * invr;
    o.s = __sinf(pf) * b3;
    o.c = __cosf(pf) * b3;
    o.sth = sth;
    o.cth = cth;
    o.kap = (float)(qbin * r);
    o.r2 = r2f;
    o.dx = (float)dxd;
    o.dy = (float)dyd;
    o.dz = (float)dzd;
}

__device__ __forceinline__ void pair_setup(PairState<double> &o, double dxd,
                                           double dyd, double dzd, bool keep,
                                           double qbin, double qbin_turns,
                                           int m0)
{
    const double r2 = fma(dxd, dxd, fma(dyd, dyd, dzd * dzd));
    const bool ok = keep && r2 > 0.0;
    const double r = sqrt(r2);
    const double invr = ok ? 1.0 / r : 0.0;
    const double u = r * qbin_turns;
    double sth, cth, s0, c0;
    sincospi(2.0 * (u - rint(u)), &sth, &cth);
    double ph = u * (double)m0;
    ph -= rint(ph);
    sincospi(2.0 * ph, &s0, &c0);
    const double b3 = invr * invr * invr;
    o.s = s0 * b3;
    o.c = c0 * b3;
    o.sth = sth;
    o.cth = cth;
    o.kap = qbin * r;
    o.r2 = r2;
    o.dx = dxd;
    o.dy = dyd;
    o.dz = dzd;
}

__device__ __forceinline__ float fma_t(float a, float b, float c)
{
    return fmaf(a, b, c);
}
__device__ __forceinline__ double fma_t(double a, double b, double c)
{
    return fma(a, b, c);
}

template <typename T, int C, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) debye_kernel(const DebyeParams p)
{
    __shared__ double sx[TJ], sy[TJ], sz[TJ];
    __shared__ float sv[TJ];
    __shared__ double sfj[MODE == MODE_FORCE ? 3 * TJ : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int chunk = blockIdx.y * nwarp + warp;
    const int m0 = chunk * C;
    const bool active = m0 < p.nq;  // idle warps still take part in barriers

    const WorkItem it = p.items[p.item_begin + (int64_t)blockIdx.x * p.item_stride];
    const bool diag = (it.info & ITEM_DIAG) != 0;
    const int btype = it.info & 0xffff;
    const int atype = p.tile_type[it.itile];
    const int gi = it.itile * TILE_I + lane;
    const double xi = p.x[gi], yi = p.y[gi], zi = p.z[gi];
    const bool vi = p.valid[gi] != 0.f;

    const T *ftab = reinterpret_cast<const T *>(p.ftab);
    const T *fa = ftab + (size_t)atype * p.qp;
    const T *fb = ftab + (size_t)btype * p.qp;
    const T *inv_na = reinterpret_cast<const T *>(p.inv_na);

    // accumulators of this thread's (atom, chunk)
    T accF[MODE != MODE_FORCE ? C : 1];
    T accX[MODE == MODE_GRAD ? C : 1], accY[MODE == MODE_GRAD ? C : 1],
        accZ[MODE == MODE_GRAD ? C : 1];
    T w0[MODE == MODE_FORCE ? C : 1], w1[MODE == MODE_FORCE ? C : 1];
    T fix = 0, fiy = 0, fiz = 0;
#pragma unroll
    for (int m = 0; m < C; ++m) {
        if constexpr (MODE != MODE_FORCE) accF[m] = 0;
        if constexpr (MODE == MODE_GRAD) { accX[m] = 0; accY[m] = 0; accZ[m] = 0; }
        if constexpr (MODE == MODE_FORCE) {
            const int bin = m0 + m;
            T w = 0;
            if (bin < p.nq)
                w = (T)(p.wq[bin]) * fa[bin] * fb[bin] * inv_na[bin];
            w0[m] = w;
            w1[m] = w * (T)bin;
        }
    }

    for (int jt = it.jbegin; jt < it.jend; jt += TJ) {
        const int cnt = min(TJ, it.jend - jt);
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
            if (t < cnt) {
                sx[t] = p.x[jt + t];
                sy[t] = p.y[jt + t];
                sz[t] = p.z[jt + t];
                sv[t] = p.valid[jt + t];
            }
            if constexpr (MODE == MODE_FORCE) {
                sfj[3 * t] = 0.0; sfj[3 * t + 1] = 0.0; sfj[3 * t + 2] = 0.0;
            }
        }
        __syncthreads();
        if (active) {
            // software pipeline: the set-up of pair jj+1 (a long dependent
            // chain through the FP64 / conversion / SFU pipes) is issued
            // alongside the FP32 bin loop of pair jj
            PairState<T> nxt;
            pair_setup(nxt, sx[0] - xi, sy[0] - yi, sz[0] - zi,
                       vi && sv[0] != 0.f, p.qbin, p.qbin_turns, m0);
            for (int jj = 0; jj < cnt; ++jj) {
                const PairState<T> ps = nxt;
                {
                    const int jn = min(jj + 1, cnt - 1);
                    pair_setup(nxt, sx[jn] - xi, sy[jn] - yi, sz[jn] - zi,
                               vi && sv[jn] != 0.f, p.qbin, p.qbin_turns, m0);
                }
                T s = ps.s, c = ps.c;
                T mk = ps.kap * (T)m0;
                T p0 = 0, p1 = 0;
#pragma unroll
                for (int m = 0; m < C; ++m) {
                    if constexpr (MODE != MODE_FORCE) accF[m] = fma_t(s, ps.r2, accF[m]);
                    if constexpr (MODE == MODE_GRAD) {
                        const T a = fma_t(mk, c, -s);
                        accX[m] = fma_t(a, ps.dx, accX[m]);
                        accY[m] = fma_t(a, ps.dy, accY[m]);
                        accZ[m] = fma_t(a, ps.dz, accZ[m]);
                        mk += ps.kap;
                    }
                    if constexpr (MODE == MODE_FORCE) {
                        p1 = fma_t(w1[m], c, p1);
                        p0 = fma_t(w0[m], s, p0);
                    }
                    const T sn = fma_t(s, ps.cth, c * ps.sth);
                    const T cn = fma_t(c, ps.cth, -(s * ps.sth));
                    s = sn;
                    c = cn;
                }
                if constexpr (MODE == MODE_FORCE) {
                    // sum_m w (m kappa c - s) = kappa*p1 - p0
                    const T phi = fma_t(ps.kap, p1, -p0);
                    fix = fma_t(phi, ps.dx, fix);
                    fiy = fma_t(phi, ps.dy, fiy);
                    fiz = fma_t(phi, ps.dz, fiz);
                    if (!diag) {
                        // Newton's third law for the j atom: reduce over the
                        // 32 i atoms of this warp, then one shared atomic
                        const double jx = warp_sum((double)(-phi * ps.dx));
                        const double jy = warp_sum((double)(-phi * ps.dy));
                        const double jz = warp_sum((double)(-phi * ps.dz));
                        if (lane == 0) {
                            atomicAdd(&sfj[3 * jj], jx);
                            atomicAdd(&sfj[3 * jj + 1], jy);
                            atomicAdd(&sfj[3 * jj + 2], jz);
                        }
                    }
                }
            }
        }
        if constexpr (MODE == MODE_FORCE) if (!diag) {
            __syncthreads();
            for (int t = threadIdx.x; t < 3 * cnt; t += blockDim.x) {
                const int oj = p.orig[jt + t / 3];
                if (oj >= 0) atomicAdd(&p.force[(size_t)oj * 3 + t % 3], sfj[t]);
            }
        }
    }

    if (!active) return;
    const int oi = p.orig[gi];
    if constexpr (MODE == MODE_FORCE) {
        if (oi >= 0) {
            atomicAdd(&p.force[(size_t)oi * 3 + 0], (double)fix);
            atomicAdd(&p.force[(size_t)oi * 3 + 1], (double)fiy);
            atomicAdd(&p.force[(size_t)oi * 3 + 2], (double)fiz);
        }
        return;
    } else {
    // flush: f_a f_b (/na) factors applied once per (atom, chunk, item)
    const double fweight = (MODE == MODE_GRAD || diag) ? 0.5 : 1.0;
    T *G = reinterpret_cast<T *>(p.G);
#pragma unroll
    for (int m = 0; m < C; ++m) {
        const int bin = m0 + m;
        if (bin < p.nq) {  // warp-uniform
            const T ff = fa[bin] * fb[bin];
            if constexpr (MODE == MODE_GRAD) if (oi >= 0) {
                const T sc = ff * inv_na[bin];
                T *row = G + (size_t)oi * 3 * p.nq + bin;
                atomicAdd(row, accX[m] * sc);
                atomicAdd(row + p.nq, accY[m] * sc);
                atomicAdd(row + 2 * (size_t)p.nq, accZ[m] * sc);
            }
            if (p.S != nullptr) {
                const double v = warp_sum((double)accF[m] * (double)ff);
                if (lane == 0) atomicAdd(&p.S[bin], fweight * v);
            }
        }
    }
    }
}

}  // namespace iid
