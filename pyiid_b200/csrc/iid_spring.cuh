// iid_spring.cuh -- spring restraint potentials (sm_100a).
//
// The restraint calculators that are summed with the Rw / chi^2 potential in a
// refinement (reference calc/spring_calc.py):
//   rep: pairs closer than rt repel     E = sum_{i != j, r < rt} k/2 (r - rt)^2
//   att: pairs further than rt attract  E = sum_{i != j, r > rt} k/2 (r - rt)^2
//   com: atoms further than rt from the centre of mass are pulled back
// with force[i] = sum_j (q_j - q_i)/r * k (r - rt) (ordered pairs, as the
// reference's N x N arrays).  The reference builds N x N x 3 numpy arrays per
// call; here thread = atom i, the j atoms are staged through shared memory and
// nothing of size N^2 exists.  O(N^2) with ~20 FP32/FP64 instructions per
// ordered pair and 24 B/atom of traffic: issue bound, a few microseconds at
// the sampler sizes, which is why it rides in the same CUDA graph as the
// fused energy+forces sequence.
//
// F32 = true reproduces the reference's arithmetic operation by operation
// (positions rounded to float32, float32 differences / r^2 accumulation /
// sqrt / (r - rt) / k (r - rt), float64 from there on, spring_calc.py:107-147);
// F32 = false is the same formulas in float64 throughout (FP64 mode).
#pragma once
#include "iid_debye.cuh"

namespace iid {

constexpr int SP_BLOCK = 128;
enum SpringType { SPRING_REP = 0, SPRING_COM = 1, SPRING_ATT = 2 };

template <bool F32> struct SpringReal { using type = double; };
template <> struct SpringReal<true> { using type = float; };

// one ordered pair: hit, (r - rt), mag = k (r - rt), unit direction d / r
template <bool F32> struct SpringPair {
    using R = typename SpringReal<F32>::type;
    R dx, dy, dz, r, dr, mag;
    bool hit;
    __device__ __forceinline__ SpringPair(R xi, R yi, R zi, R xj, R yj, R zj, R k, R rt,
                                          bool att, bool self)
    {
        if constexpr (F32) {
            // kernels/cpu_nxn.py:17-52: d = q_j - q_i, tmp += d*d (no FMA), sqrt
            dx = __fsub_rn(xj, xi);
            dy = __fsub_rn(yj, yi);
            dz = __fsub_rn(zj, zi);
            float t = __fmul_rn(dx, dx);
            t = __fadd_rn(t, __fmul_rn(dy, dy));
            t = __fadd_rn(t, __fmul_rn(dz, dz));
            r = __fsqrt_rn(t);
            dr = __fsub_rn(r, rt);
            mag = __fmul_rn(k, dr);
        } else {
            dx = __dsub_rn(xj, xi);
            dy = __dsub_rn(yj, yi);
            dz = __dsub_rn(zj, zi);
            double t = __dmul_rn(dx, dx);
            t = __dadd_rn(t, __dmul_rn(dy, dy));
            t = __dadd_rn(t, __dmul_rn(dz, dz));
            r = __dsqrt_rn(t);
            dr = __dsub_rn(r, rt);
            mag = __dmul_rn(k, dr);
        }
        hit = !self && (att ? r > rt : r < rt);
    }
    __device__ __forceinline__ double unit(R d) const
    {
        if constexpr (F32) return (double)__fdiv_rn(d, r);
        else return __ddiv_rn(d, r);
    }
};

__device__ __forceinline__ double block_sum_128(double v, double *sh)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < SP_BLOCK / 32; ++k) t += sh[k];
    __syncthreads();
    return t;  // valid in thread 0
}

// rep / att.  pos = [n,3] float64 in the caller's atom order.  Block row =
// SP_BLOCK atoms i (rows row_begin, row_begin + row_stride, ... belong to this
// rank); blockIdx.y takes a share of the j atoms in units of SP_JUNIT.  Results
// are ADDED to energy[0], force[n,3] and atomwise[n] (each may be null).
//
// Most ordered pairs are no hit (rep with a short rt), so four j atoms are
// screened at a time on the squared distance alone -- r = sqrt_rn(t) is
// monotonic in t, so t beyond rt^2 (1 +- margin) decides the comparison
// without the square root -- and only candidate pairs take the exact path.
constexpr int SP_JUNIT = 32;
template <bool F32> struct SpringVec { using type = double4; };
template <> struct SpringVec<true> { using type = float4; };

template <bool F32>
__global__ void __launch_bounds__(SP_BLOCK) spring_pair_kernel(
    const double *__restrict__ pos, int n, int att, double k, double rt, int jsplit,
    int row_begin, int row_stride, double *__restrict__ energy, double *__restrict__ force,
    double *__restrict__ atomwise)
{
    using R = typename SpringReal<F32>::type;
    using V = typename SpringVec<F32>::type;
    __shared__ V sq[SP_BLOCK];
    __shared__ double red[SP_BLOCK / 32];
    const int row = row_begin + blockIdx.x * row_stride;
    const int i = row * SP_BLOCK + threadIdx.x;
    const bool vi = i < n;
    const R xi = vi ? (R)pos[(size_t)i * 3 + 0] : (R)0;
    const R yi = vi ? (R)pos[(size_t)i * 3 + 1] : (R)0;
    const R zi = vi ? (R)pos[(size_t)i * 3 + 2] : (R)0;
    const R kk = (R)k, rtt = (R)rt;
    const R half_k = (R)(0.5 * k);
    // screening bound on t = r^2: the margin is far above the rounding of
    // t and of the square root (6e-8 / 1e-16), so no hit is ever screened out
    const double margin = F32 ? 1e-5 : 1e-12;
    const R thr = (R)((double)rtt * (double)rtt * (att ? 1.0 - margin : 1.0 + margin));
    const int nunits = (n + SP_JUNIT - 1) / SP_JUNIT;
    const int per = (nunits + jsplit - 1) / jsplit;
    const int jb = blockIdx.y * per * SP_JUNIT, je = min(n, jb + per * SP_JUNIT);
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0, aw = 0.0;

    auto exact = [&](const V &q, int j) {
        const SpringPair<F32> p(xi, yi, zi, q.x, q.y, q.z, kk, rtt, att != 0, j == i);
        if (!p.hit) return;
        const double mag = (double)p.mag, dr = (double)p.dr;
        e += mag / 2. * dr;  // spring_calc.py:121
        if (p.r > (R)0) {    // 0/0 -> NaN -> 0 (spring_calc.py:143)
            fx = fma(p.unit(p.dx), mag, fx);
            fy = fma(p.unit(p.dy), mag, fy);
            fz = fma(p.unit(p.dz), mag, fz);
        }
        // atomwise (spring_calc.py:171-185): .5 k (r - rt)^2 in R
        if constexpr (F32) aw += (double)__fmul_rn(half_k, __fmul_rn(p.dr, p.dr));
        else aw += __dmul_rn(half_k, __dmul_rn(p.dr, p.dr));
    };
    auto t_of = [&](const V &q) -> R {
        if constexpr (F32) {
            const float dx = __fsub_rn(q.x, xi), dy = __fsub_rn(q.y, yi), dz = __fsub_rn(q.z, zi);
            return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        } else {
            const double dx = __dsub_rn(q.x, xi), dy = __dsub_rn(q.y, yi),
                         dz = __dsub_rn(q.z, zi);
            return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        }
    };
    auto maybe = [&](R t) { return att ? t > thr : t < thr; };

    for (int j0 = jb; j0 < je; j0 += SP_BLOCK) {
        const int gj = j0 + threadIdx.x;
        __syncthreads();
        if (gj < je) {
            V q;
            q.x = (R)pos[(size_t)gj * 3 + 0];
            q.y = (R)pos[(size_t)gj * 3 + 1];
            q.z = (R)pos[(size_t)gj * 3 + 2];
            q.w = (R)0;
            sq[threadIdx.x] = q;
        }
        __syncthreads();
        if (!vi) continue;
        const int cnt = min(SP_BLOCK, je - j0);
        int jj = 0;
        for (; jj + 4 <= cnt; jj += 4) {
            const V q0 = sq[jj], q1 = sq[jj + 1], q2 = sq[jj + 2], q3 = sq[jj + 3];
            const R t0 = t_of(q0), t1 = t_of(q1), t2 = t_of(q2), t3 = t_of(q3);
            const bool m0 = maybe(t0), m1 = maybe(t1), m2 = maybe(t2), m3 = maybe(t3);
            if (m0 | m1 | m2 | m3) {
                if (m0) exact(q0, j0 + jj);
                if (m1) exact(q1, j0 + jj + 1);
                if (m2) exact(q2, j0 + jj + 2);
                if (m3) exact(q3, j0 + jj + 3);
            }
        }
        for (; jj < cnt; ++jj) {
            const V q = sq[jj];
            if (maybe(t_of(q))) exact(q, j0 + jj);
        }
    }
    if (vi) {
        if (force && (fx != 0.0 || fy != 0.0 || fz != 0.0)) {
            atomicAdd(&force[(size_t)i * 3 + 0], fx);
            atomicAdd(&force[(size_t)i * 3 + 1], fy);
            atomicAdd(&force[(size_t)i * 3 + 2], fz);
        }
        if (atomwise && aw != 0.0) atomicAdd(&atomwise[i], -2.0 * aw);
    }
    if (energy) {
        const double tot = block_sum_128(e, red);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(energy, tot);
    }
}

// com (spring_calc.py:188-236): disp = q - com in float64 (float32 positions
// in FP32 mode), restraint outside the sphere of radius rt.
template <bool F32>
__global__ void __launch_bounds__(SP_BLOCK) spring_com_kernel(
    const double *__restrict__ pos, int n, double k, double rt, double cx, double cy, double cz,
    int row_begin, int row_stride, double *__restrict__ energy, double *__restrict__ force,
    double *__restrict__ atomwise)
{
    using R = typename SpringReal<F32>::type;
    __shared__ double red[SP_BLOCK / 32];
    const int row = row_begin + blockIdx.x * row_stride;
    const int i = row * SP_BLOCK + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        const double dx = __dsub_rn((double)(R)pos[(size_t)i * 3 + 0], cx);
        const double dy = __dsub_rn((double)(R)pos[(size_t)i * 3 + 1], cy);
        const double dz = __dsub_rn((double)(R)pos[(size_t)i * 3 + 2], cz);
        const double dist = __dsqrt_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double dr = dist - rt;
        if (dist > rt) {
            const double mag = k * dr;
            e = mag / 2. * dr;
            if (force) {
                atomicAdd(&force[(size_t)i * 3 + 0], -(dx / dist * mag));
                atomicAdd(&force[(size_t)i * 3 + 1], -(dy / dist * mag));
                atomicAdd(&force[(size_t)i * 3 + 2], -(dz / dist * mag));
            }
        }
        // spring_calc.py:259-266 zeroes dist < rt only
        if (atomwise && !(dist < rt)) atomicAdd(&atomwise[i], 2.0 * (.5 * k * dr * dr));
    }
    if (energy) {
        const double tot = block_sum_128(e, red);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(energy, tot);
    }
}

// Energy added by a probe atom at every voxel centre of an nx x ny x nz grid
// (spring_calc.py:150-168, :240-256, :314-332).  Thread = voxel (C order).
template <bool F32>
__global__ void __launch_bounds__(SP_BLOCK) spring_voxel_kernel(
    const double *__restrict__ pos, int n, int sp_type, double k, double rt, double cx, double cy,
    double cz, double res, int nx, int ny, int nz, double *__restrict__ voxels)
{
    using R = typename SpringReal<F32>::type;
    __shared__ double sx[SP_BLOCK], sy[SP_BLOCK], sz[SP_BLOCK];
    const long long v = (long long)blockIdx.x * SP_BLOCK + threadIdx.x;
    const long long nv = (long long)nx * ny * nz;
    const int iz = (int)(v % nz), iy = (int)((v / nz) % ny), ix = (int)(v / ((long long)nz * ny));
    const double x = (ix + .5) * res, y = (iy + .5) * res, z = (iz + .5) * res;
    double acc = 0.0;
    if (sp_type == SPRING_COM) {
        const double ax = x - cx, ay = y - cy, az = z - cz;
        const double temp = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)),
                                           __dmul_rn(az, az)));
        if (temp > rt) acc = .5 * k * ((temp - rt) * (temp - rt));
    } else {
        for (int j0 = 0; j0 < n; j0 += SP_BLOCK) {
            const int gj = j0 + threadIdx.x;
            __syncthreads();
            sx[threadIdx.x] = gj < n ? (double)(R)pos[(size_t)gj * 3 + 0] : 0.0;
            sy[threadIdx.x] = gj < n ? (double)(R)pos[(size_t)gj * 3 + 1] : 0.0;
            sz[threadIdx.x] = gj < n ? (double)(R)pos[(size_t)gj * 3 + 2] : 0.0;
            __syncthreads();
            const int cnt = min(SP_BLOCK, n - j0);
            for (int jj = 0; jj < cnt; ++jj) {
                const double ax = x - sx[jj], ay = y - sy[jj], az = z - sz[jj];
                const double temp = __dsqrt_rn(__dadd_rn(
                    __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
                const bool hit = sp_type == SPRING_ATT ? temp > rt : temp < rt;
                if (hit) acc += .5 * k * ((temp - rt) * (temp - rt));
            }
        }
    }
    if (v < nv) voxels[v] = acc * 2;
}

}  // namespace iid
