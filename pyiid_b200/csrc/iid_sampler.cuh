// iid_sampler.cuh -- device-resident leapfrog step for the HMC / NUTS samplers
// (sm_100a).  Replaces the host arithmetic of pyiid/sim/__init__.py:10-38
// (leapfrog: half kick, drift, force evaluation, half kick, centre) around the
// fused energy + forces sequence, so that a phase-space point (q, p, f) never
// leaves the GPU between leapfrog steps.
//
// States live in numbered slots of one slab: slot s = q [n][3], p [n][3],
// f [n][3] float64.  The per-step parameters (step size, source and
// destination slot, centring) are read from a small device buffer that the
// captured graph refreshes from pinned host memory, so one instantiated graph
// serves every step of a trajectory.
//
// The arithmetic follows numpy's operation order (separate multiply and add,
// true division by the mass) so that positions and momenta equal the
// array-level host path bit for bit; only the kinetic-energy sum has a
// different summation order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iid {

constexpr int LF_CTL = 8;  // step, src, dst, centre flag, cell centre x y z, -
constexpr int LF_CHAIN_MAX = 64;  // steps one fused launch can walk (= IID_LF_CHAIN)

__device__ __forceinline__ double *lf_slot(double *slab, int n, int slot, int which)
{
    return slab + ((size_t)slot * 3 + which) * 3 * (size_t)n;
}

// p_half = p + (step/2) f;  q' = q + step (p_half / m), fused with the staging of
// the evaluation (prep_kernel): thread k = slot k of the element-sorted, padded
// atom order.  Writes p_half into the destination slot, q' into `pos` (caller
// order) and into the sorted x / y / z arrays (float32-rounded in FP32 mode),
// and clears the accumulators S and force of the passes that follow.
__global__ void lf_stage_kernel(const double *__restrict__ ctl, double *__restrict__ slab,
                                const double *__restrict__ mass, int n,
                                const int *__restrict__ orig, int np, int round_f32,
                                double *__restrict__ pos, double *__restrict__ x,
                                double *__restrict__ y, double *__restrict__ z,
                                float *__restrict__ valid, double *__restrict__ zero_a, int na,
                                double *__restrict__ zero_b, int nb)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    for (int e = k; e < na; e += gridDim.x * blockDim.x) zero_a[e] = 0.0;
    for (int e = k; e < nb; e += gridDim.x * blockDim.x) zero_b[e] = 0.0;
    if (k >= np) return;
    const int o = orig[k];
    double c[3] = {0.0, 0.0, 0.0};
    if (o >= 0) {
        const double step = ctl[0];
        const int src = (int)ctl[1], dst = (int)ctl[2];
        const double *q = lf_slot(slab, n, src, 0), *p = lf_slot(slab, n, src, 1),
                     *f = lf_slot(slab, n, src, 2);
        double *pd = lf_slot(slab, n, dst, 1);
        const double half = __dmul_rn(0.5, step), m = mass[o];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const size_t e = 3 * (size_t)o + w;
            const double ph = __dadd_rn(p[e], __dmul_rn(half, f[e]));
            pd[e] = ph;
            const double qn = __dadd_rn(q[e], __dmul_rn(step, __ddiv_rn(ph, m)));
            pos[e] = qn;
            c[w] = round_f32 ? (double)(float)qn : qn;
        }
    }
    x[k] = c[0];
    y[k] = c[1];
    z[k] = c[2];
    valid[k] = o >= 0 ? 1.f : 0.f;
}

// p = p_half + (step/2) f_new;  KE = sum p.p/m / 2;  q = q' + (cell centre -
// (min + max)/2) when centring.  One block: the sampler's structures are small
// and the three reductions need the whole array.  Mirrors q, p and the scalars
// into one contiguous buffer for the single device-to-host copy of the step.
// (Also the last phase of the fused evaluation kernel, where p, pos, force and
// out4 were written by other blocks of the SAME launch: they are read with
// ld.cg, never through the non-coherent path.)
__device__ __forceinline__ void lf_finish_body(const double *ctl, double *slab,
                                               const double *__restrict__ mass, int n,
                                               const double *pos, const double *force,
                                               double *mirror, const double *out4)
{
    // mirror = q [3n] | p [3n] | energy, scale, value, scale_true, restraint
    // energy, kinetic energy, shift x y z: ONE device-to-host copy per step
    double *out = mirror + 6 * (size_t)n;
    __shared__ double red[7][32];
    __shared__ double shift[3];
    const double step = ctl[0];
    const int dst = (int)ctl[2];
    const bool centre = ctl[3] != 0.0;
    double *q = lf_slot(slab, n, dst, 0), *p = lf_slot(slab, n, dst, 1),
           *f = lf_slot(slab, n, dst, 2);
    const double half = __dmul_rn(0.5, step);
    double ke = 0.0;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const double m = mass[a];
        // all loads of the atom first: the stores below may alias them as far as
        // the compiler knows, and one dependent L2 round trip per component made
        // this phase 8 us of a 56 us step
        double fk[3], pk[3], xk[3];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            fk[w] = __ldcg(force + 3 * a + w);
            pk[w] = __ldcg(p + 3 * a + w);
            xk[w] = __ldcg(pos + 3 * a + w);
        }
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int k = 3 * a + w;
            const double pn = __dadd_rn(pk[w], __dmul_rn(half, fk[w]));
            p[k] = pn;
            f[k] = fk[w];
            mirror[3 * (size_t)n + k] = pn;
            ke = fma(pn, __ddiv_rn(pn, m), ke);
            lo[w] = fmin(lo[w], xk[w]);
            hi[w] = fmax(hi[w], xk[w]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[7] = {ke, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            v[1 + w] = fmin(v[1 + w], __shfl_xor_sync(0xffffffffu, v[1 + w], o));
            v[4 + w] = fmax(v[4 + w], __shfl_xor_sync(0xffffffffu, v[4 + w], o));
        }
    }
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 7; ++c) red[c][warp] = v[c];
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double r[7];
        r[0] = lane < nw ? red[0][lane] : 0.0;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            r[1 + w] = lane < nw ? red[1 + w][lane] : 1e300;
            r[4 + w] = lane < nw ? red[4 + w][lane] : -1e300;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            r[0] += __shfl_xor_sync(0xffffffffu, r[0], o);
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                r[1 + w] = fmin(r[1 + w], __shfl_xor_sync(0xffffffffu, r[1 + w], o));
                r[4 + w] = fmax(r[4 + w], __shfl_xor_sync(0xffffffffu, r[4 + w], o));
            }
        }
        if (lane < 5) out[lane] = __ldcg(out4 + lane);
        if (lane == 0) {
            out[5] = 0.5 * r[0];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                // numpy: q + (centre - 0.5 * (min + max))
                const double s = centre ? __dsub_rn(ctl[4 + w], __dmul_rn(0.5, __dadd_rn(r[1 + w], r[4 + w]))) : 0.0;
                shift[w] = s;
                out[6 + w] = s;
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * n; k += blockDim.x) {
        const double x0 = __ldcg(pos + k);
        const double x = centre ? __dadd_rn(x0, shift[k % 3]) : x0;
        q[k] = x;
        mirror[k] = x;
    }
}

// The same second half of the step spread over a whole grid (the fused
// evaluation kernel): every block computes the extent of the drifted positions
// for itself (min / max are exact in any order) -- the centring shift follows
// from the bounding box -- then element e of the state is finished by whichever
// thread holds its force.  The kinetic energy is summed by the host from the
// mirrored momenta.
//
// ext (shared, [8]): bounding box lo x y z, hi x y z of the drifted positions
// (phase 0 of the fused kernel).
// numpy: q + (centre - 0.5 * (min + max))
__device__ __forceinline__ double lf_shift_of(const double *ctl, const double *ext, int w)
{
    return ctl[3] != 0.0 ? __dsub_rn(ctl[4 + w], __dmul_rn(0.5, __dadd_rn(ext[w], ext[3 + w]))) : 0.0;
}

// Second half kick and centring of element e into the destination state; the
// new coordinate and momentum are returned for the host mirror.
__device__ __forceinline__ void lf_kick(const double *ctl, double *slab, int n, const double *pos,
                                        const double *shift, size_t e, int w, double f, double &x,
                                        double &pn)
{
    const double step = ctl[0];
    const int dst = (int)ctl[2];
    const bool centre = ctl[3] != 0.0;
    double *qd = lf_slot(slab, n, dst, 0), *pd = lf_slot(slab, n, dst, 1),
           *fd = lf_slot(slab, n, dst, 2);
    const double ph = __ldcg(pd + e), x0 = __ldcg(pos + e);
    pn = __dadd_rn(ph, __dmul_rn(__dmul_rn(0.5, step), f));
    x = centre ? __dadd_rn(x0, shift[w]) : x0;
    pd[e] = pn;
    fd[e] = f;
    qd[e] = x;
}

__global__ void __launch_bounds__(1024) lf_finish_kernel(const double *__restrict__ ctl,
                                                         double *__restrict__ slab,
                                                         const double *__restrict__ mass, int n,
                                                         const double *__restrict__ pos,
                                                         const double *__restrict__ force,
                                                         double *__restrict__ mirror,
                                                         const double *__restrict__ out4)
{
    lf_finish_body(ctl, slab, mass, n, pos, force, mirror, out4);
}

}  // namespace iid
