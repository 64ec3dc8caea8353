// iid_force_table.cuh -- the fused force pass as a tabulated radial function.
//
// The force of the Rw / chi^2 potential is
//   force[i] = sum_j (q_j - q_i) * Phi_ab(r_ij),
//   Phi_ab(r) = sum_m wq[m] f_a(m) f_b(m)/na(m) * (Q_m r cos(Q_m r) - sin(Q_m r)) / r^3,
// i.e. for fixed chain-rule weights wq the sum over the Q bins depends on the
// pair only through r and the two element types.  Summing over m FIRST turns
// the O(N^2 Q) force pass into
//   1. phi_table_kernel:  Phi_ab on a uniform r grid, O(K Q) (K ~ r_max / 1e-3 A),
//   2. force_table_kernel: one cubic interpolation per ordered pair, O(N^2).
// Phi is band limited by Q_max, so with h = 0.05 / Q_max the 4-point Lagrange
// interpolation error is ~0.023 (Q_max h)^4 |Phi| = 1.5e-7 |Phi|, and the table
// is stored in float32 (6e-8) -- both far below the FP32-mode tolerance of
// 1e-5; geometry, interpolation and accumulation are float64.  FP64 mode and
// small structures (where the table costs more than it saves) keep the direct
// kernel.  The reference gets the same forces
// from the N x 3 x R gradient array (calc/calc_1d.py:89-95 with
// master_kernel.get_grad_rw :293-347).
#pragma once
#include "iid_debye.cuh"

namespace iid {

constexpr int PHI_KMAX = 1 << 18;  // table entries per element pair (r = 0 .. (KMAX-1) h)
constexpr int PHI_PAD = 2;         // entries stored before r = 0 / after the end

// info[0] = h, info[1] = 1/h, info[2] = entries in use
__global__ void phi_grid_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, const float *__restrict__ valid,
                                int np, double h_target, double *__restrict__ info)
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
    if (threadIdx.x == 0) {
        double d2 = 0.0;
        const int nw = blockDim.x >> 5;
        for (int a = 0; a < 3; ++a) {
            double l = 1e300, u = -1e300;
            for (int k = 0; k < nw; ++k) { l = fmin(l, smin[a][k]); u = fmax(u, smax[a][k]); }
            const double e = u > l ? u - l : 0.0;
            d2 += e * e;
        }
        const double rmax = sqrt(d2) * 1.0000001 + 1e-9;  // longest possible pair distance
        // the grid step is never widened (the interpolation error grows like
        // (Q_max h)^4): a structure whose bounding-box diagonal exceeds the
        // table (524 A at Q_max = 25) has its distant pairs summed directly
        // over the Q bins by force_table_kernel
        const double h = h_target;
        info[0] = h;
        info[1] = 1.0 / h;
        info[2] = fmin(ceil(rmax / h) + 2.0, (double)(PHI_KMAX - 4));
    }
}

// PT_G lanes per table entry, each summing a contiguous share of the Q bins
// from its own phase seed (a 330-step sequential float64 recurrence per thread
// left small structures latency bound); blockIdx.y = a * ntypes + b, the blocks
// of a row stride over the groups of PT_EPB entries in use.
constexpr int PT_G = 8;
constexpr int PT_EPB = 128 / PT_G;  // entries per block and pass
__global__ void __launch_bounds__(128) phi_table_kernel(const double *__restrict__ wq,
                                                        const float *__restrict__ ftab,
                                                        const float *__restrict__ inv_na,
                                                        int nq, int qp, int ntypes, double qbin,
                                                        const double *__restrict__ info,
                                                        float *__restrict__ tab)
{
    extern __shared__ double wab[];  // [nq] weights of this element pair
    const int a = blockIdx.y / ntypes, b = blockIdx.y % ntypes;
    const int last = (int)info[2] + 2 * PHI_PAD;  // last entry in use
    if ((int)blockIdx.x * PT_EPB > last) return;  // block-uniform
    for (int m = threadIdx.x; m < nq; m += blockDim.x)
        wab[m] = wq[m] * (double)ftab[(size_t)a * qp + m] * (double)ftab[(size_t)b * qp + m] *
                 (double)inv_na[m];
    __syncthreads();
    const int g = threadIdx.x & (PT_G - 1);
    const int per = (nq + PT_G - 1) / PT_G;
    const int mb = min(nq, g * per), me = min(nq, mb + per);
    const double h = info[0];
    for (int e0 = blockIdx.x * PT_EPB; e0 <= last; e0 += gridDim.x * PT_EPB) {
        const int e = e0 + (threadIdx.x / PT_G);               // entry, r = (e - PHI_PAD) h
        const double r = fabs((double)(e - PHI_PAD)) * h;      // Phi is even in r
        const bool series = qbin * (double)nq * r < 0.05;      // uniform over the PT_G lanes
        double phi = 0.0;
        if (series) {
            // (x cos x - sin x)/r^3 = Q^3 (-1/3 + x^2/30 - x^4/840), x = Q r
            for (int m = mb; m < me; ++m) {
                const double q = qbin * (double)m, xx = q * r * q * r;
                phi += wab[m] * q * q * q *
                       (-1.0 / 3.0 + xx * (1.0 / 30.0 - xx * (1.0 / 840.0)));
            }
        } else {
            const double turns = qbin * r * 0.15915494309189533577;  // theta / 2 pi
            double sth, cth, s, c;
            sincospi(2.0 * (turns - rint(turns)), &sth, &cth);
            const double ph = turns * (double)mb;
            sincospi(2.0 * (ph - rint(ph)), &s, &c);
            const double kap = qbin * r;
            double mk = kap * (double)mb;
            for (int m = mb; m < me; ++m) {
                phi = fma(wab[m], fma(mk, c, -s), phi);
                const double sn = fma(s, cth, c * sth);
                const double cn = fma(c, cth, -(s * sth));
                s = sn;
                c = cn;
                mk += kap;
            }
        }
#pragma unroll
        for (int o = PT_G / 2; o > 0; o >>= 1) phi += __shfl_xor_sync(0xffffffffu, phi, o);
        if (!series) phi /= r * r * r;
        if (g == 0 && e <= last)
            tab[(size_t)blockIdx.y * (PHI_KMAX + 2 * PHI_PAD) + e] = (float)phi;
    }
}

// Phi_ab(r) summed directly over the Q bins (float64 rotation recurrence): the
// pairs beyond the tabulated range.
__device__ double phi_direct(double r, int a, int b, const double *__restrict__ wq,
                             const float *__restrict__ ftab, const float *__restrict__ inv_na,
                             int nq, int qp, double qbin)
{
    const double turns = qbin * r * 0.15915494309189533577;
    double sth, cth;
    sincospi(2.0 * (turns - rint(turns)), &sth, &cth);
    double s = 0.0, c = 1.0, mk = 0.0, phi = 0.0;
    const double kap = qbin * r;
    for (int m = 0; m < nq; ++m) {
        const double w = wq[m] * (double)ftab[(size_t)a * qp + m] *
                         (double)ftab[(size_t)b * qp + m] * (double)inv_na[m];
        phi = fma(w, fma(mk, c, -s), phi);
        const double sn = fma(s, cth, c * sth);
        c = fma(c, cth, -(s * sth));
        s = sn;
        mk += kap;
    }
    return phi / (r * r * r);
}

// force[i] = sum_j Phi_{ab}(r_ij) (q_j - q_i): thread = atom i (sorted/padded
// order), the j range [jbegin, jend) of this block row is staged through shared
// memory.  Rows are split over blockIdx.y / ranks; partials meet by atomicAdd.
constexpr int FT_BLOCK = 128;
__global__ void __launch_bounds__(FT_BLOCK) force_table_kernel(
    const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
    const float *__restrict__ valid, const int *__restrict__ orig,
    const int *__restrict__ tile_type, int np, int ntypes, const double *__restrict__ info,
    const float *__restrict__ tab, int jsplit, int row_begin, int row_stride,
    double *__restrict__ force, const double *__restrict__ wq, const float *__restrict__ ftab,
    const float *__restrict__ inv_na, int nq, int qp, double qbin,
    double *__restrict__ fpart)  // [jsplit][np][3]: deterministic partial sums (or null: atomics)
{
    __shared__ double sx[FT_BLOCK], sy[FT_BLOCK], sz[FT_BLOCK];
    __shared__ int st[FT_BLOCK];
    const int row = row_begin + blockIdx.x * row_stride;  // block of FT_BLOCK atoms i
    const int gi = row * FT_BLOCK + threadIdx.x;
    const bool vi = gi < np && valid[gi] != 0.f;
    const double xi = gi < np ? x[gi] : 0.0, yi = gi < np ? y[gi] : 0.0, zi = gi < np ? z[gi] : 0.0;
    const int ta = gi < np ? tile_type[gi / TILE_I] : 0;
    const double inv_h = info[1];
    const int klast = (int)info[2] - 2;  // last interval with all four nodes tabulated
    const size_t tstride = PHI_KMAX + 2 * PHI_PAD;
    const float *taba = tab + (size_t)ta * ntypes * tstride + PHI_PAD;
    // this block's share of the j atoms, in units of one 32-atom tile
    const int nunits = np / TILE_I;
    const int per = (nunits + jsplit - 1) / jsplit;
    const int jb = blockIdx.y * per * TILE_I, je = min(np, jb + per * TILE_I);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int j0 = jb; j0 < je; j0 += FT_BLOCK) {
        const int gj = j0 + threadIdx.x;
        __syncthreads();
        const bool vj = gj < je && valid[gj] != 0.f;
        sx[threadIdx.x] = vj ? x[gj] : 0.0;
        sy[threadIdx.x] = vj ? y[gj] : 0.0;
        sz[threadIdx.x] = vj ? z[gj] : 0.0;
        st[threadIdx.x] = vj ? tile_type[gj / TILE_I] : -1;
        __syncthreads();
        if (!vi) continue;
        const int cnt = min(FT_BLOCK, je - j0);
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const int tb = st[jj];
            const double dx = sx[jj] - xi, dy = sy[jj] - yi, dz = sz[jj] - zi;
            const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
            if (tb < 0 || r2 <= 0.0) continue;  // ghost atom, self pair, r == 0
            double yv = (double)rsqrtf((float)r2);
            yv = yv * fma(-0.5 * r2, yv * yv, 1.5);
            yv = yv * fma(-0.5 * r2, yv * yv, 1.5);  // second Newton step: 1e-15
            const double tpos = r2 * yv * inv_h;
            double phi;
            if (tpos < (double)klast) {
                const int k = (int)tpos;
                const double u = tpos - (double)k;
                const float *tp = taba + (size_t)tb * tstride + k;
                const double pm = tp[-1], p0 = tp[0], p1 = tp[1], p2 = tp[2];
                // 4-point Lagrange on the uniform grid
                const double um = u - 1.0, u2 = u - 2.0, up = u + 1.0;
                phi = -(u * um * u2) * (1.0 / 6.0) * pm + (up * um * u2) * 0.5 * p0 -
                      (up * u * u2) * 0.5 * p1 + (up * u * um) * (1.0 / 6.0) * p2;
            } else {
                phi = phi_direct(r2 * yv, ta, tb, wq, ftab, inv_na, nq, qp, qbin);
            }
            fx = fma(phi, dx, fx);
            fy = fma(phi, dy, fy);
            fz = fma(phi, dz, fz);
        }
    }
    if (fpart != nullptr) {
        if (gi < np) {
            double *f = fpart + ((size_t)blockIdx.y * np + gi) * 3;
            f[0] = vi ? fx : 0.0;
            f[1] = vi ? fy : 0.0;
            f[2] = vi ? fz : 0.0;
        }
    } else if (vi) {
        const int oi = orig[gi];
        atomicAdd(&force[(size_t)oi * 3 + 0], fx);
        atomicAdd(&force[(size_t)oi * 3 + 1], fy);
        atomicAdd(&force[(size_t)oi * 3 + 2], fz);
    }
}

// force[orig[g]] (+)= sum over the j shares of fpart[s][g], in share order, for
// the rows this rank computed (rows row_begin, row_begin + row_stride, ...).
__global__ void force_table_reduce_kernel(const double *__restrict__ fpart, int jsplit, int np,
                                          const int *__restrict__ orig, int row_begin,
                                          int row_stride, double *__restrict__ force)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * np) return;
    const int g = e / 3, w = e - 3 * g;
    const int row = g / FT_BLOCK;
    if (row < row_begin || (row - row_begin) % row_stride != 0) return;
    const int o = orig[g];
    if (o < 0) return;
    double t = 0.0;
    for (int s = 0; s < jsplit; ++s) t += fpart[((size_t)s * np + g) * 3 + w];
    force[(size_t)o * 3 + w] += t;  // the only writer of this element in this pass
}

}  // namespace iid
