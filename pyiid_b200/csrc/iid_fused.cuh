// iid_fused.cuh -- one cooperative kernel for a whole Rw / chi^2 energy + force
// evaluation of a SMALL structure (the sampler's regime: Au561, one work item
// per SM), sm_100a.
//
// pyiid/sim/__init__.py:10-38 (leapfrog) + calc/calc_1d.py:78-95 cost the
// reference 1 x grad PDF + 2 x PDF per step.  Round 1 ran the evaluation as a
// graph of 6-7 small launches (staging, F(Q) pass, G(r), potential, weights,
// force pass) that at 561 atoms spent more time in launch gaps and in the
// one-block float64 stages than in the two pair sums.  Here the sequence is ONE
// launch with three grid-wide barriers:
//
//   0  staging.  Leapfrog: every block computes the half kick + drift of the
//      ~70 atoms of ITS work item straight into shared memory (no barrier; the
//      owner thread of each atom also writes the drifted position and the half-
//      kicked momentum for phase 4).  Plain evaluation: positions from the
//      caller's pinned buffer to the device, barrier, then the same fill.
//   1  F(Q) pass of the block's work item (debye2_body<MODE_FQ>), partial sums
//      added to FIXED-POINT accumulators: integer atomics commute, so the sum
//      is bit-reproducible without per-item partials or an ordered second pass
//      --- barrier ---
//   2  F = 2 S / na;  (M F)[m] for this block's rows of M = T^T T
//      --- barrier ---
//   3  every block, redundantly: Rw / chi^2, scale and the chain-rule weights
//      in Q SPACE -- with gc = T F:  gc.go = F.(T^T go),  gc.gc = F.(M F), so
//      neither G(r) nor the R x Q matrix is touched -- then the force pass of
//      its work item (debye2_body<MODE_FORCE>) into fixed-point accumulators
//      whose scale comes from a bound of the force sum (see force_fix_scale)
//      --- barrier ---
//   4  forces to float64; plain evaluation: straight into the caller's pinned
//      buffer; leapfrog: second half kick + centring into the destination state
//      and its host mirror, spread over the grid.  A chain of leapfrog steps
//      loops over 0-4 inside the launch (one more barrier per step).
//
// The Q-space scalars: a = F.vgo, b = F.MF, c = go.go (once per target);
// scale s = a/b (<= 0: the reference's branches, master_kernel.py:229-230,
// 263-264), |go - s gc|^2 = c - 2 s a + s^2 b, gc.(go - s gc) = a - s b.
#pragma once
#include <cooperative_groups.h>
#include "iid_debye2.cuh"
#include "iid_sampler.cuh"
#include "iid_stencil.cuh"
#include "iid_fq_hist.cuh"

namespace iid {

struct FusedParams {
    DebyeParams fq;  // F(Q) pass (Sfix = fixed-point accumulator, fix_scale from the host)
    DebyeParams fo;  // force pass (wq and fix_scale are set per block)
    int n_items;
    int ps_off, pl;           // shared-memory position cache: byte offset, atoms per row
    double s_scale_inv;       // 1 / fq.fix_scale
    const double *gforce;     // [qp] bound of |pair force| per unit |wq| (force_fix_scale)
    // force pass as a radial table built inside the launch (fused_table_*), or null.
    // One table per step of a chain: the lookups go through L1 (the pairs of a
    // work item crowd on a few neighbour distances), which is only safe for
    // addresses written once per launch.
    double *phi_tab;          // [chain step][ntypes^2][phi_stride] float64
    int phi_stride, phi_cap;  // doubles per type pair; entries r = 0 .. (cap-1) h
    int ntypes;
    double tab_h;             // grid step: FT_QH / Q_max
    double *ext_ref;          // [4] box centre of the previous evaluation + valid flag
    // F(Q) phase through the radial pair histogram (iid_fq_hist.cuh), or null: the
    // block spreads the pairs of its work item into a shared-memory histogram,
    // adds it to the global one, and after a barrier every block turns its share
    // of the nodes into partial pair sums (fixed point as the direct pass)
    unsigned long long *fq_C;  // [element pairs][fq_cstride] units of 2^-28, zero between passes
    int fq_cstride;            // nodes per element pair = shared-memory capacity
    // staging
    int lf;  // 1 = leapfrog staging from the state slab, 0 = positions in `pos`
    const double *ctl;
    double *slab;
    const double *mass;
    double *pos;  // [n][3] caller order (device)
    // zero-copy I/O through pinned, mapped host memory (no copy nodes around the
    // launch: at 561 atoms each small DMA costs as much as a phase of the kernel)
    const double *pos_in;  // plain staging: positions in host memory (or null: `pos`)
    double *force_out;     // [n][3] host copy of the forces (or null)
    double *out_host;      // [5] host copy of out4 (or null)
    double *lf_mirror;     // leapfrog: finish the step in this launch (half kick, kinetic
                           // energy, centring) and mirror (q, p, scalars) here (or null)
    int n_chain;           // leapfrog steps walked by this launch (ctl / lf_mirror of step s
    int chain_stride;      // at + s * chain_stride doubles); 1 unless lf_mirror is set
    int n, np, round_f32;
    // Q-space stages
    const double *inv_na_d, *Mq, *vgo;
    double gogo;
    double *MF;      // [qp]
    double *wq_blk;  // [grid][qp]
    int potential;
    double conv;
    double *out4;
    unsigned long long *stamps;  // developer timing: globaltimer at the phase boundaries (or null)
};

__device__ __forceinline__ void fused_stamp(const FusedParams &q, int k, bool on = true)
{
    if (on && q.stamps && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        q.stamps[k] = t;
    }
}

// the latest block to pass this point
__device__ __forceinline__ void fused_stamp_max(const FusedParams &q, int k, bool on = true)
{
    if (on && q.stamps && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(q.stamps + k, t);
    }
}

__device__ __forceinline__ double block_sum(double v, double *sm)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nw; ++k) t += sm[k];  // same order in every block
    return t;
}

// ---- the force pass as a radial table (small structures) ---------------------
// For fixed weights the Q sum of a pair's force depends on the pair only through
// r and the two element types (iid_force_table.cuh):
//   Phi_ab(r) = sum_m w_ab[m] (Q_m r cos(Q_m r) - sin(Q_m r)) / r^3.
// The fused kernel tabulates it in FLOAT64 on a uniform grid with Q_max h = 1/3
// (FT_G lanes per entry, three-term recurrences from exact seeds), barrier, then
// takes one 12-point Lagrange interpolation per pair: error 2.2e-4 (Q_max h)^12
// = 4e-10 |Phi|, so the forces carry the float32 rounding of the positions and
// nothing else.  At Au561 that is 2 200 entries x 330 bins (0.7 M terms)
// instead of 157 000 pairs x 330 bins; a coarse grid with a long stencil keeps
// the build short when a hot trajectory has spread the structure out.
constexpr int FT_G = 8;               // lanes per table entry

__device__ __forceinline__ void fused_table_build(double *tab, int stride, int ntp, int K,
                                                  double h, const double *wab, int nq, int qp,
                                                  double qbin)
{
    const int Kp = K + 2 * FT_PAD, total = ntp * Kp;
    const int epb = blockDim.x / FT_G;  // entries per block and round
    const int g = threadIdx.x & (FT_G - 1);
    const int per = (nq + FT_G - 1) / FT_G;
    const int mb = min(nq, g * per), me = min(nq, mb + per);
    // every block builds an equal, contiguous share of the entries
    const int share = (total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * share, t_end = min(total, t_begin + share);
    for (int t0 = t_begin; t0 < t_end; t0 += epb) {  // block-uniform
        const int t = t0 + threadIdx.x / FT_G;
        const bool ok = t < t_end;
        const int tt = ok ? t : total - 1;
        const int pair = tt / Kp, e = tt - pair * Kp;
        const double r = fabs((double)(e - FT_PAD)) * h;  // Phi is even in r
        const double *w = wab + (size_t)pair * qp;
        const bool series = qbin * (double)nq * r < 0.05;  // uniform over the FT_G lanes
        double phi = 0.0;
        if (series) {
            // (x cos x - sin x)/r^3 = Q^3 (-1/3 + x^2/30 - x^4/840), x = Q r
            for (int m = mb; m < me; ++m) {
                const double qq = qbin * (double)m, xx = qq * r * qq * r;
                phi += w[m] * qq * qq * qq * (-1.0 / 3.0 + xx * (1.0 / 30.0 - xx * (1.0 / 840.0)));
            }
        } else {
            const double turns = qbin * r * 0.15915494309189533577;  // theta / 2 pi
            double sth, cth, sn, cn;
            sincospi(2.0 * (turns - rint(turns)), &sth, &cth);
            const double ph = turns * (double)mb;
            sincospi(2.0 * (ph - rint(ph)), &sn, &cn);
            // three-term recurrences s[m+1] = 2 cos(theta) s[m] - s[m-1] from bins mb-1, mb
            double sp = fma(sn, cth, -(cn * sth)), cp = fma(cn, cth, sn * sth);
            const double tc = cth + cth, kap = qbin * r;
            double mk = kap * (double)mb;
            for (int m = mb; m < me; ++m) {
                phi = fma(w[m], fma(mk, cn, -sn), phi);
                const double s2 = fma(tc, sn, -sp), c2 = fma(tc, cn, -cp);
                sp = sn;
                sn = s2;
                cp = cn;
                cn = c2;
                mk += kap;
            }
        }
#pragma unroll
        for (int o = FT_G / 2; o > 0; o >>= 1) phi += __shfl_xor_sync(0xffffffffu, phi, o);
        if (!series) phi /= r * r * r;
        if (g == 0 && ok) tab[(size_t)pair * stride + e] = phi;
    }
}

// Phi summed directly over the Q bins (float64 rotation recurrence): the pairs
// beyond the tabulated range of an unusually extended structure.
__device__ double fused_phi_direct(double r, const double *w, int nq, double qbin)
{
    const double turns = qbin * r * 0.15915494309189533577;
    double sth, cth;
    sincospi(2.0 * (turns - rint(turns)), &sth, &cth);
    double s = 0.0, c = 1.0, mk = 0.0, phi = 0.0;
    const double kap = qbin * r;
    for (int m = 0; m < nq; ++m) {
        phi = fma(w[m], fma(mk, c, -s), phi);
        const double sn = fma(s, cth, c * sth);
        c = fma(c, cth, -(s * sth));
        s = sn;
        mk += kap;
    }
    return phi / (r * r * r);
}

// FT_PTS-point Lagrange interpolation on the uniform grid: nodes k - FT_LEFT ..
// k + FT_PTS - 1 - FT_LEFT at t[-FT_LEFT] .., u in [0, 1) measured from node k.
// Error of 12 points: max |prod (u - j)| / 12! = 2.2e-4 times (Q_max h)^12.
__device__ __forceinline__ double lagrange_pts(const double *t, double u)
{
    double v[FT_PTS], d[FT_PTS], pre[FT_PTS], L = 0.0;
#pragma unroll
    for (int i = 0; i < FT_PTS; ++i) {
        v[i] = t[i - FT_LEFT];  // (L1: neighbouring pairs share these lines; see phi_tab)
        d[i] = u - (double)(i - FT_LEFT);
    }
    pre[0] = 1.0;
#pragma unroll
    for (int i = 1; i < FT_PTS; ++i) pre[i] = pre[i - 1] * d[i - 1];
    double suf = 1.0;
#pragma unroll
    for (int i = FT_PTS - 1; i >= 0; --i) {
        L = fma(v[i] * ft_bary(i), pre[i] * suf, L);
        suf *= d[i];
    }
    return L;
}

// The force pass of one work item from the table: one warp iteration per j atom
// (lane = atom i); the i side is summed over the warps in shared memory.
__device__ __forceinline__ void fused_table_forces(const FusedParams &q, const WorkItem wi,
                                                   const double *ps, int pl, const double *tab,
                                                   int K, double inv_h, const double *w,
                                                   double fscale, double *fs /* [nw][3][32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool diag = (wi.info & ITEM_DIAG) != 0;
    const int len = wi.jend - wi.jbegin;
    const double xi = ps[lane], yi = ps[pl + lane], zi = ps[2 * pl + lane];
    const bool vi = ps[3 * pl + lane] != 0.0;
    const double klast = (double)(K - FT_PTS);  // every node of the stencil is tabulated
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int jj = warp; jj < len; jj += nw) {
        const int sj = TILE_I + jj;
        const double dx = ps[sj] - xi, dy = ps[pl + sj] - yi, dz = ps[2 * pl + sj] - zi;
        const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
        double phi = 0.0;
        if (vi && ps[3 * pl + sj] != 0.0 && r2 > 0.0) {  // not a ghost atom, self pair, r == 0
            double y = (double)rsqrtf((float)r2);
            y = y * fma(-0.5 * r2, y * y, 1.5);
            y = y * fma(-0.5 * r2, y * y, 1.5);  // second Newton step: 1e-15
            const double r = r2 * y, tpos = r * inv_h;
            if (tpos < klast) {
                const int k = (int)tpos;
                phi = lagrange_pts(tab + k, tpos - (double)k);
            } else {
                phi = fused_phi_direct(r, w, q.fq.nq, q.fq.qbin);
            }
        }
        fx = fma(phi, dx, fx);
        fy = fma(phi, dy, fy);
        fz = fma(phi, dz, fz);
        if (!diag) {  // Newton's third law for the j atom
            const double jx = warp_sum(-phi * dx), jy = warp_sum(-phi * dy),
                         jz = warp_sum(-phi * dz);
            const int oj = q.fq.orig[wi.jbegin + jj];
            if (lane == 0 && oj >= 0) {
                fix_add(q.fo.Ffix + (size_t)oj * 3 + 0, jx, fscale);
                fix_add(q.fo.Ffix + (size_t)oj * 3 + 1, jy, fscale);
                fix_add(q.fo.Ffix + (size_t)oj * 3 + 2, jz, fscale);
            }
        }
    }
    fs[(warp * 3 + 0) * 32 + lane] = fx;
    fs[(warp * 3 + 1) * 32 + lane] = fy;
    fs[(warp * 3 + 2) * 32 + lane] = fz;
    __syncthreads();
    for (int e = threadIdx.x; e < 96; e += blockDim.x) {  // (a block may be one warp)
        const int a = e & 31, c = e >> 5;
        double t = 0.0;
        for (int k = 0; k < nw; ++k) t += fs[(k * 3 + c) * 32 + a];
        const int oa = q.fq.orig[wi.itile * TILE_I + a];
        if (oa >= 0) fix_add(q.fo.Ffix + (size_t)oa * 3 + c, t, fscale);
    }
}

// Power-of-two scale of the fixed-point force accumulators.  One pair adds
// phi * d to a force component with phi = sum_q w_q [x cos x - sin x] / r^3,
// x = Q r, |d| <= r, and |x cos x - sin x| <= min(x^3 / 3, x + 1), so
// |phi d| <= sum_q |w_q| Q^2 max_x min(x / 3, (x + 1) / x^2) < 0.6 sum_q |w_q| Q^2.
// gforce[m] = 0.6 Q_m^2 max_types f^2 / na (host), bound = n sum_m |wq[m]| gforce[m];
// every block computes the same number in the same order.
__device__ __forceinline__ double force_fix_scale(double bound)
{
    if (!(bound > 0.0) || !(bound < 1e300)) return 1.0;
    int e;
    frexp(bound, &e);  // bound < 2^e
    return ldexp(1.0, 60 - e);
}

template <bool CHEB>
__global__ void __launch_bounds__(384, 1) fused_eval_kernel(const FusedParams q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int nq = q.fq.nq, qp = q.fq.qp;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const bool has_item = (int)blockIdx.x < q.n_items;
    double *ps = reinterpret_cast<double *>(smem_raw + q.ps_off);  // [4][pl]
    const int pl = q.pl;
    double *pall = ps + 4 * pl;  // [3][n] all positions, caller order
    __shared__ double ext[8];    // their extent: table range, centring shift
    __shared__ double cref[4];   // reference point of the extent (phase 0)
    if (threadIdx.x < 4) cref[threadIdx.x] = q.ext_ref[threadIdx.x];
    __syncthreads();

  // the step parameters of the whole chain: ONE round trip to the pinned staging
  // per block and launch
  __shared__ double ctl_all[LF_CHAIN_MAX * LF_CTL];
  if (q.lf) {
      for (int k = threadIdx.x; k < q.n_chain * LF_CTL; k += blockDim.x)
          ctl_all[k] = q.ctl[(size_t)(k / LF_CTL) * q.chain_stride + k % LF_CTL];
      __syncthreads();
  }
  const int scs = q.n_chain > 1 ? 1 : 0;  // the step whose phases are stamped (developer timing)
  for (int cs = 0; cs < q.n_chain; ++cs) {
    // a chain of leapfrog steps: step cs mirrors its state to ring slot cs of the
    // pinned staging.  Everything another block wrote in an EARLIER step of this
    // launch is read past L1 (ld.cg).
    const double *ctl = ctl_all + cs * LF_CTL;
    fused_stamp(q, 0, cs == scs);
    fused_stamp(q, 12 + cs, cs < 4);  // start of the first four steps of a chain
    // ---- phase 0: staging ---------------------------------------------------------
    // Every block holds ALL positions in shared memory (13 KB at Au561): their
    // extent (table range, centring shift) and the atoms of its own work item come
    // from there without another trip to L2.
    // Extent of the positions while they pass through registers: bounding box and
    // the largest distance from a reference point c (the box centre of the previous
    // evaluation, i.e. the structure's centre up to one step's drift).  Every pair
    // distance is <= min(box diagonal, 2 max |x - c|) for ANY c; the table range
    // only decides how many entries are built, not their values.
    const bool want_ext = q.phi_tab != nullptr || q.lf_mirror != nullptr || q.fq_C != nullptr;
    double ev[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300}, ed2 = 0.0;
    const double rcx = cref[0], rcy = cref[1], rcz = cref[2];
    auto ext_acc = [&](double x, double y, double z) {
        ev[0] = fmin(ev[0], x); ev[3] = fmax(ev[3], x);
        ev[1] = fmin(ev[1], y); ev[4] = fmax(ev[4], y);
        ev[2] = fmin(ev[2], z); ev[5] = fmax(ev[5], z);
        const double dx = x - rcx, dy = y - rcy, dz = z - rcz;
        ed2 = fmax(ed2, fma(dx, dx, fma(dy, dy, dz * dz)));
    };
    if (q.lf) {
        // leapfrog: half kick + drift from the state slab, computed by every block
        // for itself (no barrier); block (a mod grid) also writes atom a's drifted
        // position and half-kicked momentum for phase 4
        const double step = ctl[0], half = __dmul_rn(0.5, step);
        const int src = (int)ctl[1], dst = (int)ctl[2];
        const double *qq = lf_slot(q.slab, q.n, src, 0), *pp = lf_slot(q.slab, q.n, src, 1),
                     *ff = lf_slot(q.slab, q.n, src, 2);
        double *pd = lf_slot(q.slab, q.n, dst, 1);
        // two atoms per thread and trip, all loads first: the stores below may
        // alias them as far as the compiler knows, and one L2 round trip per atom
        // in a row was 3 us at Au561
        for (int a0 = threadIdx.x; a0 < q.n; a0 += 2 * blockDim.x) {
            double xq[2][3], xp[2][3], xf[2][3], m[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int a = min(a0 + u * (int)blockDim.x, q.n - 1);
                m[u] = q.mass[a];
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const size_t e = 3 * (size_t)a + w;
                    xq[u][w] = __ldcg(qq + e);
                    xp[u][w] = __ldcg(pp + e);
                    xf[u][w] = __ldcg(ff + e);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int a = a0 + u * (int)blockDim.x;
                if (a >= q.n) break;
                const bool owner = a % (int)gridDim.x == (int)blockIdx.x;
                double qn[3];
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    // p_half = p + (step/2) f;  q' = q + step p_half / m (numpy's
                    // operation order, see lf_stage_kernel)
                    const double ph = __dadd_rn(xp[u][w], __dmul_rn(half, xf[u][w]));
                    qn[w] = __dadd_rn(xq[u][w], __dmul_rn(step, __ddiv_rn(ph, m[u])));
                    pall[w * q.n + a] = qn[w];
                    if (owner) {
                        q.pos[3 * (size_t)a + w] = qn[w];
                        pd[3 * (size_t)a + w] = ph;
                    }
                }
                ext_acc(qn[0], qn[1], qn[2]);
            }
        }
        fused_stamp(q, 1, cs == scs);
    } else {
        if (q.pos_in) {
            for (int e = gtid; e < 3 * q.n; e += gsz) q.pos[e] = q.pos_in[e];
            fused_stamp(q, 1, cs == scs);
            grid.sync();
        } else {
            fused_stamp(q, 1, cs == scs);
        }
        for (int a = threadIdx.x; a < q.n; a += blockDim.x) {
            double c[3];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                c[w] = __ldcg(q.pos + 3 * (size_t)a + w);
                pall[w * q.n + a] = c[w];
            }
            ext_acc(c[0], c[1], c[2]);
        }
    }
    __shared__ double extp[7][12];
    if (want_ext) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                ev[w] = fmin(ev[w], __shfl_xor_sync(0xffffffffu, ev[w], o));
                ev[3 + w] = fmax(ev[3 + w], __shfl_xor_sync(0xffffffffu, ev[3 + w], o));
            }
            ed2 = fmax(ed2, __shfl_xor_sync(0xffffffffu, ed2, o));
        }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) extp[c][threadIdx.x >> 5] = ev[c];
            extp[6][threadIdx.x >> 5] = ed2;
        }
    }
    __syncthreads();
    if (want_ext && threadIdx.x < 7) {
        const int c = threadIdx.x, nw = blockDim.x >> 5;
        double v = extp[c][0];
        for (int k = 1; k < nw; ++k) v = (c < 3) ? fmin(v, extp[c][k]) : fmax(v, extp[c][k]);
        ext[c] = v;  // [6]: largest squared distance from the reference point, for now
    }
    // the atoms of this block's work item (32 i atoms, then its j range), element-
    // sorted order, float32-rounded in FP32 mode: both pair passes read these
    if (has_item) {
        const WorkItem wi = q.fq.items[blockIdx.x];
        const int cnt = TILE_I + (wi.jend - wi.jbegin);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int g = k < TILE_I ? wi.itile * TILE_I + k : wi.jbegin + (k - TILE_I);
            const int o = q.fq.orig[g];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const double c = o >= 0 ? pall[w * q.n + o] : 0.0;
                ps[w * pl + k] = q.round_f32 ? (double)(float)c : c;
            }
            ps[3 * pl + k] = o >= 0 ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    if (want_ext && threadIdx.x == 0) {
        const double ex = ext[3] - ext[0], ey = ext[4] - ext[1], ez = ext[5] - ext[2];
        const double half_diag = 0.5 * sqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
        ext[6] = cref[3] != 0.0 ? fmin(half_diag, sqrt(ext[6])) : half_diag;
        ext[7] = half_diag;  // (independent of the reference point)
        cref[0] = 0.5 * (ext[0] + ext[3]);
        cref[1] = 0.5 * (ext[1] + ext[4]);
        cref[2] = 0.5 * (ext[2] + ext[5]);
        cref[3] = 1.0;  // (read again at the next step's staging, barriers from here)
    }
    fused_stamp(q, 2, cs == scs);

    // ---- phase 1: F(Q) pass -------------------------------------------------------
    // Through the radial pair histogram when the structure's bounding box fits the
    // shared-memory histogram (block-uniform AND grid-uniform: every block has the
    // same extent), else the direct pass.  The CHOICE follows the box diagonal alone:
    // ext[6] also depends on the reference point, i.e. on the previous evaluation,
    // and the two passes round differently -- the same positions must take the same
    // pass whatever came before.  The number of nodes in use (Kh) may follow ext[6]:
    // it bounds every pair distance, and unused nodes hold zeros.
    __syncthreads();  // ext[6], ext[7] (thread 0, above)
    const double hgrid = q.tab_h;
    const int Kh = (int)fmin(2e9, ceil(2.0000002 * ext[6] / hgrid) + (double)(FT_PTS + 2));
    const bool hist = q.fq_C != nullptr &&
                      fmin(2e9, ceil(2.0000002 * ext[7] / hgrid)) + (double)(FT_PTS + 2 + 2 * FT_PAD) <=
                          (double)q.fq_cstride;
    if (hist) {
        const int Kp = Kh + 2 * FT_PAD;
        int *hi_s = reinterpret_cast<int *>(smem_raw);
        unsigned *lo_s = reinterpret_cast<unsigned *>(hi_s + q.fq_cstride);
        for (int e = threadIdx.x; e < Kp; e += blockDim.x) { hi_s[e] = 0; lo_s[e] = 0u; }
        __syncthreads();
        if (has_item) {
            const WorkItem wi = q.fq.items[blockIdx.x];
            const bool diag = (wi.info & ITEM_DIAG) != 0;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
            const int len = wi.jend - wi.jbegin;
            const double xi = ps[lane], yi = ps[pl + lane], zi = ps[2 * pl + lane];
            const bool vi = ps[3 * pl + lane] != 0.0;
            const double inv_h = 1.0 / hgrid;
            for (int jj = warp; jj < len; jj += nw) {
                const int sj = TILE_I + jj;
                // a diagonal item holds both orders of its pairs: count i > j
                if (vi && ps[3 * pl + sj] != 0.0 && (!diag || jj < lane)) {
                    const double dx = ps[sj] - xi, dy = ps[pl + sj] - yi, dz = ps[2 * pl + sj] - zi;
                    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                    if (r2 > 0.0) {
                        double y = (double)rsqrtf((float)r2);
                        y = y * fma(-0.5 * r2, y * y, 1.5);
                        y = y * fma(-0.5 * r2, y * y, 1.5);  // second Newton step: 1e-15
                        const double tpos = r2 * y * inv_h;
                        const int k = (int)tpos;
                        fq_hist_spread<FT_PTS>(hi_s, lo_s, k, (float)(tpos - (double)k));
                    }
                }
            }
            __syncthreads();
            const int ta = q.fq.tile_type[wi.itile], tb = wi.info & 0xffff;
            const int pair = ta >= tb ? ta * (ta + 1) / 2 + tb : tb * (tb + 1) / 2 + ta;
            unsigned long long *C = q.fq_C + (size_t)pair * q.fq_cstride;
            for (int e = threadIdx.x; e < Kp; e += blockDim.x) {
                const long long v = (long long)hi_s[e] * 65536ll + (long long)lo_s[e];
                if (v != 0) atomicAdd(C + e, (unsigned long long)v);
            }
        }
        fused_stamp(q, 3, cs == scs);
        grid.sync();
        // Every block: chunks of FUSED_HCH nodes dealt round-robin, all Q bins (thread
        // = bin, three-term recurrence over a chunk's nodes from exact seeds), partial
        // sums into Sfix.  The chunk grid is FIXED (not derived from the number of
        // nodes in use, which follows the extent's reference point and so the history
        // of evaluations): a chunk past the occupied range holds zeros and adds
        // nothing, so the pair sums are bit-identical whatever K was.
        {
            constexpr int FUSED_HCH = 16;
            const int ntp = q.ntypes * (q.ntypes + 1) / 2;
            double *cs_s = reinterpret_cast<double *>(smem_raw);  // [ntp][FUSED_HCH] C / |r|
            const int m = threadIdx.x;
            const double Q = q.fq.qbin * (double)m;
            const double turn = Q * hgrid * 0.15915494309189533577;
            double sth, cth;
            sincospi(2.0 * (turn - rint(turn)), &sth, &cth);
            const double tc = cth + cth;
            const float *ftab = reinterpret_cast<const float *>(q.fq.ftab);
            double tot = 0.0;
            bool added = false;
            for (int e0 = blockIdx.x * FUSED_HCH; e0 < Kp; e0 += gridDim.x * FUSED_HCH) {
                const int ne = min(FUSED_HCH, Kp - e0);
                // (seeds first: they do not depend on the histogram's loads)
                double s0, c0;
                const double t0 = turn * (double)(e0 - FT_PAD);
                sincospi(2.0 * (t0 - rint(t0)), &s0, &c0);
                bool any = false;
                __syncthreads();  // cs_s of the previous chunk is consumed
                for (int idx = threadIdx.x; idx < ntp * FUSED_HCH; idx += blockDim.x) {
                    const int pr = idx / FUSED_HCH, e = idx - pr * FUSED_HCH;
                    double v = 0.0;
                    if (e < ne) {
                        unsigned long long *c = q.fq_C + (size_t)pr * q.fq_cstride + e0 + e;
                        const long long iv = (long long)__ldcg(c);
                        if (iv != 0) *c = 0ull;  // clear for the next pass
                        const double r = fabs((double)(e0 + e - FT_PAD)) * hgrid;
                        v = (double)iv * (1.0 / 268435456.0) * (r > 0.0 ? 1.0 / r : 1.0);
                        any |= iv != 0;
                    }
                    cs_s[idx] = v;
                }
                if (__syncthreads_or(any) == 0) continue;  // an empty chunk (block-uniform)
                if (m < nq) {
                    const double sp0 = fma(s0, cth, -(c0 * sth));
                    const int ezero = FT_PAD - e0;  // index of the r = 0 node in this chunk (if any)
                    int a = 0, b = 0;
                    for (int pr = 0; pr < ntp; ++pr) {
                        double acc = 0.0, sn0 = s0, sp = sp0;
#pragma unroll
                        for (int e = 0; e < FUSED_HCH; ++e) {
                            const double sv = (e < ezero) ? -sn0 : sn0;  // sin(Q |r|)
                            acc = fma(cs_s[pr * FUSED_HCH + e], e == ezero ? Q : sv, acc);
                            const double sn = fma(tc, sn0, -sp);
                            sp = sn0;
                            sn0 = sn;
                        }
                        tot = fma(acc, (double)ftab[(size_t)a * qp + m] * (double)ftab[(size_t)b * qp + m], tot);
                        if (++b > a) { ++a; b = 0; }  // pair = a (a + 1) / 2 + b, a >= b
                    }
                    added = true;
                }
            }
            if (added) fix_add2(q.fq.Sfix + m, q.fq.Sfix + qp + m, tot, q.fq.fix_scale);
        }
    } else {
        if (has_item)
            debye2_body<32, MODE_FQ, 8, CHEB, 1, true>(q.fq, smem_raw, (int)blockIdx.x, 0, ps, pl);
        fused_stamp(q, 3, cs == scs);
    }
    grid.sync();
    fused_stamp(q, 4, cs == scs);

    // ---- phase 2: F and this block's rows of M F ------------------------------------
    double *Fs = reinterpret_cast<double *>(smem_raw);  // [qp]
    double *Ms = Fs + qp;                               // [qp] M F, then the weights
    double *red = Ms + qp;                              // [32]
    double *wab = red + 32;                             // [ntypes^2][qp] (table force pass)
    for (int m = threadIdx.x; m < qp; m += blockDim.x) {
        const double sm = m < nq ? ((double)(long long)__ldcg(q.fq.Sfix + m) +
                                    (double)(long long)__ldcg(q.fq.Sfix + qp + m) * (1.0 / FIX_LOW)) *
                                       q.s_scale_inv
                                 : 0.0;
        Fs[m] = 2.0 * sm * q.inv_na_d[m];
        if (blockIdx.x == 0) q.fq.S[m] = sm;  // the handle's pair sums stay current
    }
    __syncthreads();
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int row = blockIdx.x + warp * gridDim.x; row < nq; row += nw * gridDim.x) {
            const double *mrow = q.Mq + (size_t)row * qp;
            double acc = 0.0;
            for (int m = lane; m < nq; m += 32) acc = fma(mrow[m], Fs[m], acc);
            acc = warp_sum(acc);
            if (lane == 0) q.MF[row] = acc;
        }
    }
    fused_stamp(q, 5, cs == scs);
    grid.sync();
    fused_stamp(q, 6, cs == scs);

    // ---- phase 3: potential + weights (every block), then the force pass ------------
    // (every block has read the F(Q) accumulators: clear them for the next pass)
    if (blockIdx.x == 0)
        for (int m = threadIdx.x; m < 2 * qp; m += blockDim.x) q.fq.Sfix[m] = 0ull;
    double la = 0.0, lb = 0.0;
    for (int m = threadIdx.x; m < nq; m += blockDim.x) {
        const double mf = __ldcg(q.MF + m);
        Ms[m] = mf;
        la = fma(Fs[m], q.vgo[m], la);
        lb = fma(Fs[m], mf, lb);
    }
    const double a = block_sum(la, red);
    const double b = block_sum(lb, red);
    const double c = q.gogo;
    const double scale_true = b > 0.0 ? a / b : 0.0;
    const bool pos = scale_true > 0.0;
    const double scale = pos ? scale_true : 1.0;
    double dd = c - 2.0 * scale * a + scale * scale * b;  // |go - scale gc|^2
    if (dd < 0.0) dd = 0.0;
    const double gd = a - scale * b;                      // gc . (go - scale gc)
    double value, pref;
    if (q.potential == 0) {  // Rw
        value = pos ? sqrt(dd / c) : 1.0;
        pref = dd > 0.0 ? -value / dd : 0.0;
    } else {  // chi^2
        value = dd;
        pref = -2.0;
    }
    const double gdb = b > 0.0 ? gd / b : 0.0;
    const double coef0 = pref * (scale + gdb);
    const double coef1 = pref * (scale * scale + 2.0 * gdb * scale_true);
    double *wq = q.wq_blk + (size_t)blockIdx.x * qp;
    const bool table = q.phi_tab != nullptr;
    double lw = 0.0;
    for (int m = threadIdx.x; m < qp; m += blockDim.x) {
        const double w = m < nq ? q.conv * (coef0 * q.vgo[m] - coef1 * Ms[m]) : 0.0;
        if (table) Ms[m] = w;  // (each thread reads and writes its own elements)
        else wq[m] = w;
        lw = fma(fabs(w), q.gforce[m], lw);
    }
    const double fscale = force_fix_scale((double)q.n * block_sum(lw, red));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        q.out4[0] = value * q.conv;
        q.out4[1] = scale;
        q.out4[2] = value;
        q.out4[3] = scale_true;
        q.out4[4] = 0.0;  // restraint energy, summed by the spring kernels that follow
    }
    __syncthreads();  // the weights are complete (block scope) and `red` is free again
    fused_stamp(q, 7, cs == scs);
    if (table) {
        // weights of each element pair, the table over r = 0 .. 2 x radius, barrier,
        // one interpolation per pair of the block's work item
        const int nt = q.ntypes, ntp = nt * nt;
        const float *ftab = reinterpret_cast<const float *>(q.fq.ftab);
        const float *inv_na = reinterpret_cast<const float *>(q.fq.inv_na);
        for (int idx = threadIdx.x; idx < ntp * qp; idx += blockDim.x) {
            const int pr = idx / qp, m = idx - pr * qp;
            wab[idx] = Ms[m] * (double)ftab[(size_t)(pr / nt) * qp + m] *
                       (double)ftab[(size_t)(pr % nt) * qp + m] * (double)inv_na[m];
        }
        __syncthreads();
        const double h = q.tab_h;
        const int K = (int)fmin((double)q.phi_cap, ceil(2.0000002 * ext[6] / h) + (double)(FT_PTS + 2));
        if (q.stamps && blockIdx.x == 0 && threadIdx.x == 0) {  // developer timing
            q.ext_ref[4] = ext[6];
            q.ext_ref[5] = (double)K;
        }
        double *tab = q.phi_tab + (size_t)cs * ntp * q.phi_stride;
        fused_table_build(tab, q.phi_stride, ntp, K, h, wab, nq, qp, q.fq.qbin);
        fused_stamp(q, 16, cs == scs);
        fused_stamp_max(q, 18, cs == scs);
        grid.sync();
        fused_stamp(q, 17, cs == scs);
        if (has_item) {
            const WorkItem wi = q.fq.items[blockIdx.x];
            const int pr = q.fq.tile_type[wi.itile] * nt + (wi.info & 0xffff);
            fused_table_forces(q, wi, ps, pl, tab + (size_t)pr * q.phi_stride + FT_PAD, K,
                               1.0 / h, wab + (size_t)pr * qp, fscale, wab + (size_t)ntp * qp);
        }
    } else if (has_item) {
        DebyeParams fo = q.fo;
        fo.wq = wq;
        fo.fix_scale = fscale;
        debye2_body<32, MODE_FORCE, 8, CHEB, 1, true>(fo, smem_raw, (int)blockIdx.x, 0, ps, pl);
    }
    fused_stamp(q, 8, cs == scs);
    fused_stamp_max(q, 19, cs == scs);
    // ---- phase 4: forces to float64; plain evaluation: results straight into the
    // caller's pinned buffer; leapfrog: second half kick into the destination state
    // and its host mirror ----------------------------------------------------------
    const bool finish = q.lf_mirror != nullptr;
    double *mirror = finish ? q.lf_mirror + (size_t)cs * q.chain_stride : nullptr;
    __shared__ double lf_shift[3];
    if (finish && threadIdx.x < 3) lf_shift[threadIdx.x] = lf_shift_of(ctl, ext, threadIdx.x);
    grid.sync();  // (also a block barrier: lf_shift is visible)
    fused_stamp(q, 9, cs == scs);
    {
        const double inv = 1.0 / fscale;  // a power of two
        for (int e = gtid; e < 3 * q.n; e += gsz) {
            const double f = (double)(long long)__ldcg(q.fo.Ffix + e) * inv;
            q.fo.Ffix[e] = 0ull;  // clear for the next pass
            q.fo.force[e] = f;
            if (q.force_out) q.force_out[e] = f;
            if (finish) {
                double mx, mp;
                lf_kick(ctl, q.slab, q.n, q.pos, lf_shift, e, e % 3, f, mx, mp);
                mirror[e] = mx;
                mirror[3 * (size_t)q.n + e] = mp;
            }
        }
        if (q.force_out && gtid < 5) q.out_host[gtid] = __ldcg(q.out4 + gtid);
        if (finish) {
            // scalars of the step: energy, scale, value, scale_true, restraint energy,
            // (kinetic energy: summed by the host from the mirrored momenta), shift
            double *out = mirror + 6 * (size_t)q.n;
            if (blockIdx.x == 0 && threadIdx.x < 9) {
                const int k = threadIdx.x;
                out[k] = k < 5 ? __ldcg(q.out4 + k) : (k == 5 ? 0.0 : lf_shift[k - 6]);
            }
            fused_stamp(q, 10, cs == scs);
            if (cs + 1 < q.n_chain) {
                grid.sync();  // the next step's staging reads the new state
                // Every block's share of the mirror was written before the barrier:
                // tell the host that step cs is complete (iid_leapfrog_chain_next
                // polls this word; the last step's completion is the end of the
                // launch).  The system-wide fence costs its thread ~1 us: the LAST
                // block raises the flag, whose work item is the shortest (or none).
                if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
                    __threadfence_system();
                    *reinterpret_cast<volatile double *>(out + 15) = ctl[7];
                }
            }
            fused_stamp(q, 11, cs == scs);
        }
    }
  }
  // (cref was last written by thread 0, barriers ago)
  if (blockIdx.x == 0 && threadIdx.x < 4 && (q.phi_tab != nullptr || q.lf_mirror != nullptr))
      q.ext_ref[threadIdx.x] = cref[threadIdx.x];
}

}  // namespace iid
