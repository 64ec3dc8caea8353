// iid_fused.cuh -- one cooperative kernel for a whole Rw / chi^2 energy + force
// evaluation of a SMALL structure (the sampler's regime: Au561, one work item
// per SM), sm_100a.
//
// pyiid/sim/__init__.py:10-38 (leapfrog) + calc/calc_1d.py:78-95 cost the
// reference 1 x grad PDF + 2 x PDF per step.  Round 1 ran the evaluation as a
// graph of 6-7 small launches (staging, F(Q) pass, G(r), potential, weights,
// force pass) that at 561 atoms spent more time in launch gaps and in the
// one-block float64 stages than in the two pair sums.  Here the sequence is ONE
// launch with grid-wide barriers between its phases:
//
//   0  staging: (leapfrog: half kick + drift) element sort, float32 rounding,
//      clearing of the S and force accumulators
//   1  F(Q) pass, one work item per block (debye2_body<MODE_FQ>)
//   2  F = 2 S / na;  (M F)[m] for this block's rows of M = T^T T
//   3  every block, redundantly: Rw / chi^2, scale and the chain-rule weights
//      in Q SPACE -- with gc = T F:  gc.go = F.(T^T go),  gc.gc = F.(M F), so
//      neither G(r) nor the R x Q matrix is touched -- then the force pass of
//      its work item (debye2_body<MODE_FORCE>)
//   4  forces complete (deterministic mode: the items' partial forces added in
//      item order); leapfrog: second half kick + centring into the destination
//      state and its host mirror, spread over the grid
//
// The Q-space scalars: a = F.vgo, b = F.MF, c = go.go (once per target);
// scale s = a/b (<= 0: the reference's branches, master_kernel.py:229-230,
// 263-264), |go - s gc|^2 = c - 2 s a + s^2 b, gc.(go - s gc) = a - s b.
#pragma once
#include <cooperative_groups.h>
#include "iid_debye2.cuh"
#include "iid_sampler.cuh"

namespace iid {

#ifdef IID_EXP_NOLDCG  // developer A/B only (wrong for chains)
constexpr bool FUSED_LDCG = false;
#else
constexpr bool FUSED_LDCG = true;
#endif

struct FusedParams {
    DebyeParams fq;  // F(Q) pass (S = the handle's accumulator)
    DebyeParams fo;  // force pass (wq is set per block)
    int n_items;
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

template <bool CHEB>
__global__ void __launch_bounds__(384, 1) fused_eval_kernel(const FusedParams q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int nq = q.fq.nq, qp = q.fq.qp;

  // the step parameters of the whole chain: ONE round trip to the pinned staging
  // per block and launch
  __shared__ double ctl_all[LF_CHAIN_MAX * LF_CTL];
  if (q.lf) {
      for (int k = threadIdx.x; k < q.n_chain * LF_CTL; k += blockDim.x)
          ctl_all[k] = q.ctl[(size_t)(k / LF_CTL) * q.chain_stride + k % LF_CTL];
      __syncthreads();
  }
  for (int cs = 0; cs < q.n_chain; ++cs) {
    // a chain of leapfrog steps: step cs mirrors its state to ring slot cs of the
    // pinned staging.  Everything another block wrote in an EARLIER step of this
    // launch is read past L1 (ld.cg).
    const double *ctl = ctl_all + cs * LF_CTL;
    fused_stamp(q, 0, cs == 0);
    // ---- phase 0: staging ---------------------------------------------------------
    {
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
        for (int e = gtid; e < qp; e += gsz) q.fq.S[e] = 0.0;
        for (int e = gtid; e < 3 * q.n; e += gsz) q.fo.force[e] = 0.0;
        for (int k = gtid; k < q.np; k += gsz) {
            const int o = q.fq.orig[k];
            double c[3] = {0.0, 0.0, 0.0};
            if (o >= 0) {
                if (q.lf) {
                    // p_half = p + (step/2) f; q' = q + step p_half / m (numpy's
                    // operation order, see lf_stage_kernel)
                    const double step = ctl[0];
                    const int src = (int)ctl[1], dst = (int)ctl[2];
                    const double *qq = lf_slot(q.slab, q.n, src, 0),
                                 *pp = lf_slot(q.slab, q.n, src, 1),
                                 *ff = lf_slot(q.slab, q.n, src, 2);
                    double *pd = lf_slot(q.slab, q.n, dst, 1);
                    const double half = __dmul_rn(0.5, step), m = q.mass[o];
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const size_t e = 3 * (size_t)o + w;
                        const double ph = __dadd_rn(__ldcg(pp + e), __dmul_rn(half, __ldcg(ff + e)));
                        pd[e] = ph;
                        const double qn = __dadd_rn(__ldcg(qq + e), __dmul_rn(step, __ddiv_rn(ph, m)));
                        q.pos[e] = qn;
                        c[w] = qn;
                    }
                } else if (q.pos_in) {
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        c[w] = q.pos_in[3 * (size_t)o + w];
                        q.pos[3 * (size_t)o + w] = c[w];  // device copy for the kernels that follow
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < 3; ++w) c[w] = q.pos[3 * (size_t)o + w];
                }
                if (q.round_f32)
#pragma unroll
                    for (int w = 0; w < 3; ++w) c[w] = (double)(float)c[w];
            }
            const_cast<double *>(q.fq.x)[k] = c[0];
            const_cast<double *>(q.fq.y)[k] = c[1];
            const_cast<double *>(q.fq.z)[k] = c[2];
            const_cast<float *>(q.fq.valid)[k] = o >= 0 ? 1.f : 0.f;
        }
    }
    fused_stamp(q, 1, cs == 0);
    grid.sync();
    fused_stamp(q, 2, cs == 0);

    // ---- phase 1: F(Q) pass -------------------------------------------------------
    if ((int)blockIdx.x < q.n_items)
        debye2_body<32, MODE_FQ, 8, CHEB, 1, FUSED_LDCG>(q.fq, smem_raw, (int)blockIdx.x, 0);
    fused_stamp(q, 3, cs == 0);
    grid.sync();
    if (q.fq.Sitem != nullptr) {
        // deterministic F(Q): the items' partial sums added in item order, one
        // warp per Q bin (lanes over items, butterfly sum)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int m = blockIdx.x * nw + warp; m < nq; m += gridDim.x * nw) {
            double acc = 0.0;
            for (int it = lane; it < q.n_items; it += 32) acc += __ldcg(q.fq.Sitem + (size_t)it * qp + m);
            acc = warp_sum(acc);
            if (lane == 0) q.fq.S[m] = acc;
        }
        grid.sync();
    }
    fused_stamp(q, 4, cs == 0);

    // ---- phase 2: F and this block's rows of M F ------------------------------------
    double *Fs = reinterpret_cast<double *>(smem_raw);  // [qp]
    double *Ms = Fs + qp;                               // [qp] M F
    double *red = Ms + qp;                              // [32]
    for (int m = threadIdx.x; m < qp; m += blockDim.x)
        Fs[m] = m < nq ? 2.0 * __ldcg(q.fq.S + m) * q.inv_na_d[m] : 0.0;
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
    fused_stamp(q, 5, cs == 0);
    grid.sync();
    fused_stamp(q, 6, cs == 0);

    // ---- phase 3: potential + weights (every block), then the force pass ------------
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
    for (int m = threadIdx.x; m < qp; m += blockDim.x)
        wq[m] = m < nq ? q.conv * (coef0 * q.vgo[m] - coef1 * Ms[m]) : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        q.out4[0] = value * q.conv;
        q.out4[1] = scale;
        q.out4[2] = value;
        q.out4[3] = scale_true;
        q.out4[4] = 0.0;  // restraint energy, summed by the spring kernels that follow
    }
    __syncthreads();  // wq is complete (block scope) and the scratch is free again
    fused_stamp(q, 7, cs == 0);
    if ((int)blockIdx.x < q.n_items) {
        DebyeParams fo = q.fo;
        fo.wq = wq;
        debye2_body<32, MODE_FORCE, 8, CHEB, 1, FUSED_LDCG>(fo, smem_raw, (int)blockIdx.x, 0);
    }
    fused_stamp(q, 8, cs == 0);
    // ---- phase 4: the forces are complete after one more barrier.  Deterministic
    // mode adds the items' partial forces in item order; the results go straight
    // into the caller's pinned buffer (plain evaluation) or through the leapfrog's
    // second half kick into the destination state and its host mirror ------------
    const bool finish = q.lf_mirror != nullptr;
    double *mirror = finish ? q.lf_mirror + (size_t)cs * q.chain_stride : nullptr;
    __shared__ double lf_shift[3];
    if (finish) lf_shift_block(ctl, q.pos, q.n, lf_shift);  // while the slowest block finishes
    if (q.fo.Fi != nullptr) {
        grid.sync();
        fused_stamp(q, 9, cs == 0);
        // one warp per atom, lanes over the items (each lane adds its items in item
        // order, then a butterfly sum: the same order on every run)
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int g = blockIdx.x * nw + warp; g < q.np; g += gridDim.x * nw) {
            const int o = q.fq.orig[g];
            if (o < 0) continue;  // warp-uniform
            const int tile = g / TILE_I, a = g - tile * TILE_I;
            double fx = 0.0, fy = 0.0, fz = 0.0;
            for (int it = lane; it < q.n_items; it += 32) {
                const WorkItem wi = q.fq.items[it];
                if (wi.itile == tile) {
                    const double *fi = q.fo.Fi + ((size_t)it * 32 + a) * 3;
                    fx += __ldcg(fi);
                    fy += __ldcg(fi + 1);
                    fz += __ldcg(fi + 2);
                }
                if (!(wi.info & ITEM_DIAG) && g >= wi.jbegin && g < wi.jend) {
                    const double *fj = q.fo.Fj + ((size_t)it * q.fo.fj_len + (g - wi.jbegin)) * 3;
                    fx += __ldcg(fj);
                    fy += __ldcg(fj + 1);
                    fz += __ldcg(fj + 2);
                }
            }
            fx = warp_sum(fx);
            fy = warp_sum(fy);
            fz = warp_sum(fz);
            if (lane < 3) {
                const double f = lane == 0 ? fx : (lane == 1 ? fy : fz);
                const size_t e = (size_t)o * 3 + lane;
                q.fo.force[e] = f;
                if (q.force_out) q.force_out[e] = f;
                if (finish) lf_kick(ctl, q.slab, q.n, q.pos, mirror, lf_shift, e, lane, f);
            }
        }
        if (q.force_out && gtid < 5) q.out_host[gtid] = __ldcg(q.out4 + gtid);
    } else if (q.force_out || finish) {
        grid.sync();
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
        for (int e = gtid; e < 3 * q.n; e += gsz) {
            const double f = __ldcg(q.fo.force + e);
            if (q.force_out) q.force_out[e] = f;
            if (finish) lf_kick(ctl, q.slab, q.n, q.pos, mirror, lf_shift, e, e % 3, f);
        }
        if (q.force_out && gtid < 5) q.out_host[gtid] = __ldcg(q.out4 + gtid);
    }
    if (finish) {
        // scalars of the step: energy, scale, value, scale_true, restraint energy,
        // (kinetic energy: summed by the host from the mirrored momenta), shift
        double *out = mirror + 6 * (size_t)q.n;
        if (blockIdx.x == 0 && threadIdx.x < 9) {
            const int k = threadIdx.x;
            out[k] = k < 5 ? __ldcg(q.out4 + k) : (k == 5 ? 0.0 : lf_shift[k - 6]);
        }
        // the next step's staging reads the state written here and clears the
        // accumulators read here
        fused_stamp(q, 10, cs == 0);
        if (cs + 1 < q.n_chain) grid.sync();
        fused_stamp(q, 11, cs == 0);
    }
  }
}

}  // namespace iid
