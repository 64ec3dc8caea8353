// iid_debye2.cuh -- FP32 Debye-sum kernel, producer/consumer version (sm_100a).
//
// Same mapping and arithmetic as iid_debye.cuh (block = one (i-tile, j-slab)
// work item, warp = Q chunk of C bins, lane = atom i, accumulators in
// registers, FP32 rotation recurrence), but the per-pair set-up is no longer
// repeated by every warp: for each tile of TJ2 j atoms the block computes the
// 32 x TJ2 pair records ONCE, cooperatively, into shared memory
//   A  = (cos theta, sin theta, kappa, r^2)        theta = qbin r, kappa = qbin r
//   B  = (dx, dy, dz, -)                           q_j - q_i
//   SC = (sin, cos)(m0_w theta) / r^3 for every warp w of the block
// (distance, phase in turns and all range reductions in float64, sincos by a
// float32 minimax polynomial on the quadrant-reduced angle, the chunk seeds by
// rotating with e^{i C theta}), double-buffered so that the records of tile
// t+1 are produced while tile t is consumed.  The consumer's per-(pair, chunk)
// overhead is a handful of shared-memory loads; its bin loop is pure FP32 pipe
// work on packed (two-bin) instructions: 8.25 per bin step for F + grad F, 6.7
// for the gradient-only items of the square walk, ~2.6 for F only, ~4.3 for
// the force.
//
// Packed arithmetic.  With one atom pair per lane every FFMA operand is a
// per-lane register, and on sm_100 a scalar FFMA whose two multiplicands sit
// in the same register bank costs an extra dispatch cycle (measured: the scalar
// loop reaches 73 % of the FP32 pipe, scripts/ubench2.cu).  The loop is
// therefore written on float2 values with the sm_100 packed instructions
// (FFMA2/FMUL2/FADD2): a warp advances TWO half-chunks of C/2 bins per
// instruction (bins m0+k and m0+C/2+k, two independent rotation recurrences
// with their own seeds), the accumulators live in aligned register pairs and
// the per-pair constants enter as broadcast scalars; this reaches 87 % of the
// pipe in the same micro-benchmark.
#pragma once
#include <type_traits>
#include "iid_debye.cuh"

namespace iid {

constexpr int NREC = 10;  // floats per pair record shared by all warps

// TJ2 = j atoms per produced tile (template parameter of the kernel)
__host__ __device__ inline size_t debye2_buf_bytes(int nwarp, int tj)
{
    return (size_t)tj * 32 * sizeof(float) * (NREC + 4 * (size_t)nwarp);
}
// MODE_FORCE keeps the per-(pair, warp) scalars of one tile for the j-side sum
__host__ __device__ inline size_t debye2_phi_bytes(int nwarp, int tj)
{
    return (size_t)tj * 32 * sizeof(float) * (size_t)nwarp;
}

// sin, cos of 2 pi u for any u >= 0 held in float64: quadrant in float64,
// polynomial in float32 (error ~1e-7).
__device__ __forceinline__ void sincos_turns(double u, float &s, float &c)
{
    const double q4 = rint(u * 4.0);
    const float fr = (float)fma(-0.25, q4, u);
    const int quad = ((int)q4) & 3;
    float sf, cf;
    sincos_eighth(fr, sf, cf);
    s = (quad & 1) ? cf : sf;
    c = (quad & 1) ? sf : cf;
    if (quad == 1 || quad == 2) c = -c;
    if (quad >= 2) s = -s;
}

// CHEB selects how sin/cos are advanced from bin to bin: false = complex
// rotation everywhere (2 FMUL + 2 FFMA per bin, error ~1.4e-7 rms); true = the
// three-term recurrence x[k+1] = 2 cos(theta) x[k] - x[k-1] (1 FFMA per
// sequence per bin) inside quarter-chunks of C/4 bins.  The float32 rounding
// of 2 cos(theta) is a small frequency error that grows linearly with the
// number of steps, so every quarter-chunk restarts from an exact seed (the
// second one is the half-chunk seed rotated by C/4 bins) and takes its first
// step with the rotation; over 8 bins the error is 1.3-2.5e-7 rms, the same
// as the rotation ("recurrences").
// The kernel body, callable from the stand-alone kernel below and from the
// fused evaluation kernel (iid_fused.cuh): block (bx, by) of a launch with
// dynamic shared memory `smem_raw`.
// PS: the positions of the item's atoms are staged in shared memory by the
// caller (the fused evaluation kernel): ps[c * pl + k], c = x, y, z, validity;
// k = lane for the i atoms, 32 + (j - jbegin) for the j atoms.
template <int C, int MODE, int TJ2, bool CHEB, int PU = 1, bool PS = false>
__device__ __forceinline__ void debye2_body(const DebyeParams &p, unsigned char *smem_raw,
                                            const int bx, const int by,
                                            const double *ps = nullptr, const int pl = 0)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int chunk0 = by * nwarp;
    const int m0 = (chunk0 + warp) * C;
    const bool active = m0 < p.nq;

    // MODE_GRAD: a row job (i-tile, consecutive j segments); the other modes: one
    // work item = one segment
    int itile, seg = 0, seg_end = 1, dest = -1;
    RowSeg sg;
    if constexpr (MODE == MODE_GRAD) {
        const RowJob job = p.jobs[bx];
        itile = job.itile;
        seg = job.seg_begin;
        seg_end = job.seg_end;
        dest = job.dest;
        sg = p.segs[seg];
    } else {
        const WorkItem it = p.items[p.item_begin + (int64_t)bx * p.item_stride];
        itile = it.itile;
        sg.jbegin = it.jbegin;
        sg.jend = it.jend;
        sg.info = it.info;
    }
    // per segment: diagonal tile (both orders present: F(Q) weight 1/2, applied
    // by the producer to the r^2 of the record), gradient-only range (above the
    // diagonal, or every segment when no F(Q) is wanted), element type of the j run
    bool diag = (sg.info & ITEM_DIAG) != 0;
    bool nof = MODE == MODE_GRAD && p.grad_split && ((sg.info & ITEM_NOF) != 0 || p.S == nullptr);
    float fw = (MODE == MODE_GRAD && (diag || !p.grad_split)) ? 0.5f : 1.f;
    const int atype = p.tile_type[itile];
    const int gi = itile * TILE_I + lane;
    const double xi = PS ? ps[lane] : p.x[gi], yi = PS ? ps[pl + lane] : p.y[gi],
                 zi = PS ? ps[2 * pl + lane] : p.z[gi];
    const bool vi = PS ? ps[3 * pl + lane] != 0.0 : p.valid[gi] != 0.f;

    const float *ftab = reinterpret_cast<const float *>(p.ftab);
    const float *fa = ftab + (size_t)atype * p.qp;
    const float *fb = ftab + (size_t)(sg.info & 0xffff) * p.qp;
    const float *inv_na = reinterpret_cast<const float *>(p.inv_na);

    constexpr int NPAIR = TJ2 * 32;
    const size_t buf_bytes = debye2_buf_bytes(nwarp, TJ2);
    float *phis = reinterpret_cast<float *>(smem_raw + 2 * buf_bytes);  // MODE_FORCE
    const int nactive = min(nwarp, (p.nq - chunk0 * C + C - 1) / C);
    // structure of arrays, [field][jj][lane]: scalar, conflict-free loads that
    // leave the register allocator free to keep the FFMA operand pairs in
    // opposite register banks (vector loads would pin their parities)
    auto tab = [&](int b) { return reinterpret_cast<float *>(smem_raw + (size_t)b * buf_bytes); };

    static_assert(C % 2 == 0, "C must be even");
    constexpr int H = C / 2;  // bins per half-chunk; .x = bin m0+k, .y = bin m0+H+k
    float2 accF[MODE != MODE_FORCE ? H : 1];
    float2 accX[MODE == MODE_GRAD ? H : 1], accY[MODE == MODE_GRAD ? H : 1],
        accZ[MODE == MODE_GRAD ? H : 1];
    float2 w0[MODE == MODE_FORCE ? H : 1], w1[MODE == MODE_FORCE ? H : 1];
    float fix = 0.f, fiy = 0.f, fiz = 0.f;
    // MODE_FORCE weights of this warp's bins: lane L forms the weight of bin
    // m0 + L once (coalesced loads), the others take it by shuffle (128 broadcast
    // loads per thread in front of a 40-atom item were a visible share of it)
    float wlane = 0.f;
    if constexpr (MODE == MODE_FORCE) {
        const int bin = m0 + lane;
        if (lane < C && bin < p.nq) wlane = (float)(p.wq[bin]) * fa[bin] * fb[bin] * inv_na[bin];
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
        if constexpr (MODE != MODE_FORCE) accF[k] = make_float2(0.f, 0.f);
        if constexpr (MODE == MODE_GRAD) {
            accX[k] = make_float2(0.f, 0.f);
            accY[k] = make_float2(0.f, 0.f);
            accZ[k] = make_float2(0.f, 0.f);
        }
        if constexpr (MODE == MODE_FORCE) {
            const float wa = __shfl_sync(0xffffffffu, wlane, k);
            const float wb = __shfl_sync(0xffffffffu, wlane, H + k);
            w0[k] = make_float2(wa, wb);
            w1[k] = make_float2(wa * (float)(m0 + k), wb * (float)(m0 + H + k));
        }
    }

    // ---- producer: the pair records of one j tile -----------------------------
    // PU pairs per thread are set up side by side (PU = 2 for the 16-j tiles of
    // the gradient kernel): the set-up of one pair is a single dependent chain
    // (global load -> float64 distance -> range reductions -> polynomials ->
    // the seed rotations), and the ncu source view showed a producing warp at
    // ~0.16 instructions per cycle; two independent chains halve that time.
    // The seeds advance by ONE rotation per warp (e^{i C theta} = the square of
    // the half-chunk rotation) with the half-chunk seed off the critical path.
    auto produce = [&](int jt, int b) {
        float *T = tab(b);
        float *S = T + NREC * NPAIR;
        float4 *RA = reinterpret_cast<float4 *>(T);
        float2 *S2 = reinterpret_cast<float2 *>(S);
        for (int pr0 = threadIdx.x; pr0 < NPAIR; pr0 += PU * blockDim.x) {
            int pr[PU];
            bool live[PU];
            float s[PU], c[PU], sC[PU], cC[PU], sW[PU], cW[PU];
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                const int want = pr0 + u * blockDim.x;
                live[u] = want < NPAIR;
                pr[u] = live[u] ? want : pr0;  // a dead slot recomputes pair pr0, stores nothing
                const int jj = pr[u] >> 5;     // (pr & 31) == lane: this thread's own atom i
                const int gj = jt + jj;
                const int sj = 32 + gj - sg.jbegin;
                const double dxd = (PS ? ps[sj] : p.x[gj]) - xi, dyd = (PS ? ps[pl + sj] : p.y[gj]) - yi,
                             dzd = (PS ? ps[2 * pl + sj] : p.z[gj]) - zi;
                const bool keep = vi && (PS ? ps[3 * pl + sj] != 0.0 : p.valid[gj] != 0.f);
                const double r2 = fma(dxd, dxd, fma(dyd, dyd, dzd * dzd));
                const float r2f = (float)r2;
                double y = (double)rsqrtf(r2f);
                y = y * fma(-0.5 * r2, y * y, 1.5);  // one Newton step in float64
                if (!(keep && r2f > 0.f)) y = 0.0;   // self pair, ghost atom, r == 0
                const double r = r2 * y;
                const double ut = r * p.qbin_turns;  // turns per Q bin
                float sth, cth, sQ, cQ;
                sincos_turns(ut, sth, cth);
                sincos_turns(ut * (double)H, sC[u], cC[u]);   // rotation by one half-chunk
                sincos_turns(ut * (double)(H / 2), sQ, cQ);   // rotation by one quarter-chunk
                if (chunk0 == 0) { s[u] = 0.f; c[u] = 1.f; }
                else sincos_turns(ut * (double)(chunk0 * C), s[u], c[u]);
                const float invr = (float)y;
                const float b3 = invr * invr * invr;
                s[u] *= b3;
                c[u] *= b3;
                if (live[u]) {
                    RA[pr[u]] = make_float4(cth, sth, (float)(p.qbin * r), r2f * fw);
                    RA[NPAIR + pr[u]] = make_float4((float)dxd, (float)dyd, (float)dzd, cQ);
                    T[8 * NPAIR + pr[u]] = sQ;
                }
                if constexpr (PU > 1) {  // rotation by a whole chunk
                    cW[u] = fmaf(cC[u], cC[u], -(sC[u] * sC[u]));
                    sW[u] = 2.f * sC[u] * cC[u];
                }
            }
            for (int w = 0; w < nwarp; ++w) {
#pragma unroll
                for (int u = 0; u < PU; ++u) {
                    const float s1 = fmaf(s[u], cC[u], c[u] * sC[u]);
                    const float c1 = fmaf(c[u], cC[u], -(s[u] * sC[u]));
                    if (live[u]) {
                        S2[(2 * w) * NPAIR + pr[u]] = make_float2(s[u], s1);      // sin at m0, m0+H
                        S2[(2 * w + 1) * NPAIR + pr[u]] = make_float2(c[u], c1);  // cos at m0, m0+H
                    }
                    if constexpr (PU > 1) {
                        const float sn = fmaf(s[u], cW[u], c[u] * sW[u]);
                        c[u] = fmaf(c[u], cW[u], -(s[u] * sW[u]));
                        s[u] = sn;
                    } else {
                        s[u] = fmaf(s1, cC[u], c1 * sC[u]);
                        c[u] = fmaf(c1, cC[u], -(s1 * sC[u]));
                    }
                }
            }
        }
    };

    // ---- consumer: this warp's chunk of bins for pairs (lane, jj) --------------
    // one pair record as the consumer holds it in registers
    struct Rec {
        float4 a;   // cos(theta), sin(theta), kappa, r^2
        float4 b;   // dx, dy, dz, cos(HQ theta)
        float sq;   // sin(HQ theta)
        float2 s, c;  // seeds of this warp's two half-chunks
    };
    auto load_rec = [&](int b, int jj) {
        Rec r;
        const float *T = tab(b);
        const float4 *RA = reinterpret_cast<const float4 *>(T) + jj * 32 + lane;
        const float2 *S2 = reinterpret_cast<const float2 *>(T + NREC * NPAIR) +
                           (2 * warp) * NPAIR + jj * 32 + lane;
        r.a = RA[0];
        r.b = RA[NPAIR];
        r.sq = CHEB ? T[8 * NPAIR + jj * 32 + lane] : 0.f;
        r.s = S2[0];
        r.c = S2[NPAIR];
        return r;
    };
    auto bins = [&](auto nof_tag, const Rec &rec, int jj) {
        constexpr bool NOF = decltype(nof_tag)::value;
        const float cth = rec.a.x, sth = rec.a.y, kap = rec.a.z, r2 = rec.a.w;
        const float dx = rec.b.x, dy = rec.b.y, dz = rec.b.z, cq = rec.b.w, sq = rec.sq;
        float2 s = rec.s, c = rec.c;
        if constexpr (MODE == MODE_GRAD && NOF && CHEB) {
            // Gradient only (items above the diagonal; F(Q) is taken from the
            // items below it).  Basis change that removes the running m*kappa:
            // with C_k = kappa c_k and t_k = m_q kappa c_k - s_k (m_q = first bin
            // of the quarter-chunk) the coefficient is a_k = k C_k + t_k with a
            // COMPILE-TIME k, and both sequences obey the same three-term
            // recurrence as s and c (it is linear).  The one rotation step per
            // quarter is done on (s, c) -- the (t, C) basis is not orthogonal
            // and would amplify its rounding by m kappa -- and converted.
            // 6.7 packed instructions per bin instead of 8.25.
            constexpr int HQ = H / 2;
            const float2 cth2 = make_float2(cth, cth), sth2 = make_float2(sth, sth),
                         nsth2 = make_float2(-sth, -sth), kap2 = make_float2(kap, kap);
            const float tc = cth + cth;
            const float2 tc2 = make_float2(tc, tc);
            const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy),
                         dz2 = make_float2(dz, dz);
            float2 mkq = make_float2(kap * (float)m0, kap * (float)(m0 + H));
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float2 s0 = s, c0 = c;
                if (q == 1) {  // exact seed of the second quarter: rotate by HQ bins
                    const float2 cq2 = make_float2(cq, cq), sq2 = make_float2(sq, sq),
                                 nsq2 = make_float2(-sq, -sq);
                    s0 = __ffma2_rn(s, cq2, __fmul2_rn(c, sq2));
                    c0 = __ffma2_rn(c, cq2, __fmul2_rn(s, nsq2));
                    const float hk = kap * (float)HQ;
                    mkq = __fadd2_rn(mkq, make_float2(hk, hk));
                }
                const float2 s1 = __ffma2_rn(s0, cth2, __fmul2_rn(c0, sth2));
                const float2 c1 = __ffma2_rn(c0, cth2, __fmul2_rn(s0, nsth2));
                float2 tp = __ffma2_rn(mkq, c0, make_float2(-s0.x, -s0.y));
                float2 t = __ffma2_rn(mkq, c1, make_float2(-s1.x, -s1.y));
                float2 Cp = __fmul2_rn(kap2, c0);
                float2 Cc = __fmul2_rn(kap2, c1);
                // bin 0 of the quarter: a = t_0
                accX[q * HQ] = __ffma2_rn(tp, dx2, accX[q * HQ]);
                accY[q * HQ] = __ffma2_rn(tp, dy2, accY[q * HQ]);
                accZ[q * HQ] = __ffma2_rn(tp, dz2, accZ[q * HQ]);
#pragma unroll
                for (int kk = 1; kk < HQ; ++kk) {
                    const float kf = (float)kk;
                    const float2 a = __ffma2_rn(make_float2(kf, kf), Cc, t);
                    accX[q * HQ + kk] = __ffma2_rn(a, dx2, accX[q * HQ + kk]);
                    accY[q * HQ + kk] = __ffma2_rn(a, dy2, accY[q * HQ + kk]);
                    accZ[q * HQ + kk] = __ffma2_rn(a, dz2, accZ[q * HQ + kk]);
                    if (kk + 1 < HQ) {
                        const float2 tn = __ffma2_rn(tc2, t, make_float2(-tp.x, -tp.y));
                        const float2 Cn = __ffma2_rn(tc2, Cc, make_float2(-Cp.x, -Cp.y));
                        tp = t;
                        t = tn;
                        Cp = Cc;
                        Cc = Cn;
                    }
                }
            }
            return;
        }
            const float2 cth2 = make_float2(cth, cth), sth2 = make_float2(sth, sth),
                         nsth2 = make_float2(-sth, -sth), r22 = make_float2(r2, r2),
                         kap2 = make_float2(kap, kap);
            const float tc = cth + cth;
            const float2 tc2 = make_float2(tc, tc);
            float2 mk = make_float2(kap * (float)m0, kap * (float)(m0 + H));
            // four independent partial sums each: a single FFMA2 chain over the
            // bins would be latency-bound
            float2 p0[4], p1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { p0[q] = make_float2(0.f, 0.f); p1[q] = make_float2(0.f, 0.f); }
            float2 sp = s, cp = c;  // previous bin (three-term recurrence)
            constexpr int HQ = H / 2;  // bins per quarter-chunk
            // seed of the second quarter-chunk: the exact seed rotated by HQ bins
            float2 s8 = s, c8 = c;
            if constexpr (CHEB) {
                const float2 cq2 = make_float2(cq, cq), sq2 = make_float2(sq, sq),
                             nsq2 = make_float2(-sq, -sq);
                s8 = __ffma2_rn(s, cq2, __fmul2_rn(c, sq2));
                c8 = __ffma2_rn(c, cq2, __fmul2_rn(s, nsq2));
            }
#pragma unroll
            for (int k = 0; k < H; ++k) {
                if constexpr (MODE != MODE_FORCE) accF[k] = __ffma2_rn(s, r22, accF[k]);
                if constexpr (MODE == MODE_GRAD) {
                    const float2 a = __ffma2_rn(mk, c, make_float2(-s.x, -s.y));
                    accX[k] = __ffma2_rn(a, make_float2(dx, dx), accX[k]);
                    accY[k] = __ffma2_rn(a, make_float2(dy, dy), accY[k]);
                    accZ[k] = __ffma2_rn(a, make_float2(dz, dz), accZ[k]);
                    mk = __fadd2_rn(mk, kap2);
                }
                if constexpr (MODE == MODE_FORCE) {
                    p1[k & 3] = __ffma2_rn(w1[k], c, p1[k & 3]);
                    p0[k & 3] = __ffma2_rn(w0[k], s, p0[k & 3]);
                }
                if (k + 1 < H) {
                    if (CHEB && k + 1 == HQ) {
                        s = s8;  // fresh seed for the second quarter-chunk
                        c = c8;
                    } else if (!CHEB || k % HQ == 0) {
                        const float2 sn = __ffma2_rn(s, cth2, __fmul2_rn(c, sth2));
                        const float2 cn = __ffma2_rn(c, cth2, __fmul2_rn(s, nsth2));
                        sp = s;
                        cp = c;
                        s = sn;
                        c = cn;
                    } else {
                        const float2 sn = __ffma2_rn(tc2, s, make_float2(-sp.x, -sp.y));
                        sp = s;
                        s = sn;
                        // cos is only needed for a (MODE_GRAD), the force sum and
                        // the rotation steps; MODE_FQ skips it in the last
                        // Chebyshev steps of a quarter where nothing reads it
                        if (MODE != MODE_FQ || (k % HQ) < 1) {
                            const float2 cn = __ffma2_rn(tc2, c, make_float2(-cp.x, -cp.y));
                            cp = c;
                            c = cn;
                        }
                    }
                }
            }
            if constexpr (MODE == MODE_FORCE) {
                const float2 q1 = __fadd2_rn(__fadd2_rn(p1[0], p1[1]), __fadd2_rn(p1[2], p1[3]));
                const float2 q0 = __fadd2_rn(__fadd2_rn(p0[0], p0[1]), __fadd2_rn(p0[2], p0[3]));
                const float phi = fmaf(kap, q1.x + q1.y, -(q0.x + q0.y));
                fix = fmaf(phi, dx, fix);
                fiy = fmaf(phi, dy, fiy);
                fiz = fmaf(phi, dz, fiz);
                // the j atom's share is summed over warps and lanes once per
                // tile (reduce_j below) from this scalar
                if (!diag) phis[warp * NPAIR + jj * 32 + lane] = phi;
            }
    };
    // two pairs per trip with ping-pong register sets: the record of the next
    // pair is in flight while the bin loop of the current one runs
    auto consume = [&](auto nof_tag, int b) {
        Rec r0 = load_rec(b, 0);
#pragma unroll 1
        for (int jj = 0; jj < TJ2; jj += 2) {
            Rec r1 = load_rec(b, jj + 1);
            bins(nof_tag, r0, jj);
            r0 = load_rec(b, min(jj + 2, TJ2 - 1));
            bins(nof_tag, r1, jj + 1);
        }
    };

    // ---- MODE_FORCE: Newton's third law for the j atoms of one tile -----------
    auto reduce_j = [&](int b, int jt) {
        const float4 *RB = reinterpret_cast<const float4 *>(tab(b)) + NPAIR + lane;
        for (int jj = warp; jj < TJ2; jj += nwarp) {
            float phi = 0.f;
            for (int w = 0; w < nactive; ++w) phi += phis[w * NPAIR + jj * 32 + lane];
            const float4 d = RB[jj * 32];  // dx, dy, dz of pair (lane, jj)
            const float jx = warp_sum(-phi * d.x);
            const float jy = warp_sum(-phi * d.y);
            const float jz = warp_sum(-phi * d.z);
            const int oj = p.orig[jt + jj];
            if (p.Ffix != nullptr) {  // fixed point: bit-reproducible in any order
                if (lane == 0 && oj >= 0) {
                    fix_add(p.Ffix + (size_t)oj * 3 + 0, (double)jx, p.fix_scale);
                    fix_add(p.Ffix + (size_t)oj * 3 + 1, (double)jy, p.fix_scale);
                    fix_add(p.Ffix + (size_t)oj * 3 + 2, (double)jz, p.fix_scale);
                }
            } else if (lane == 0 && oj >= 0) {
                atomicAdd(&p.force[(size_t)oj * 3 + 0], (double)jx);
                atomicAdd(&p.force[(size_t)oj * 3 + 1], (double)jy);
                atomicAdd(&p.force[(size_t)oj * 3 + 2], (double)jz);
            }
        }
    };

    // ---- MODE_GRAD flush: the block owns its rows -> plain stores ---------------
    // Transpose the warp's (bin x atom) accumulators through its own slice of
    // the idle pair-record buffers so that lane L holds bin m0 + L of atom a:
    // every store instruction then writes one contiguous 128-byte piece of ONE
    // row (coalesced for HBM and for mapped host memory alike).  The first flush
    // of a job stores, later ones (further element runs) add to what the same
    // thread stored before.
    static_assert(C <= 32, "one lane per bin of the chunk");
    const int oi = p.orig[gi];
    double srun = 0.0;  // MODE_GRAD: this job's F(Q) pair sum of bin m0 + lane
    // Float32 accuracy over long rows.  One float32 accumulator per (atom, bin)
    // running over all j of a 50 000-atom row loses a digit against the
    // float64 mode (measured 1.3e-5 instead of 1.7e-6 on grad F, 1.1e-6 instead
    // of 8e-8 on F(Q)).  Every acc_j j atoms the block therefore parks its
    // gradient partial sums in a scratch slot (thread-private layout: each
    // thread reads back only what it wrote; [value][thread] so the accesses
    // coalesce; the slots in use fit the L2) and folds the F(Q) partial sums
    // into the float64 `srun`; the row flush adds the parked sums back.  The
    // order of all these additions is fixed: results stay bit-reproducible.
    int slot = -1;           // scratch slot of this block (acquired on first use)
    bool parked = false;     // the slot holds partial sums of the current run
    int jacc = 0;            // j atoms accumulated since the last park
    __shared__ int slot_sh;
    [[maybe_unused]] auto park = [&](int b) {
      if constexpr (MODE == MODE_GRAD) {
        if (slot < 0) {
            if (threadIdx.x == 0) {
                // start at this SM's own slots: consecutive blocks of an SM reuse the
                // same slot, so the scratch in use (one slot per resident block)
                // stays in the L2 instead of wandering over all the slots
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                int s0 = (int)((smid * 4u) % (unsigned)p.n_slots), got = -1;
                while (got < 0)
                    for (int k = 0; k < p.n_slots && got < 0; ++k) {
                        const int t = (s0 + k) % p.n_slots;
                        if (atomicCAS(&p.slot_busy[t], 0, 1) == 0) got = t;
                    }
                slot_sh = got;
            }
            __syncthreads();
            slot = slot_sh;
        }
        if (active) {
            float2 *scr = reinterpret_cast<float2 *>(p.Gscr) +
                          (size_t)slot * (3 * H) * blockDim.x + threadIdx.x;
            // first park of a run: plain stores; later ones: fire-and-forget
            // vector RED (no read-back stall).  Every address belongs to one
            // thread and its operations arrive in program order, so the sums
            // are the same on every run.
            auto put = [&](float2(&acc)[MODE == MODE_GRAD ? H : 1], int comp) {
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    float2 *q = scr + (size_t)(comp * H + k) * blockDim.x;
                    if (parked) atomicAdd(q, acc[k]);
                    else *q = acc[k];
                    acc[k] = make_float2(0.f, 0.f);
                }
            };
            put(accX, 0);
            put(accY, 1);
            put(accZ, 2);
            if (p.S != nullptr && !nof) {
                // F(Q): transpose through the record buffer just consumed
                float *tr = tab(b) + warp * (C * 33);
#pragma unroll
                for (int m = 0; m < C; ++m) {
                    tr[m * 33 + lane] = m >= H ? accF[m % H].y : accF[m % H].x;
                }
                __syncwarp();
                const int bin = m0 + lane;
                if (lane < C && bin < p.nq) {
                    double v0 = 0.0, v1 = 0.0;
#pragma unroll
                    for (int a = 0; a < 32; a += 2) {
                        v0 += (double)tr[lane * 33 + a];
                        v1 += (double)tr[lane * 33 + a + 1];
                    }
                    srun += (v0 + v1) * (double)(fa[bin] * fb[bin]);
                }
#pragma unroll
                for (int k = 0; k < H; ++k) accF[k] = make_float2(0.f, 0.f);
            }
        }
        parked = true;
        jacc = 0;
        __syncthreads();  // buffer b is produced into at the start of the next tile
      }
    };
    [[maybe_unused]] auto flush_rows = [&](bool add) {
      if constexpr (MODE == MODE_GRAD) {
        if (parked) {  // add the parked partial sums of this run back
            __threadfence();  // this thread's REDs have been performed (L2); loads bypass L1
            const float2 *scr = reinterpret_cast<const float2 *>(p.Gscr) +
                                (size_t)slot * (3 * H) * blockDim.x + threadIdx.x;
            auto back = [&](float2(&acc)[MODE == MODE_GRAD ? H : 1], int comp) {
                float2 t[H];  // all loads of a component in flight together
#pragma unroll
                for (int k = 0; k < H; ++k) t[k] = __ldcg(scr + (size_t)(comp * H + k) * blockDim.x);
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    acc[k].x += t[k].x;
                    acc[k].y += t[k].y;
                }
            };
            back(accX, 0);
            back(accY, 1);
            back(accZ, 2);
        }
        float *tr = reinterpret_cast<float *>(smem_raw) + warp * (C * 33);
        const int bin = m0 + lane;
        const bool binok = lane < C && bin < p.nq;
        const float sc = binok ? fa[bin] * fb[bin] * inv_na[bin] : 0.f;
        float *Gp = dest < 0 ? reinterpret_cast<float *>(p.G)
                             : reinterpret_cast<float *>(p.Gside) + (size_t)dest * (32 * 3) * p.nq;
        auto put = [&](const float2(&acc)[MODE == MODE_GRAD ? H : 1], int comp) {
#pragma unroll
            for (int m = 0; m < C; ++m) tr[m * 33 + lane] = m >= H ? acc[m % H].y : acc[m % H].x;
            __syncwarp();
#pragma unroll 4
            for (int a = 0; a < 32; ++a) {
                const int oa = dest < 0 ? __shfl_sync(0xffffffffu, oi, a) : a;
                if (binok && oa >= 0) {
                    float *q = Gp + ((size_t)oa * 3 + comp) * p.nq + bin;
                    float v = tr[lane * 33 + a] * sc;
                    if (add) v += *q;
                    *q = v;
                }
            }
            __syncwarp();
        };
        put(accX, 0);
        put(accY, 1);
        put(accZ, 2);
        if (p.S != nullptr) {
#pragma unroll
            for (int m = 0; m < C; ++m) tr[m * 33 + lane] = m >= H ? accF[m % H].y : accF[m % H].x;
            __syncwarp();
            if (binok) {
                double v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int a = 0; a < 32; a += 2) {
                    v0 += (double)tr[lane * 33 + a];
                    v1 += (double)tr[lane * 33 + a + 1];
                }
                srun += (v0 + v1) * (double)(fa[bin] * fb[bin]);
            }
            __syncwarp();
        }
      }
    };

    // warps w and w+4 share a scheduler: one produces the next tile before its
    // consume phase, the other after it, so one of them always feeds the FP32 pipe
    // (measured: pairing by warp & 1, or no stagger at all, costs 6 %)
    const bool early = ((warp >> 2) & 1) == 0;
    bool first = true;
    for (;;) {
        const int ntile = (sg.jend - sg.jbegin) / TJ2;  // segments are multiples of 32
        produce(sg.jbegin, 0);
        __syncthreads();
        for (int t = 0; t < ntile; ++t) {
            const int b = t & 1;
            const bool has_next = t + 1 < ntile;
            const int jnext = sg.jbegin + (t + 1) * TJ2;
            if (has_next && early) produce(jnext, b ^ 1);
            if (active) {
                if constexpr (MODE == MODE_GRAD && CHEB) {
                    if (nof) consume(std::true_type{}, b);
                    else consume(std::false_type{}, b);
                } else {
                    consume(std::false_type{}, b);
                }
            }
            if (has_next && !early) produce(jnext, b ^ 1);
            __syncthreads();
            if constexpr (MODE == MODE_FORCE) {
                if (!diag) {
                    reduce_j(b, sg.jbegin + t * TJ2);
                    __syncthreads();  // records of buffer b are rewritten next iteration
                }
            }
            if constexpr (MODE == MODE_GRAD) {
                jacc += TJ2;
                if (p.acc_j > 0 && jacc >= p.acc_j && has_next) park(b);
            }
        }
        if constexpr (MODE != MODE_GRAD) {
            break;
        } else {
            ++seg;
            const bool last = seg >= seg_end;
            if ((sg.info & SEG_FLUSH) || last) {
                if (active) {
                    flush_rows(!first);
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        accF[k] = make_float2(0.f, 0.f);
                        accX[k] = make_float2(0.f, 0.f);
                        accY[k] = make_float2(0.f, 0.f);
                        accZ[k] = make_float2(0.f, 0.f);
                    }
                }
                first = false;
                parked = false;
                jacc = 0;
                if (!last) __syncthreads();  // the scratch is the next segment's record buffer
            }
            if (last) break;
            sg = p.segs[seg];
            diag = (sg.info & ITEM_DIAG) != 0;
            nof = p.grad_split && ((sg.info & ITEM_NOF) != 0 || p.S == nullptr);
            fw = (diag || !p.grad_split) ? 0.5f : 1.f;
            fb = ftab + (size_t)(sg.info & 0xffff) * p.qp;
        }
    }

    if constexpr (MODE == MODE_GRAD) {
        if (slot >= 0) {  // every thread has read its parked sums back: free the slot
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicExch(&p.slot_busy[slot], 0);
            }
        }
    }
    if constexpr (MODE == MODE_FORCE) {
        if (p.Ffix != nullptr) {
            // i side: the warps' partial forces meet in shared memory, one
            // fixed-point atomic per (item, atom, component)
            float *fs = reinterpret_cast<float *>(smem_raw);  // [nwarp][3][32]
            fs[(warp * 3 + 0) * 32 + lane] = active ? fix : 0.f;
            fs[(warp * 3 + 1) * 32 + lane] = active ? fiy : 0.f;
            fs[(warp * 3 + 2) * 32 + lane] = active ? fiz : 0.f;
            __syncthreads();
            for (int e = threadIdx.x; e < 96; e += blockDim.x) {  // (a block may be one warp)
                const int a = e & 31, w = e >> 5;
                double t = 0.0;
                for (int k = 0; k < nwarp; ++k) t += (double)fs[(k * 3 + w) * 32 + a];
                const int oa = p.orig[itile * TILE_I + a];
                if (oa >= 0) fix_add(p.Ffix + (size_t)oa * 3 + w, t, p.fix_scale);
            }
            return;
        }
    }
    if (!active) return;
    if constexpr (MODE == MODE_GRAD) {
        // per-job partial of the F(Q) pair sum; reduce_spart_kernel adds the jobs
        // in a fixed order
        const int bin = m0 + lane;
        if (p.S != nullptr && lane < C && bin < p.nq) p.S[(size_t)bx * p.qp + bin] = srun;
    } else if constexpr (MODE == MODE_FORCE) {
        if (oi >= 0) {
            atomicAdd(&p.force[(size_t)oi * 3 + 0], (double)fix);
            atomicAdd(&p.force[(size_t)oi * 3 + 1], (double)fiy);
            atomicAdd(&p.force[(size_t)oi * 3 + 2], (double)fiz);
        }
    } else {
        const double fweight = diag ? 0.5 : 1.0;
        // S[bin] += sum over the 32 atoms i of this warp.  Transpose the
        // warp's (bin x atom) accumulators through its own slice of the
        // (now idle) pair-record buffers, so that lane L sums bin m0 + L
        // in float64 and the warp issues ONE 32-wide atomic instead of 32
        // butterfly reductions (the flush was 1/3 of the instructions of a
        // 32 x 32-atom item).
        float *tr = reinterpret_cast<float *>(smem_raw) + warp * (C * 33);
#pragma unroll
        for (int m = 0; m < C; ++m) {
            const int k = m % H;
            tr[m * 33 + lane] = m >= H ? accF[k].y : accF[k].x;
        }
        __syncwarp();
        const int bin = m0 + lane;
        if (lane < C && bin < p.nq) {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int a = 0; a < 32; a += 2) {
                v0 += (double)tr[lane * 33 + a];
                v1 += (double)tr[lane * 33 + a + 1];
            }
            const double part = fweight * (v0 + v1) * (double)(fa[bin] * fb[bin]);
            if (p.Sfix != nullptr) fix_add2(p.Sfix + bin, p.Sfix + p.qp + bin, part, p.fix_scale);
            else if (p.Sitem != nullptr) p.Sitem[(size_t)bx * p.qp + bin] = part;  // deterministic
            else atomicAdd(&p.S[bin], part);
        }
    }
}

template <int C, int MODE, int MAXT, int MINB, int TJ2, bool CHEB, int PU = 1>
__global__ void __launch_bounds__(MAXT, MINB) debye2_kernel(const DebyeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (MODE == MODE_FQ && p.gate != nullptr && p.gate[3] != 0.0) return;  // the histogram pass ran
    debye2_body<C, MODE, TJ2, CHEB, PU>(p, smem_raw, (int)blockIdx.x, (int)blockIdx.y);
}

// G rows of the split i-tiles = sum of their pieces in the side buffer, in a
// fixed order.  One block per split row (32 atoms x 3 x nq values).
constexpr int FIX_SPLIT = 8;  // blocks per split row (4 atoms each)
template <typename TG>
__global__ void __launch_bounds__(256) rows_fixup_kernel(const RowFix *__restrict__ fix,
                                                         const int *__restrict__ orig,
                                                         const TG *__restrict__ side, int nq,
                                                         TG *__restrict__ G)
{
    const RowFix f = fix[blockIdx.x];
    const int per_atom = 3 * nq;
    constexpr int APB = TILE_I / FIX_SPLIT;
    const int a0 = blockIdx.y * APB;
    for (int e = threadIdx.x; e < APB * per_atom; e += blockDim.x) {
        const int a = a0 + e / per_atom, rem = e % per_atom;
        const int oa = orig[f.itile * TILE_I + a];
        if (oa < 0) continue;
        TG v = side[((size_t)f.d0 * 32 + a) * per_atom + rem];
        for (int d = f.d0 + 1; d < f.d1; ++d) v += side[((size_t)d * 32 + a) * per_atom + rem];
        G[(size_t)oa * per_atom + rem] = v;
    }
}

// S[m] = sum over the row jobs of their partial sums, fixed order: thread
// (tx, ty) adds jobs ty, ty + 32, ... of bin 32 blockIdx.x + tx, then the 32
// slices are added in order.
__global__ void __launch_bounds__(1024) reduce_spart_kernel(const double *__restrict__ part,
                                                            int njobs, int nq, int qp,
                                                            double *__restrict__ S)
{
    __shared__ double sl[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (m < nq)
        for (int j = ty; j < njobs; j += 32) acc += part[(size_t)j * qp + m];
    sl[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && m < nq) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sl[k][tx];
        S[m] = t;
    }
}

}  // namespace iid
