// iid_debye64.cuh -- FP64 Debye-sum kernel, producer/consumer version (sm_100a).
//
// The float64 twin of iid_debye2.cuh: block = one (i-tile, j-slab) work item,
// warp = Q chunk of C bins, lane = atom i, accumulators in registers; the
// 32 x TJ pair records of a j tile (cos/sin theta, kappa, r^2, d, and one
// (sin, cos)(m0 theta)/r^3 seed per warp) are produced once per block into
// double-buffered shared memory.  Everything is float64 (there is no packed
// DFMA and no FP64 SFU on sm_100; the bound is the 64 lane/clk/SM DFMA pipe).
// Inside a chunk sin/cos advance by one rotation step from the seed and then
// by the three-term recurrence x[k+1] = 2cos(theta) x[k] - x[k-1]; in float64
// its rounding (1e-16 k / sin(theta)) is far below the 1e-10 tolerance, so no
// re-seeding inside the chunk is needed.  F + grad F: 8 DFMA-class
// instructions per bin per ordered pair.
#pragma once
#include <type_traits>
#include "iid_debye.cuh"

namespace iid {

constexpr int NREC64 = 8;  // doubles per pair record shared by all warps

__host__ __device__ inline size_t debye64_buf_bytes(int nwarp, int tj)
{
    return (size_t)tj * 32 * sizeof(double) * (NREC64 + 2 * (size_t)nwarp);
}
__host__ __device__ inline size_t debye64_phi_bytes(int nwarp, int tj)
{
    return (size_t)tj * 32 * sizeof(double) * (size_t)nwarp;
}

// sin, cos of 2 pi u for any u held in float64.  Exact reduction to an eighth of
// a turn (rint and one fma, both exact in float64), then the fdlibm kernel
// polynomials on |x| <= pi/4 (k_sin.c / k_cos.c: errors below 2^-58): about 25
// DFMA-class instructions against ~70 for sincospi, which repeats a general
// range reduction and handles special values the phases here never take.
__device__ __forceinline__ void sincos_turns64(double u, double &s, double &c)
{
    const double q4 = rint(u * 4.0);
    const double fr = fma(-0.25, q4, u);            // |fr| <= 1/8 turn, exact
    const int quad = (int)(long long)q4 & 3;
    const double x = 6.283185307179586476925286766559 * fr;
    const double z = x * x;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sf = fma(x * z, ps, x);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cf = fma(z * z, pc, fma(-0.5, z, 1.0));
    s = (quad & 1) ? cf : sf;
    c = (quad & 1) ? sf : cf;
    if (quad == 1 || quad == 2) c = -c;
    if (quad >= 2) s = -s;
}

template <int C, int MODE, int MAXT, int TJ>
__global__ void __launch_bounds__(MAXT, 1) debye64_kernel(const DebyeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw64[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int chunk0 = blockIdx.y * nwarp;
    const int m0 = (chunk0 + warp) * C;
    const bool active = m0 < p.nq;

    // MODE_GRAD: a row job (i-tile, consecutive j segments), see iid_debye.cuh;
    // the other modes: one work item = one segment
    int itile, seg = 0, seg_end = 1, dest = -1;
    RowSeg sg;
    if constexpr (MODE == MODE_GRAD) {
        const RowJob job = p.jobs[blockIdx.x];
        itile = job.itile;
        seg = job.seg_begin;
        seg_end = job.seg_end;
        dest = job.dest;
        sg = p.segs[seg];
    } else {
        const WorkItem it = p.items[p.item_begin + (int64_t)blockIdx.x * p.item_stride];
        itile = it.itile;
        sg.jbegin = it.jbegin;
        sg.jend = it.jend;
        sg.info = it.info;
    }
    bool diag = (sg.info & ITEM_DIAG) != 0;
    bool nof = MODE == MODE_GRAD && p.grad_split && ((sg.info & ITEM_NOF) != 0 || p.S == nullptr);
    double fw = (MODE == MODE_GRAD && (diag || !p.grad_split)) ? 0.5 : 1.0;
    const int atype = p.tile_type[itile];
    const int gi = itile * TILE_I + lane;
    const double xi = p.x[gi], yi = p.y[gi], zi = p.z[gi];
    const bool vi = p.valid[gi] != 0.f;

    const double *ftab = reinterpret_cast<const double *>(p.ftab);
    const double *fa = ftab + (size_t)atype * p.qp;
    const double *fb = ftab + (size_t)(sg.info & 0xffff) * p.qp;
    const double *inv_na = reinterpret_cast<const double *>(p.inv_na);

    constexpr int NPAIR = TJ * 32;
    const size_t buf_bytes = debye64_buf_bytes(nwarp, TJ);
    auto tab = [&](int b) { return reinterpret_cast<double2 *>(smem_raw64 + (size_t)b * buf_bytes); };
    double *phis = reinterpret_cast<double *>(smem_raw64 + 2 * buf_bytes);  // MODE_FORCE
    const int nactive = min(nwarp, (p.nq - chunk0 * C + C - 1) / C);

    double accF[MODE != MODE_FORCE ? C : 1];
    double accX[MODE == MODE_GRAD ? C : 1] = {}, accY[MODE == MODE_GRAD ? C : 1] = {},
           accZ[MODE == MODE_GRAD ? C : 1] = {};
    double w0[MODE == MODE_FORCE ? C : 1], w1[MODE == MODE_FORCE ? C : 1];
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        if constexpr (MODE != MODE_FORCE) accF[k] = 0.0;
        if constexpr (MODE == MODE_GRAD) { accX[k] = 0.0; accY[k] = 0.0; accZ[k] = 0.0; }
        if constexpr (MODE == MODE_FORCE) {
            const int bin = m0 + k;
            double w = 0.0;
            if (bin < p.nq) w = p.wq[bin] * fa[bin] * fb[bin] * inv_na[bin];
            w0[k] = w;
            w1[k] = w * (double)bin;
        }
    }

    // ---- producer -----------------------------------------------------------------
    // record layout (double2 arrays of NPAIR): [0] cos,sin theta  [1] kappa, r^2
    // [2] dx, dy  [3] dz, -   then one (sin, cos) seed per warp
    auto produce = [&](int jt, int b) {
        double2 *T = tab(b);
        for (int pr = threadIdx.x; pr < NPAIR; pr += blockDim.x) {
            const int jj = pr >> 5;  // (pr & 31) == lane
            const int gj = jt + jj;
            const double dx = p.x[gj] - xi, dy = p.y[gj] - yi, dz = p.z[gj] - zi;
            const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
            const bool ok = vi && p.valid[gj] != 0.f && (float)r2 > 0.f;
            // 1/r: MUFU.RSQ seed (2^-23) + two Newton steps (-> 1e-14, 1e-16);
            // cheaper than the IEEE sqrt + divide sequences
            double y = (double)rsqrtf((float)r2);
            y = y * fma(-0.5 * r2, y * y, 1.5);
            y = y * fma(-0.5 * r2, y * y, 1.5);
            const double invr = ok ? y : 0.0;
            const double r = r2 * invr;
            const double u = r * p.qbin_turns;  // turns per Q bin
            double sth, cth, sC, cC, s, c;
            sincos_turns64(u, sth, cth);
            // rotation by one chunk: e^{i C theta} by log2(C) squarings of
            // e^{i theta} (float64: the doubling of the 1e-16 rounding is harmless)
            sC = sth;
            cC = cth;
#pragma unroll
            for (int d = 1; d < C; d <<= 1) {
                const double s2 = 2.0 * sC * cC;
                cC = fma(cC, cC, -(sC * sC));
                sC = s2;
            }
            static_assert((C & (C - 1)) == 0, "C must be a power of two");
            if (chunk0 == 0) {
                s = 0.0;
                c = 1.0;
            } else {
                sincos_turns64(u * (double)(chunk0 * C), s, c);
            }
            const double b3 = invr * invr * invr;
            s *= b3;
            c *= b3;
            T[pr] = make_double2(cth, sth);
            T[NPAIR + pr] = make_double2(p.qbin * r, r2 * fw);
            T[2 * NPAIR + pr] = make_double2(dx, dy);
            T[3 * NPAIR + pr] = make_double2(dz, 0.0);
            for (int w = 0; w < nwarp; ++w) {
                T[(4 + w) * NPAIR + pr] = make_double2(s, c);
                const double sn = fma(s, cC, c * sC);
                const double cn = fma(c, cC, -(s * sC));
                s = sn;
                c = cn;
            }
        }
    };

    struct Rec {
        double2 cs, kr, dxy, dz, seed;
    };
    auto load_rec = [&](int b, int jj) {
        Rec r;
        const double2 *T = tab(b) + jj * 32 + lane;
        r.cs = T[0];
        r.kr = T[NPAIR];
        r.dxy = T[2 * NPAIR];
        r.dz = T[3 * NPAIR];
        r.seed = T[(4 + warp) * NPAIR];
        return r;
    };
    auto bins = [&](auto nof_tag, const Rec &rec, int jj) {
        constexpr bool NOF = decltype(nof_tag)::value;
        const double cth = rec.cs.x, sth = rec.cs.y, kap = rec.kr.x;
        [[maybe_unused]] const double r2 = rec.kr.y;
        const double dx = rec.dxy.x, dy = rec.dxy.y, dz = rec.dz.x;
        const double tc = cth + cth;
        double s = rec.seed.x, c = rec.seed.y, sp = s, cp = c;
        double mk = kap * (double)m0;
        if constexpr (MODE == MODE_GRAD && NOF) {
            // gradient only: a_k = k C_k + t_k with C = kappa c, t = m0 kappa c - s
            // (both obey the three-term recurrence; k is a compile-time constant),
            // 6.2 instead of 8 DFMA-class instructions per bin
            const double s1 = fma(s, cth, c * sth);
            const double c1 = fma(c, cth, -(s * sth));
            double tp = fma(mk, c, -s), t = fma(mk, c1, -s1);
            double Cp = kap * c, Cc = kap * c1;
            accX[0] = fma(tp, dx, accX[0]);
            accY[0] = fma(tp, dy, accY[0]);
            accZ[0] = fma(tp, dz, accZ[0]);
#pragma unroll
            for (int k = 1; k < C; ++k) {
                const double a = fma((double)k, Cc, t);
                accX[k] = fma(a, dx, accX[k]);
                accY[k] = fma(a, dy, accY[k]);
                accZ[k] = fma(a, dz, accZ[k]);
                if (k + 1 < C) {
                    const double tn = fma(tc, t, -tp);
                    const double Cn = fma(tc, Cc, -Cp);
                    tp = t; t = tn; Cp = Cc; Cc = Cn;
                }
            }
            return;
        }
        [[maybe_unused]] double p0[2] = {0.0, 0.0}, p1[2] = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < C; ++k) {
            if constexpr (MODE != MODE_FORCE) accF[k] = fma(s, r2, accF[k]);
            if constexpr (MODE == MODE_GRAD) {
                const double a = fma(mk, c, -s);
                accX[k] = fma(a, dx, accX[k]);
                accY[k] = fma(a, dy, accY[k]);
                accZ[k] = fma(a, dz, accZ[k]);
                mk += kap;
            }
            if constexpr (MODE == MODE_FORCE) {
                p1[k & 1] = fma(w1[k], c, p1[k & 1]);
                p0[k & 1] = fma(w0[k], s, p0[k & 1]);
            }
            if (k + 1 < C) {
                if (k == 0) {
                    const double sn = fma(s, cth, c * sth);
                    const double cn = fma(c, cth, -(s * sth));
                    sp = s; cp = c; s = sn; c = cn;
                } else {
                    const double sn = fma(tc, s, -sp);
                    sp = s;
                    s = sn;
                    if constexpr (MODE != MODE_FQ) {
                        const double cn = fma(tc, c, -cp);
                        cp = c;
                        c = cn;
                    }
                }
            }
        }
        if constexpr (MODE == MODE_FORCE) {
            const double phi = fma(kap, p1[0] + p1[1], -(p0[0] + p0[1]));
            fix = fma(phi, dx, fix);
            fiy = fma(phi, dy, fiy);
            fiz = fma(phi, dz, fiz);
            if (!diag) phis[warp * NPAIR + jj * 32 + lane] = phi;
        }
    };
    auto consume = [&](auto nof_tag, int b) {
        Rec r0 = load_rec(b, 0);
#pragma unroll 1
        for (int jj = 0; jj < TJ; jj += 2) {
            Rec r1 = load_rec(b, jj + 1);
            bins(nof_tag, r0, jj);
            r0 = load_rec(b, min(jj + 2, TJ - 1));
            bins(nof_tag, r1, jj + 1);
        }
    };
    auto reduce_j = [&](int b, int jt) {
        const double2 *T = tab(b) + lane;
        for (int jj = warp; jj < TJ; jj += nwarp) {
            double phi = 0.0;
            for (int w = 0; w < nactive; ++w) phi += phis[w * NPAIR + jj * 32 + lane];
            const double2 dxy = T[2 * NPAIR + jj * 32];
            const double2 dzz = T[3 * NPAIR + jj * 32];
            const double jx = warp_sum(-phi * dxy.x);
            const double jy = warp_sum(-phi * dxy.y);
            const double jz = warp_sum(-phi * dzz.x);
            const int oj = p.orig[jt + jj];
            if (lane == 0 && oj >= 0) {
                atomicAdd(&p.force[(size_t)oj * 3 + 0], jx);
                atomicAdd(&p.force[(size_t)oj * 3 + 1], jy);
                atomicAdd(&p.force[(size_t)oj * 3 + 2], jz);
            }
        }
    };

    // ---- MODE_GRAD flush: plain coalesced stores of the rows this block owns
    // (see iid_debye2.cuh)
    static_assert(C <= 32, "one lane per bin of the chunk");
    const int oi = p.orig[gi];
    double srun = 0.0;
    [[maybe_unused]] auto flush_rows = [&](bool add) {
      if constexpr (MODE == MODE_GRAD) {
        double *tr = reinterpret_cast<double *>(smem_raw64) + warp * (C * 33);
        const int bin = m0 + lane;
        const bool binok = lane < C && bin < p.nq;
        const double sc = binok ? fa[bin] * fb[bin] * inv_na[bin] : 0.0;
        double *Gp = dest < 0 ? reinterpret_cast<double *>(p.G)
                              : reinterpret_cast<double *>(p.Gside) + (size_t)dest * (32 * 3) * p.nq;
        auto put = [&](const double(&acc)[C], int comp) {
#pragma unroll
            for (int k = 0; k < C; ++k) tr[k * 33 + lane] = acc[k];
            __syncwarp();
#pragma unroll 4
            for (int a = 0; a < 32; ++a) {
                const int oa = dest < 0 ? __shfl_sync(0xffffffffu, oi, a) : a;
                if (binok && oa >= 0) {
                    double *q = Gp + ((size_t)oa * 3 + comp) * p.nq + bin;
                    double v = tr[lane * 33 + a] * sc;
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
            for (int k = 0; k < C; ++k) tr[k * 33 + lane] = accF[k];
            __syncwarp();
            if (binok) {
                double v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int a = 0; a < 32; a += 2) {
                    v0 += tr[lane * 33 + a];
                    v1 += tr[lane * 33 + a + 1];
                }
                srun += (v0 + v1) * (fa[bin] * fb[bin]);
            }
            __syncwarp();
        }
      }
    };

    const bool early = ((warp >> 2) & 1) == 0;
    bool first = true;
    for (;;) {
        const int ntile = (sg.jend - sg.jbegin) / TJ;  // segments are multiples of 32
        produce(sg.jbegin, 0);
        __syncthreads();
        for (int t = 0; t < ntile; ++t) {
            const int b = t & 1;
            const bool has_next = t + 1 < ntile;
            const int jnext = sg.jbegin + (t + 1) * TJ;
            if (has_next && early) produce(jnext, b ^ 1);
            if (active) {
                if constexpr (MODE == MODE_GRAD) {
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
                    reduce_j(b, sg.jbegin + t * TJ);
                    __syncthreads();
                }
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
                    for (int k = 0; k < C; ++k) {
                        accF[k] = 0.0;
                        accX[k] = 0.0;
                        accY[k] = 0.0;
                        accZ[k] = 0.0;
                    }
                }
                first = false;
                if (!last) __syncthreads();
            }
            if (last) break;
            sg = p.segs[seg];
            diag = (sg.info & ITEM_DIAG) != 0;
            nof = p.grad_split && ((sg.info & ITEM_NOF) != 0 || p.S == nullptr);
            fw = (diag || !p.grad_split) ? 0.5 : 1.0;
            fb = ftab + (size_t)(sg.info & 0xffff) * p.qp;
        }
    }

    if (!active) return;
    if constexpr (MODE == MODE_GRAD) {
        const int bin = m0 + lane;
        if (p.S != nullptr && lane < C && bin < p.nq) p.S[(size_t)blockIdx.x * p.qp + bin] = srun;
    } else if constexpr (MODE == MODE_FORCE) {
        if (oi >= 0) {
            atomicAdd(&p.force[(size_t)oi * 3 + 0], fix);
            atomicAdd(&p.force[(size_t)oi * 3 + 1], fiy);
            atomicAdd(&p.force[(size_t)oi * 3 + 2], fiz);
        }
    } else {
        const double fweight = diag ? 0.5 : 1.0;
        // transpose the warp's (bin x atom) accumulators through its slice of
        // the idle record buffers: lane L sums bin m0 + L, one atomic per bin
        double *tr = reinterpret_cast<double *>(smem_raw64) + warp * (C * 33);
#pragma unroll
        for (int k = 0; k < C; ++k) tr[k * 33 + lane] = accF[k];
        __syncwarp();
        const int bin = m0 + lane;
        if (lane < C && bin < p.nq) {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int a = 0; a < 32; a += 2) {
                v0 += tr[lane * 33 + a];
                v1 += tr[lane * 33 + a + 1];
            }
            const double part = fweight * (v0 + v1) * (fa[bin] * fb[bin]);
            if (p.Sitem != nullptr) p.Sitem[(size_t)blockIdx.x * p.qp + bin] = part;  // deterministic
            else atomicAdd(&p.S[bin], part);
        }
    }
}

}  // namespace iid
