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
// overhead is 3 shared-memory loads; its bin loop is pure FP32 pipe work:
// 10 instructions per bin for F + grad F, 5 for F only, 6 for the force.
#pragma once
#include "iid_debye.cuh"

namespace iid {

constexpr int TJ2 = 16;  // j atoms per produced tile

__host__ __device__ inline size_t debye2_buf_bytes(int nwarp)
{
    return (size_t)TJ2 * 32 * sizeof(float) * (7 + 2 * (size_t)nwarp);
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

template <int C, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) debye2_kernel(const DebyeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double sfj[MODE == MODE_FORCE ? 2 * 3 * TJ2 : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int chunk0 = blockIdx.y * nwarp;
    const int m0 = (chunk0 + warp) * C;
    const bool active = m0 < p.nq;

    const WorkItem it = p.items[p.item_begin + (int64_t)blockIdx.x * p.item_stride];
    const bool diag = (it.info & ITEM_DIAG) != 0;
    const int btype = it.info & 0xffff;
    const int atype = p.tile_type[it.itile];
    const int gi = it.itile * TILE_I + lane;
    const double xi = p.x[gi], yi = p.y[gi], zi = p.z[gi];
    const bool vi = p.valid[gi] != 0.f;

    const float *ftab = reinterpret_cast<const float *>(p.ftab);
    const float *fa = ftab + (size_t)atype * p.qp;
    const float *fb = ftab + (size_t)btype * p.qp;
    const float *inv_na = reinterpret_cast<const float *>(p.inv_na);

    constexpr int NPAIR = TJ2 * 32;
    const size_t buf_bytes = debye2_buf_bytes(nwarp);
    // structure of arrays, [field][jj][lane]: scalar, conflict-free loads that
    // leave the register allocator free to keep the FFMA operand pairs in
    // opposite register banks (vector loads would pin their parities)
    auto tab = [&](int b) { return reinterpret_cast<float *>(smem_raw + (size_t)b * buf_bytes); };

    float accF[MODE != MODE_FORCE ? C : 1];
    float accX[MODE == MODE_GRAD ? C : 1], accY[MODE == MODE_GRAD ? C : 1],
        accZ[MODE == MODE_GRAD ? C : 1];
    float w0[MODE == MODE_FORCE ? C : 1], w1[MODE == MODE_FORCE ? C : 1];
    float fix = 0.f, fiy = 0.f, fiz = 0.f;
#pragma unroll
    for (int m = 0; m < C; ++m) {
        if constexpr (MODE != MODE_FORCE) accF[m] = 0.f;
        if constexpr (MODE == MODE_GRAD) { accX[m] = 0.f; accY[m] = 0.f; accZ[m] = 0.f; }
        if constexpr (MODE == MODE_FORCE) {
            const int bin = m0 + m;
            float w = 0.f;
            if (bin < p.nq) w = (float)(p.wq[bin]) * fa[bin] * fb[bin] * inv_na[bin];
            w0[m] = w;
            w1[m] = w * (float)bin;
        }
    }
    if constexpr (MODE == MODE_FORCE)
        for (int t = threadIdx.x; t < 2 * 3 * TJ2; t += blockDim.x) sfj[t] = 0.0;

    // ---- producer: the pair records of one j tile -----------------------------
    auto produce = [&](int jt, int b) {
        float *T = tab(b);
        float *S = T + 7 * NPAIR;
        for (int pr = threadIdx.x; pr < NPAIR; pr += blockDim.x) {
            const int jj = pr >> 5;  // (pr & 31) == lane: this thread's own atom i
            const int gj = jt + jj;
            const double dxd = p.x[gj] - xi, dyd = p.y[gj] - yi, dzd = p.z[gj] - zi;
            const bool keep = vi && p.valid[gj] != 0.f;
            const double r2 = fma(dxd, dxd, fma(dyd, dyd, dzd * dzd));
            const float r2f = (float)r2;
            double y = (double)rsqrtf(r2f);
            y = y * fma(-0.5 * r2, y * y, 1.5);  // one Newton step in float64
            if (!(keep && r2f > 0.f)) y = 0.0;   // self pair, ghost atom, r == 0
            const double r = r2 * y;
            const double u = r * p.qbin_turns;   // turns per Q bin
            float sth, cth, sC, cC, s, c;
            sincos_turns(u, sth, cth);
            sincos_turns(u * (double)C, sC, cC);
            if (chunk0 == 0) { s = 0.f; c = 1.f; }
            else sincos_turns(u * (double)(chunk0 * C), s, c);
            const float invr = (float)y;
            const float b3 = invr * invr * invr;
            s *= b3;
            c *= b3;
            T[pr] = cth;
            T[NPAIR + pr] = sth;
            T[2 * NPAIR + pr] = (float)(p.qbin * r);
            T[3 * NPAIR + pr] = r2f;
            T[4 * NPAIR + pr] = (float)dxd;
            T[5 * NPAIR + pr] = (float)dyd;
            T[6 * NPAIR + pr] = (float)dzd;
            for (int w = 0; w < nwarp; ++w) {
                S[(2 * w) * NPAIR + pr] = s;
                S[(2 * w + 1) * NPAIR + pr] = c;
                const float sn = fmaf(s, cC, c * sC);
                const float cn = fmaf(c, cC, -(s * sC));
                s = sn;
                c = cn;
            }
        }
    };

    // ---- consumer: this warp's chunk of bins for pairs (lane, jj) --------------
    auto consume = [&](int b, int jlo, int jhi, int fbuf) {
        const float *T = tab(b) + lane;
        const float *S = T + (7 + 2 * warp) * NPAIR;
        float n_cth = T[jlo * 32], n_sth = T[NPAIR + jlo * 32], n_kap = T[2 * NPAIR + jlo * 32],
              n_r2 = T[3 * NPAIR + jlo * 32], n_dx = T[4 * NPAIR + jlo * 32],
              n_dy = T[5 * NPAIR + jlo * 32], n_dz = T[6 * NPAIR + jlo * 32],
              n_s = S[jlo * 32], n_c = S[NPAIR + jlo * 32];
        for (int jj = jlo; jj < jhi; ++jj) {
            const float cth = n_cth, sth = n_sth, kap = n_kap, r2 = n_r2;
            const float dx = n_dx, dy = n_dy, dz = n_dz;
            float s = n_s, c = n_c;
            {   // prefetch the next pair's record
                const int jn = min(jj + 1, jhi - 1) * 32;
                n_cth = T[jn]; n_sth = T[NPAIR + jn]; n_kap = T[2 * NPAIR + jn];
                n_r2 = T[3 * NPAIR + jn]; n_dx = T[4 * NPAIR + jn];
                n_dy = T[5 * NPAIR + jn]; n_dz = T[6 * NPAIR + jn];
                n_s = S[jn]; n_c = S[NPAIR + jn];
            }
            float mk = kap * (float)m0;
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int m = 0; m < C; ++m) {
                if constexpr (MODE != MODE_FORCE) accF[m] = fmaf(s, r2, accF[m]);
                if constexpr (MODE == MODE_GRAD) {
                    const float a = fmaf(mk, c, -s);
                    accX[m] = fmaf(a, dx, accX[m]);
                    accY[m] = fmaf(a, dy, accY[m]);
                    accZ[m] = fmaf(a, dz, accZ[m]);
                    mk += kap;
                }
                if constexpr (MODE == MODE_FORCE) {
                    p1 = fmaf(w1[m], c, p1);
                    p0 = fmaf(w0[m], s, p0);
                }
                const float sn = fmaf(s, cth, c * sth);
                const float cn = fmaf(c, cth, -(s * sth));
                s = sn;
                c = cn;
            }
            if constexpr (MODE == MODE_FORCE) {
                const float phi = fmaf(kap, p1, -p0);
                fix = fmaf(phi, dx, fix);
                fiy = fmaf(phi, dy, fiy);
                fiz = fmaf(phi, dz, fiz);
                if (!diag) {
                    const float jx = warp_sum(-phi * dx);
                    const float jy = warp_sum(-phi * dy);
                    const float jz = warp_sum(-phi * dz);
                    if (lane == 0) {
                        atomicAdd(&sfj[fbuf * 3 * TJ2 + 3 * jj], (double)jx);
                        atomicAdd(&sfj[fbuf * 3 * TJ2 + 3 * jj + 1], (double)jy);
                        atomicAdd(&sfj[fbuf * 3 * TJ2 + 3 * jj + 2], (double)jz);
                    }
                }
            }
        }
    };

    const int ntile = (it.jend - it.jbegin) / TJ2;  // slabs are multiples of 32
    // warps w and w+4 share a scheduler: stagger their producer phases so one
    // of them always feeds the FP32 pipe
    const bool early = ((warp >> 2) & 1) == 0;
    produce(it.jbegin, 0);
    __syncthreads();
    for (int t = 0; t < ntile; ++t) {
        const int b = t & 1;
        const bool has_next = t + 1 < ntile;
        const int jnext = it.jbegin + (t + 1) * TJ2;
        if (has_next && early) produce(jnext, b ^ 1);
        if (active) consume(b, 0, TJ2 / 2, b);
        if (has_next && !early) produce(jnext, b ^ 1);
        if (active) consume(b, TJ2 / 2, TJ2, b);
        __syncthreads();
        if constexpr (MODE == MODE_FORCE) {
            if (!diag && threadIdx.x < 3 * TJ2) {
                const int k = threadIdx.x;
                const int oj = p.orig[it.jbegin + t * TJ2 + k / 3];
                const double v = sfj[b * 3 * TJ2 + k];
                sfj[b * 3 * TJ2 + k] = 0.0;
                if (oj >= 0) atomicAdd(&p.force[(size_t)oj * 3 + k % 3], v);
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
    } else {
        const double fweight = (MODE == MODE_GRAD || diag) ? 0.5 : 1.0;
        float *G = reinterpret_cast<float *>(p.G);
#pragma unroll
        for (int m = 0; m < C; ++m) {
            const int bin = m0 + m;
            if (bin < p.nq) {  // warp-uniform
                const float ff = fa[bin] * fb[bin];
                if constexpr (MODE == MODE_GRAD) if (oi >= 0) {
                    const float sc = ff * inv_na[bin];
                    float *row = G + (size_t)oi * 3 * p.nq + bin;
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
