// iid_debye.cuh -- shared definitions of the pair-tiled Debye-sum kernels (sm_100a).
//
// The kernels cover the three pair sums of the hot path:
//   MODE_FQ    S[m]       = sum_{pairs} f_i f_j sin(Q_m r)/r            (triangle)
//   MODE_GRAD  G[i,w,m]   = sum_j f_i f_j a_ij(m) (q_j - q_i)_w / na    (square, + S)
//   MODE_FORCE force[i,w] = sum_m wq[m] G[i,w,m] without forming G      (triangle)
// with Q_m = m*qbin and a = (Q cos(Q r) - sin(Q r)/r)/r^2.  They replace the
// reference's chain of element-wise kernels that materialise K x Q and
// K x 3 x Q arrays (pyiid/experiments/elasticscatter/kernels/gpu_flat.py:11-299
// driven by atomics/gpu_atomics.py:89-280; CPU twins kernels/cpu_flat.py:14-194,
// kernels/cpu_experimental.py:8-15).  Nothing O(pairs) is ever stored.
//
// Mapping (iid_debye2.cuh FP32, iid_debye64.cuh FP64).  Atoms are sorted by
// element and every element run is padded to a multiple of 32 (ghost atoms carry
// valid = 0), so a 32-atom i-tile and a j-slab each have ONE element type and
// f_i f_j / na factors out of the pair loop.  A work item is (i-tile, j-slab).
// A block takes one item; warp w of the block owns a chunk of C Q bins and lane
// l owns atom i = 32*itile + l, so a thread keeps the accumulators of its
// (atom, chunk) in registers for the whole item.
//
// Transcendentals.  The Q grid is uniform, so sin/cos(m*theta), theta = qbin*r,
// are advanced by recurrences on the FP32 (FP64) FMA pipe instead of 2 MUFU per
// bin on the 16-lane SFU.  The state is pre-scaled by 1/r^3 so that
// a_ij(m) = m*kappa*c - s (kappa = qbin*r) costs one FMA.  Distances, phases in
// turns and their range reductions are computed in FLOAT64 even in FP32 mode
// (float32-rounded positions are exact in float64, so r and the seed phase
// m0*theta mod 2pi carry no float32 cancellation error even at Q r ~ 3000 rad).
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
constexpr int ITEM_NOF = 1 << 17;  // j range above the diagonal: gradient only
constexpr int SEG_FLUSH = 1 << 18; // row jobs: last segment of an element run -> flush

// Full-gradient pass (MODE_GRAD): ROW OWNERSHIP.  A block owns one row job =
// (i-tile, consecutive j segments); it keeps the 32 x 3 x Q accumulators of
// its atoms in registers over all segments and writes each gradient row
// exactly once with plain coalesced stores -- no zero-fill of the output, no
// atomics, bit-reproducible.  A job whose segments cover the whole j range
// (dest < 0) stores into the caller's G rows; the rows kept back for load
// balance are cut into pieces (dest >= 0) that store into a side buffer and
// are summed in a fixed order by rows_fixup_kernel.
struct RowSeg {
    int jbegin, jend;  // sorted/padded j range, one element run, multiples of 32
    int info;          // element type | ITEM_DIAG | ITEM_NOF | SEG_FLUSH
    int pad;
};
struct RowJob {
    int itile;               // atoms [32*itile, 32*itile + 32)
    int seg_begin, seg_end;  // segments [seg_begin, seg_end) of the segment array
    int dest;                // < 0: rows of G; >= 0: slot of the side buffer
};
// a split row for rows_fixup_kernel: G rows of tile itile = sum of side slots [d0, d1)
struct RowFix {
    int itile, d0, d1, pad;
};

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
    void *G;        // [n][3][nq] kernel precision (MODE_GRAD); device or mapped host memory
    double *S;      // [qp] (MODE_FQ); MODE_GRAD: [n_jobs][qp] per-job partial sums (may be null)
    double *force;  // [n][3] (MODE_FORCE)
    int grad_split;  // MODE_GRAD: 1 = F(Q) from the lower triangle only (ITEM_NOF segments skip it)
    // MODE_GRAD row jobs
    const RowJob *jobs;
    const RowSeg *segs;
    void *Gside;    // [n_pieces][32][3][nq] kernel precision
    // FP32 accuracy: a row job parks its float32 partial sums every acc_j j
    // atoms in a thread-private scratch slot (L2 resident) instead of letting
    // one float32 accumulator run over the whole row (0 = never)
    float *Gscr;    // [n_slots][3 C][block threads]
    int *slot_busy; // [n_slots] 0 = free
    int n_slots, acc_j;
    // Deterministic F(Q) of the standalone pass: every work item stores its
    // partial sums, added later in item order.  Null = atomics.
    double *Sitem = nullptr;  // MODE_FQ: [n_items][qp]
    // The fused evaluation kernel: FIXED-POINT accumulators.  Integer addition
    // is associative, so plain atomics give bit-reproducible sums in ONE phase
    // (no per-item partials, no ordered second pass).  value = count / fix_scale
    // with fix_scale a power of two chosen from a rigorous bound of the sum.
    unsigned long long *Sfix = nullptr;  // MODE_FQ: [2][qp] high word, low word (fix_add2)
    unsigned long long *Ffix = nullptr;  // MODE_FORCE: [n][3]
    double fix_scale = 0.0;
    // The stand-alone F(Q) kernel behind the histogram pass (iid_fq_hist.cuh): it
    // runs only when gate[3] == 0, i.e. the structure did not fit the histogram.
    const double *gate = nullptr;
};

__device__ __forceinline__ void fix_add(unsigned long long *a, double v, double scale)
{
    atomicAdd(a, (unsigned long long)__double2ll_rn(v * scale));
}

// Two words: the high word counts units of 1/scale, the low word what the
// rounding of the high word left, in units of 2^-40 / scale -- the sum keeps
// every bit a float64 sum could (the pair sums S feed a difference of nearly
// equal numbers in Rw).  v * scale is exact (power of two); so is the remainder.
constexpr double FIX_LOW = 1099511627776.0;  // 2^40
__device__ __forceinline__ void fix_add2(unsigned long long *hi, unsigned long long *lo, double v,
                                         double scale)
{
    const double t = v * scale;
    const long long h = __double2ll_rn(t);
    atomicAdd(hi, (unsigned long long)h);
    atomicAdd(lo, (unsigned long long)__double2ll_rn((t - (double)h) * FIX_LOW));
}

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

}  // namespace iid
