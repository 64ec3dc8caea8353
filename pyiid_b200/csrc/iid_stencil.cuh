// iid_stencil.cuh -- the interpolation stencil shared by the radial tables:
// the force table of the fused evaluation kernel (iid_fused.cuh, values
// interpolated FROM a uniform r grid) and the pair histogram of the F(Q) pass
// (iid_fq_hist.cuh, pair weights spread ONTO the same grid -- the adjoint).
// FT_PTS-point Lagrange interpolation on a grid with Q_max h = FT_QH: the
// functions involved are band limited by Q_max, so the error is
// max |prod (u - j)| / FT_PTS! * (Q_max h)^FT_PTS = 2.2e-4 * 3^-12 = 4e-10 of
// their amplitude.
#pragma once

namespace iid {

constexpr int FT_PTS = 12;               // interpolation points
constexpr int FT_LEFT = FT_PTS / 2 - 1;  // nodes k - FT_LEFT .. k + FT_PTS - 1 - FT_LEFT
constexpr int FT_PAD = FT_PTS / 2;       // entries stored before r = 0
constexpr double FT_QH = 1.0 / 3.0;      // Q_max h

// barycentric weight of node i of P equispaced nodes: (-1)^(n-i) C(n, i) / n!, n = P - 1
template <int P>
__host__ __device__ constexpr double lagrange_bary(int i)
{
    double c = 1.0, f = 1.0;
    for (int k = 1; k <= P - 1; ++k) f *= (double)k;
    for (int k = 0; k < i; ++k) c = c * (double)(P - 1 - k) / (double)(k + 1);
    return (((P - 1 - i) & 1) ? -c : c) / f;
}
__host__ __device__ constexpr double ft_bary(int i) { return lagrange_bary<FT_PTS>(i); }

// The F(Q) pair histogram of large structures spreads with a SHORTER stencil on a
// FINER grid when the structure fits it: its cost is the shared-memory atomics per
// pair (two per node), and the same error bound max |prod (u - j)| / P! *
// (Q_max h)^P = 4e-10 holds for
//   12 points, Q_max h = 1/3     (2.2e-4 * 3^-12),
//    8 points, Q_max h = 0.157   (1.07e-3 * 0.157^8),
//    6 points, Q_max h = 0.0658  (4.88e-3 * 0.0658^6)
// -- 24, 16 or 12 atomics per pair on 1, 2.1 or 5.1 times as many nodes.
constexpr int FH_PTS_FINE = 8;
constexpr double FH_QH_FINE = 0.157;
constexpr int FH_PTS_FINEST = 6;
constexpr double FH_QH_FINEST = 0.0658;

}  // namespace iid
