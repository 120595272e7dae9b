/*
 * stencil.cuh -- the 4th-order variable-coefficient finite-volume operator at one cell.
 *
 * Restates the Poisson form of the reference macro apply_op_ijk (operators.fv4.c:87-114) with the
 * SAME association order, because the parity target is bit-for-bit agreement with the gcc -O2
 * (no FMA) build of the reference: C evaluates a+b+c as (a+b)+c, and the library is compiled
 * with -fmad=false so nvcc emits separate DMUL/DADD exactly as written here.
 *
 *   A x = -b h^-2 { 1/12 * SUM_6faces  beta_face * ( 15 (x_nbr - x) - (x_nbr2 - x_opp) )
 *                 + 1/48 * SUM_12      (transverse difference of beta_face) * (4-point mixed difference of x) }
 *
 * beta_d[ijk] is the coefficient on the LOW d-face of cell ijk.  25 x-points (axis +-1,+-2 and the
 * 12 in-plane diagonals: ghost faces + edges, never corners), 30 beta-points.
 *
 * The arithmetic is written ONCE, over four "loaders" X, BI, BJ, BK that return the value at an
 * offset (di,dj,dk) from the cell: the global-memory kernels pass pointer loaders, the tiled kernels
 * pass shared-memory loaders; both therefore produce identical bits.
 */
#ifndef HPGMG_B200_STENCIL_CUH
#define HPGMG_B200_STENCIL_CUH

#define STENCIL_TWELFTH (0.0833333333333333333)

/* aii (optional): the diagonal entry of A at this cell, i.e. the same expression evaluated on the unit vector
 * (x_c = 1, every other point 0) -- what rebuild_operator_blackbox (rebuild.c:127-133) accumulates into Aii for a
 * cell whose stencil does not reach a boundary-condition ghost cell: 15*(0-1)-(0-0) = -15 on each face, every
 * mixed difference (+-0)*beta = +-0, so Aii = -b*h2inv*(1/12 * SUM_faces beta*(-15)) with the faces added in the
 * macro's order.  Dinv = 1.0/Aii (rebuild.c:181) can then be formed in registers instead of being read. */
template <class XL, class BIL, class BJL, class BKL>
__device__ __forceinline__ double fv4_apply_op_at(const XL &X, const BIL &BI, const BJL &BJ, const BKL &BK, const double b, const double h2inv, double *aii = nullptr)
{
  const double xc = X(0, 0, 0);
  const double xw = X(-1, 0, 0), xe = X(1, 0, 0), xww = X(-2, 0, 0), xee = X(2, 0, 0);
  const double xs = X(0, -1, 0), xn = X(0, 1, 0), xss = X(0, -2, 0), xnn = X(0, 2, 0);
  const double xd = X(0, 0, -1), xu = X(0, 0, 1), xdd = X(0, 0, -2), xuu = X(0, 0, 2);
  const double xwn = X(-1, 1, 0), xws = X(-1, -1, 0), xen = X(1, 1, 0), xes = X(1, -1, 0);
  const double xwu = X(-1, 0, 1), xwd = X(-1, 0, -1), xeu = X(1, 0, 1), xed = X(1, 0, -1);
  const double xsu = X(0, -1, 1), xsd = X(0, -1, -1), xnu = X(0, 1, 1), xnd = X(0, 1, -1);

  const double bi0 = BI(0, 0, 0), bi1 = BI(1, 0, 0), bj0 = BJ(0, 0, 0), bj1 = BJ(0, 1, 0), bk0 = BK(0, 0, 0), bk1 = BK(0, 0, 1);
  const double axial =
      bi0 * (15.0 * (xw - xc) - (xww - xe))
    + bi1 * (15.0 * (xe - xc) - (xee - xw))
    + bj0 * (15.0 * (xs - xc) - (xss - xn))
    + bj1 * (15.0 * (xn - xc) - (xnn - xs))
    + bk0 * (15.0 * (xd - xc) - (xdd - xu))
    + bk1 * (15.0 * (xu - xc) - (xuu - xd));
  if (aii) {
    const double unit = bi0 * (-15.0) + bi1 * (-15.0) + bj0 * (-15.0) + bj1 * (-15.0) + bk0 * (-15.0) + bk1 * (-15.0);
    *aii = -b * h2inv * (STENCIL_TWELFTH * unit);
  }

  const double mixed =
      (BI(0, 1, 0) - BI(0, -1, 0)) * (xwn - xn - xws + xs)
    + (BI(0, 0, 1) - BI(0, 0, -1)) * (xwu - xu - xwd + xd)
    + (BJ(1, 0, 0) - BJ(-1, 0, 0)) * (xes - xe - xws + xw)
    + (BJ(0, 0, 1) - BJ(0, 0, -1)) * (xsu - xu - xsd + xd)
    + (BK(1, 0, 0) - BK(-1, 0, 0)) * (xed - xe - xwd + xw)
    + (BK(0, 1, 0) - BK(0, -1, 0)) * (xnd - xn - xsd + xs)

    + (BI(1, 1, 0) - BI(1, -1, 0)) * (xen - xn - xes + xs)
    + (BI(1, 0, 1) - BI(1, 0, -1)) * (xeu - xu - xed + xd)
    + (BJ(1, 1, 0) - BJ(-1, 1, 0)) * (xen - xe - xwn + xw)
    + (BJ(0, 1, 1) - BJ(0, 1, -1)) * (xnu - xu - xnd + xd)
    + (BK(1, 0, 1) - BK(-1, 0, 1)) * (xeu - xe - xwu + xw)
    + (BK(0, 1, 1) - BK(0, -1, 1)) * (xnu - xn - xsu + xs);

  return -b * h2inv * (STENCIL_TWELFTH * axial + 0.25 * STENCIL_TWELFTH * mixed);
}

/* loader over a global-memory array: p points at cell ijk */
struct GlobalLoader {
  const double *__restrict__ p;
  int jS, kS;
  __device__ __forceinline__ double operator()(const int di, const int dj, const int dk) const { return p[di + dj * jS + dk * kS]; }
};

/* x, bi, bj, bk point at cell ijk; jS/kS are the strides (doubles). Returns A x at ijk. */
__device__ __forceinline__ double fv4_apply_op(const double *__restrict__ x,
                                               const double *__restrict__ bi,
                                               const double *__restrict__ bj,
                                               const double *__restrict__ bk,
                                               const int jS, const int kS,
                                               const double b, const double h2inv)
{
  const GlobalLoader X = { x, jS, kS }, BI = { bi, jS, kS }, BJ = { bj, jS, kS }, BK = { bk, jS, kS };
  return fv4_apply_op_at(X, BI, BJ, BK, b, h2inv);
}

#endif
