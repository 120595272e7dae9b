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
 */
#ifndef HPGMG_B200_STENCIL_CUH
#define HPGMG_B200_STENCIL_CUH

#define STENCIL_TWELFTH (0.0833333333333333333)

/* x, bi, bj, bk point at cell ijk; jS/kS are the strides (doubles). Returns A x at ijk. */
__device__ __forceinline__ double fv4_apply_op(const double *__restrict__ x,
                                               const double *__restrict__ bi,
                                               const double *__restrict__ bj,
                                               const double *__restrict__ bk,
                                               const int jS, const int kS,
                                               const double b, const double h2inv)
{
  const double xc = x[0];
  const double xw = x[-1],  xe = x[1],   xww = x[-2],      xee = x[2];
  const double xs = x[-jS], xn = x[jS],  xss = x[-2 * jS], xnn = x[2 * jS];
  const double xd = x[-kS], xu = x[kS],  xdd = x[-2 * kS], xuu = x[2 * kS];
  const double xwn = x[-1 + jS], xws = x[-1 - jS], xen = x[1 + jS], xes = x[1 - jS];
  const double xwu = x[-1 + kS], xwd = x[-1 - kS], xeu = x[1 + kS], xed = x[1 - kS];
  const double xsu = x[-jS + kS], xsd = x[-jS - kS], xnu = x[jS + kS], xnd = x[jS - kS];

  const double axial =
      bi[0]  * (15.0 * (xw - xc) - (xww - xe))
    + bi[1]  * (15.0 * (xe - xc) - (xee - xw))
    + bj[0]  * (15.0 * (xs - xc) - (xss - xn))
    + bj[jS] * (15.0 * (xn - xc) - (xnn - xs))
    + bk[0]  * (15.0 * (xd - xc) - (xdd - xu))
    + bk[kS] * (15.0 * (xu - xc) - (xuu - xd));

  const double mixed =
      (bi[jS]     - bi[-jS])     * (xwn - xn - xws + xs)
    + (bi[kS]     - bi[-kS])     * (xwu - xu - xwd + xd)
    + (bj[1]      - bj[-1])      * (xes - xe - xws + xw)
    + (bj[kS]     - bj[-kS])     * (xsu - xu - xsd + xd)
    + (bk[1]      - bk[-1])      * (xed - xe - xwd + xw)
    + (bk[jS]     - bk[-jS])     * (xnd - xn - xsd + xs)

    + (bi[1 + jS] - bi[1 - jS])  * (xen - xn - xes + xs)
    + (bi[1 + kS] - bi[1 - kS])  * (xeu - xu - xed + xd)
    + (bj[jS + 1] - bj[jS - 1])  * (xen - xe - xwn + xw)
    + (bj[jS + kS] - bj[jS - kS]) * (xnu - xu - xnd + xd)
    + (bk[kS + 1] - bk[kS - 1])  * (xeu - xe - xwu + xw)
    + (bk[kS + jS] - bk[kS - jS]) * (xnu - xn - xsu + xs);

  return -b * h2inv * (STENCIL_TWELFTH * axial + 0.25 * STENCIL_TWELFTH * mixed);
}

#endif
