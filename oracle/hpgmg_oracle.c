/*
 * hpgmg_oracle.c -- ORACLE: a plain-C, single-threaded restatement of the reference's fv4 FMG path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hpgmg_b200/ may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs use oracle/.
 *
 * Scope: a hierarchy in which EVERY level is one cubical box (the reference's `hpgmg-fv L 1`:
 * 2^L, 2^(L-1), ... 2 cells per side), homogeneous Dirichlet, Poisson
 * (a=0,b=1), GSRB or Chebyshev smoother, BiCGStab bottom solver.  With one box per level there
 * is no ghost exchange, so the whole algorithm is: boundary conditions + stencil + transfers.
 * Multi-box decompositions are checked against the reference build itself (oracle/_ref).
 *
 * Parity is PINNED: tests/test_oracle.py checks oracle_fmg() against the goldens produced by the
 * unmodified reference (`hpgmg-fv 6 1`: F-cycle norms of the 64^3/32^3/16^3 solves, the Richardson
 * error and the per-level Gershgorin bounds) and,
 * where oracle/_ref is present, against the reference's arrays cell by cell.
 *
 * Every function cites the reference lines it follows (paths relative to finite-volume/source/).
 * Arithmetic keeps the reference's association order; build with -ffp-contract=off.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hpgmg_oracle.h"

#define G 2                                   /* ghost depth = stencil radius (operators.fv4.c:138) */
enum { V_TEMP, V_U, V_F, V_E, V_R, V_DINV, V_BI, V_BJ, V_BK, V_RESERVED };   /* defines.h:28-38 */

/* ---- layout: level.c:935-938 ------------------------------------------------------------------ */
static void level_init(olevel *L, int n, int nvec)
{
  L->n = n;
  L->jS = n + 2 * G;
  while (L->jS % 4) L->jS++;
  L->kS = L->jS * (n + 2 * G);
  L->vol = L->kS * (n + 2 * G);
  L->nvec = nvec;
  L->v = (double **)calloc((size_t)nvec, sizeof(double *));
  for (int i = 0; i < nvec; i++) L->v[i] = (double *)calloc((size_t)L->vol, sizeof(double));
  L->eig = 0.0;
}
static void level_free(olevel *L)
{
  for (int i = 0; i < L->nvec; i++) free(L->v[i]);
  free(L->v);
}
static inline double *cell0(const olevel *L, int id) { return L->v[id] + G * (1 + L->jS + L->kS); }
#define IDX(L, i, j, k) ((i) + (j) * (L)->jS + (k) * (L)->kS)

/* ---- the operator: operators.fv4.c:87-114 (Poisson branch) ------------------------------------ */
#define TWELFTH (0.0833333333333333333)
static inline double apply_op_ijk(const double *x, const double *bi, const double *bj, const double *bk, int ijk, int jS, int kS, double b, double h2inv)
{
  return -b * h2inv * (
    TWELFTH * (
      + bi[ijk     ] * (15.0 * (x[ijk - 1 ] - x[ijk]) - (x[ijk - 2     ] - x[ijk + 1 ]))
      + bi[ijk + 1 ] * (15.0 * (x[ijk + 1 ] - x[ijk]) - (x[ijk + 2     ] - x[ijk - 1 ]))
      + bj[ijk     ] * (15.0 * (x[ijk - jS] - x[ijk]) - (x[ijk - 2 * jS] - x[ijk + jS]))
      + bj[ijk + jS] * (15.0 * (x[ijk + jS] - x[ijk]) - (x[ijk + 2 * jS] - x[ijk - jS]))
      + bk[ijk     ] * (15.0 * (x[ijk - kS] - x[ijk]) - (x[ijk - 2 * kS] - x[ijk + kS]))
      + bk[ijk + kS] * (15.0 * (x[ijk + kS] - x[ijk]) - (x[ijk + 2 * kS] - x[ijk - kS])))
    + 0.25 * TWELFTH * (
      + (bi[ijk + jS] - bi[ijk - jS]) * (x[ijk - 1 + jS] - x[ijk + jS] - x[ijk - 1 - jS] + x[ijk - jS])
      + (bi[ijk + kS] - bi[ijk - kS]) * (x[ijk - 1 + kS] - x[ijk + kS] - x[ijk - 1 - kS] + x[ijk - kS])
      + (bj[ijk + 1 ] - bj[ijk - 1 ]) * (x[ijk - jS + 1] - x[ijk + 1 ] - x[ijk - jS - 1] + x[ijk - 1 ])
      + (bj[ijk + kS] - bj[ijk - kS]) * (x[ijk - jS + kS] - x[ijk + kS] - x[ijk - jS - kS] + x[ijk - kS])
      + (bk[ijk + 1 ] - bk[ijk - 1 ]) * (x[ijk - kS + 1] - x[ijk + 1 ] - x[ijk - kS - 1] + x[ijk - 1 ])
      + (bk[ijk + jS] - bk[ijk - jS]) * (x[ijk - kS + jS] - x[ijk + jS] - x[ijk - kS - jS] + x[ijk - jS])
      + (bi[ijk + 1 + jS] - bi[ijk + 1 - jS]) * (x[ijk + 1 + jS] - x[ijk + jS] - x[ijk + 1 - jS] + x[ijk - jS])
      + (bi[ijk + 1 + kS] - bi[ijk + 1 - kS]) * (x[ijk + 1 + kS] - x[ijk + kS] - x[ijk + 1 - kS] + x[ijk - kS])
      + (bj[ijk + jS + 1] - bj[ijk + jS - 1]) * (x[ijk + jS + 1] - x[ijk + 1 ] - x[ijk + jS - 1] + x[ijk - 1 ])
      + (bj[ijk + jS + kS] - bj[ijk + jS - kS]) * (x[ijk + jS + kS] - x[ijk + kS] - x[ijk + jS - kS] + x[ijk - kS])
      + (bk[ijk + kS + 1] - bk[ijk + kS - 1]) * (x[ijk + kS + 1] - x[ijk + 1 ] - x[ijk + kS - 1] + x[ijk - 1 ])
      + (bk[ijk + kS + jS] - bk[ijk + kS - jS]) * (x[ijk + kS + jS] - x[ijk + jS] - x[ijk + kS - jS] + x[ijk - jS])));
}

/* ---- boundary conditions: operators/boundary_fv.c ---------------------------------------------- */
/* which of the 26 ghost regions a shape covers: level.c:420-424 */
static int shape_has(int shape, int di, int dj, int dk)
{
  int m = (di != 0) + (dj != 0) + (dk != 0);
  if (m == 0) return 0;
  if (shape == OR_SHAPE_STAR) return m == 1;
  if (shape == OR_SHAPE_NO_CORNERS) return m <= 2;
  return 1;
}
static inline void quartic(double x1, double x2, double x3, double x4, double *near, double *far)
{
  const double OneTwelfth = 1.0 / 12.0;                                       /* boundary_fv.c:296,339-340 */
  *near = OneTwelfth * (-77.0 * x1 + 43.0 * x2 - 17.0 * x3 + 3.0 * x4);
  *far  = OneTwelfth * (-505.0 * x1 + 335.0 * x2 - 145.0 * x3 + 27.0 * x4);
}

/* apply_BCs_v4 (boundary_fv.c:262-569) for one region with outward normal (di,dj,dk) of a single box:
 * faces extrapolate along the normal, edges along both normals (lower axis first, :404-425), corners
 * along i then j then k (:507-565). */
static void bc_v4_region(const olevel *L, double *x, int di, int dj, int dk)
{
  const int n = L->n, st[3] = { 1, L->jS, L->kS }, nrm[3] = { di, dj, dk };
  int lo[3], hi[3], t[3], in[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = nrm[a] ? 0 : 0;  hi[a] = nrm[a] ? 1 : n;           /* tangential axes run over the box */
    t[a] = nrm[a] < 0 ? -1 : n;                                 /* nearest ghost cell              */
    in[a] = nrm[a] < 0 ? st[a] : -st[a];                        /* one step inward                 */
  }
  for (int c = lo[2]; c < hi[2]; c++)
  for (int b = lo[1]; b < hi[1]; b++)
  for (int a = lo[0]; a < hi[0]; a++) {
    int p[3] = { a, b, c }, ijk = 0;
    for (int q = 0; q < 3; q++) ijk += (nrm[q] ? t[q] : p[q]) * st[q];
    double v[4][4][4];
    int ci = di ? 4 : 1, cj = dj ? 4 : 1, ck = dk ? 4 : 1;
    for (int K = 0; K < ck; K++) for (int J = 0; J < cj; J++) for (int I = 0; I < ci; I++)
      v[I][J][K] = x[ijk + (di ? (I + 1) * in[0] : 0) + (dj ? (J + 1) * in[1] : 0) + (dk ? (K + 1) * in[2] : 0)];
    if (di) { for (int K = 0; K < ck; K++) for (int J = 0; J < cj; J++) { double nn, ff; quartic(v[0][J][K], v[1][J][K], v[2][J][K], v[3][J][K], &nn, &ff); v[0][J][K] = nn; v[1][J][K] = ff; } ci = 2; }
    if (dj) { for (int K = 0; K < ck; K++) for (int I = 0; I < ci; I++) { double nn, ff; quartic(v[I][0][K], v[I][1][K], v[I][2][K], v[I][3][K], &nn, &ff); v[I][0][K] = nn; v[I][1][K] = ff; } cj = 2; }
    if (dk) { for (int J = 0; J < cj; J++) for (int I = 0; I < ci; I++) { double nn, ff; quartic(v[I][J][0], v[I][J][1], v[I][J][2], v[I][J][3], &nn, &ff); v[I][J][0] = nn; v[I][J][1] = ff; } ck = 2; }
    for (int K = 0; K < ck; K++) for (int J = 0; J < cj; J++) for (int I = 0; I < ci; I++)
      x[ijk - (di ? I * in[0] : 0) - (dj ? J * in[1] : 0) - (dk ? K * in[2] : 0)] = v[I][J][K];
  }
}

/* apply_BCs_v2 (boundary_fv.c:101-250): zero the region, then the first ghost layer only */
static void bc_v2_region(const olevel *L, double *x, int di, int dj, int dk)
{
  const int n = L->n, st[3] = { 1, L->jS, L->kS }, nrm[3] = { di, dj, dk };
  int lo[3], ext[3], t[3], d[3], nd = 0;
  for (int a = 0; a < 3; a++) {
    lo[a] = nrm[a] < 0 ? -G : (nrm[a] > 0 ? n : 0);
    ext[a] = nrm[a] ? G : n;
    t[a] = nrm[a] < 0 ? -1 : n;
    if (nrm[a]) d[nd++] = nrm[a] < 0 ? st[a] : -st[a];
  }
  for (int c = 0; c < ext[2]; c++) for (int b = 0; b < ext[1]; b++) for (int a = 0; a < ext[0]; a++)
    x[IDX(L, a + lo[0], b + lo[1], c + lo[2])] = 0.0;                          /* :139-145 (box_ghosts>1) */
  for (int c = 0; c < (nrm[2] ? 1 : n); c++)
  for (int b = 0; b < (nrm[1] ? 1 : n); b++)
  for (int a = 0; a < (nrm[0] ? 1 : n); a++) {
    int p[3] = { a, b, c }, ijk = 0;
    for (int q = 0; q < 3; q++) ijk += (nrm[q] ? t[q] : p[q]) * st[q];
    if (nd == 1) {
      x[ijk] = -2.5 * x[ijk + d[0]] + 0.5 * x[ijk + 2 * d[0]];                                     /* :169 */
    } else if (nd == 2) {
      x[ijk] = 6.25 * x[ijk + d[0] + d[1]] - 1.25 * x[ijk + 2 * d[0] + d[1]]
             - 1.25 * x[ijk + d[0] + 2 * d[1]] + 0.25 * x[ijk + 2 * d[0] + 2 * d[1]];              /* :206-209 */
    } else {
      x[ijk] = -15.625 * x[ijk + d[0] + d[1] + d[2]]
              + 3.125 * x[ijk + 2 * d[0] + d[1] + d[2]] + 3.125 * x[ijk + d[0] + 2 * d[1] + d[2]] + 3.125 * x[ijk + d[0] + d[1] + 2 * d[2]]
              - 0.625 * x[ijk + 2 * d[0] + 2 * d[1] + d[2]] - 0.625 * x[ijk + d[0] + 2 * d[1] + 2 * d[2]] - 0.625 * x[ijk + 2 * d[0] + d[1] + 2 * d[2]]
              + 0.125 * x[ijk + 2 * d[0] + 2 * d[1] + 2 * d[2]];                                   /* :238-245 */
    }
  }
}

void oracle_apply_BCs_v2(const olevel *L, int id, int shape)
{
  double *x = cell0(L, id);
  for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++)
    if (shape_has(shape, di, dj, dk)) bc_v2_region(L, x, di, dj, dk);
}
void oracle_apply_BCs_v4(const olevel *L, int id, int shape)
{
  if (L->n < 4) { oracle_apply_BCs_v2(L, id, shape); return; }               /* boundary_fv.c:269 */
  double *x = cell0(L, id);
  for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++)
    if (shape_has(shape, di, dj, dk)) bc_v4_region(L, x, di, dj, dk);
}

/* extrapolate_betas (boundary_fv.c:573-681): every ghost region of the box, in region order 0..26,
 * cells k,j,i ascending; beta_d is not extrapolated along d */
void oracle_extrapolate_betas(const olevel *L)
{
  const int n = L->n, jS = L->jS, kS = L->kS;
  double *b[3] = { cell0(L, V_BI), cell0(L, V_BJ), cell0(L, V_BK) };
  for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++) {
    if (!di && !dj && !dk) continue;
    const int nrm[3] = { di, dj, dk };
    int lo[3], ext[3];
    for (int a = 0; a < 3; a++) { lo[a] = nrm[a] < 0 ? -G : (nrm[a] > 0 ? n : 0); ext[a] = nrm[a] ? G : n; }
    /* inward strides with the component along the coefficient's own direction removed (:636-638) */
    const int s[3] = { -dj * jS - dk * kS, -di - dk * kS, -di - dj * jS };
    const int skip[3] = { (dj == 0 && dk == 0), (di == 0 && dk == 0), (di == 0 && dj == 0) };   /* pure d-face */
    for (int c = 0; c < ext[2]; c++) for (int bb = 0; bb < ext[1]; bb++) for (int a = 0; a < ext[0]; a++) {
      const int ijk = IDX(L, a + lo[0], bb + lo[1], c + lo[2]);
      for (int q = 0; q < 3; q++) {
        if (skip[q]) continue;
        double *be = b[q];
        const int d = s[q];
        if (n >= 5)      be[ijk] = 5.0 * be[ijk + d] - 10.0 * be[ijk + 2 * d] + 10.0 * be[ijk + 3 * d] - 5.0 * be[ijk + 4 * d] + be[ijk + 5 * d];
        else if (n >= 4) be[ijk] = 4.0 * be[ijk + d] - 6.0 * be[ijk + 2 * d] + 4.0 * be[ijk + 3 * d] - be[ijk + 4 * d];
        else if (n >= 2) be[ijk] = 2.0 * be[ijk + d] - be[ijk + 2 * d];
      }
    }
  }
}

/* ---- smoothers / residual / apply_op ------------------------------------------------------------ */
static void fill_ghosts(const olevel *L, int id) { oracle_apply_BCs_v4(L, id, OR_SHAPE_NO_CORNERS); }  /* exchange is empty for one box */

void oracle_apply_op(const olevel *L, int Ax_id, int x_id, double b)                 /* apply_op.c:9-50 */
{
  fill_ghosts(L, x_id);
  const double h2inv = 1.0 / (L->h * L->h);
  const double *x = cell0(L, x_id), *bi = cell0(L, V_BI), *bj = cell0(L, V_BJ), *bk = cell0(L, V_BK);
  double *Ax = cell0(L, Ax_id);
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) {
    const int ijk = IDX(L, i, j, k);
    Ax[ijk] = apply_op_ijk(x, bi, bj, bk, ijk, L->jS, L->kS, b, h2inv);
  }
}
void oracle_residual(const olevel *L, int res_id, int x_id, int rhs_id, double b)    /* residual.c:9-51 */
{
  fill_ghosts(L, x_id);
  const double h2inv = 1.0 / (L->h * L->h);
  const double *x = cell0(L, x_id), *rhs = cell0(L, rhs_id), *bi = cell0(L, V_BI), *bj = cell0(L, V_BJ), *bk = cell0(L, V_BK);
  double *res = cell0(L, res_id);
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) {
    const int ijk = IDX(L, i, j, k);
    res[ijk] = rhs[ijk] - apply_op_ijk(x, bi, bj, bk, ijk, L->jS, L->kS, b, h2inv);
  }
}
void oracle_smooth_gsrb(const olevel *L, int x_id, int rhs_id, double b)             /* gsrb.c:24-132, NUM_SMOOTHS=3, GSRB_OOP */
{
  const double h2inv = 1.0 / (L->h * L->h);
  const double *rhs = cell0(L, rhs_id), *bi = cell0(L, V_BI), *bj = cell0(L, V_BJ), *bk = cell0(L, V_BK), *Dinv = cell0(L, V_DINV);
  for (int s = 0; s < 6; s++) {
    const int src = (s & 1) ? V_TEMP : x_id, dst = (s & 1) ? x_id : V_TEMP;
    fill_ghosts(L, src);
    const double *xn = cell0(L, src);
    double *xp = cell0(L, dst);
    const int color000 = s & 1;                                                       /* box low = (0,0,0) */
    for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) {
      for (int i = 0; i < L->n; i++) xp[IDX(L, i, j, k)] = xn[IDX(L, i, j, k)];
      for (int i = ((j ^ k ^ color000) & 1); i < L->n; i += 2) {
        const int ijk = IDX(L, i, j, k);
        const double Ax = apply_op_ijk(xn, bi, bj, bk, ijk, L->jS, L->kS, b, h2inv);
        xp[ijk] = xn[ijk] + Dinv[ijk] * (rhs[ijk] - Ax);
      }
    }
  }
}
void oracle_smooth_cheby(const olevel *L, int x_id, int rhs_id, double b)            /* chebyshev.c:8-100, degree 6 */
{
  const double h2inv = 1.0 / (L->h * L->h);
  double beta = 1.000 * L->eig, alpha = 0.125000 * beta, theta = 0.5 * (beta + alpha), delta = 0.5 * (beta - alpha);
  double sigma = theta / delta, rho_n = 1 / sigma, c1[6], c2[6];
  c1[0] = 0.0;  c2[0] = 1 / theta;
  for (int s = 1; s < 6; s++) { double rho_nm1 = rho_n; rho_n = 1.0 / (2.0 * sigma - rho_nm1); c1[s] = rho_n * rho_nm1; c2[s] = rho_n * 2.0 / delta; }
  const double *rhs = cell0(L, rhs_id), *bi = cell0(L, V_BI), *bj = cell0(L, V_BJ), *bk = cell0(L, V_BK), *Dinv = cell0(L, V_DINV);
  for (int s = 0; s < 6; s++) {
    const int src = (s & 1) ? V_TEMP : x_id, dst = (s & 1) ? x_id : V_TEMP;
    fill_ghosts(L, src);
    const double *xn = cell0(L, src);
    double *xp = cell0(L, dst);                                                       /* x_nm1 aliases x_np1 */
    for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) {
      const int ijk = IDX(L, i, j, k);
      const double Ax = apply_op_ijk(xn, bi, bj, bk, ijk, L->jS, L->kS, b, h2inv);
      xp[ijk] = xn[ijk] + c1[s] * (xn[ijk] - xp[ijk]) + c2[s] * Dinv[ijk] * (rhs[ijk] - Ax);
    }
  }
}

/* ---- BLAS1: operators/misc.c --------------------------------------------------------------------- */
static void zero_vec(const olevel *L, int id) { memset(L->v[id], 0, (size_t)L->vol * sizeof(double)); }   /* interior + ghosts, misc.c:6-44 */
static void add_vec(const olevel *L, int c, double sa, int a, double sb, int b)
{
  double *C = cell0(L, c); const double *A = cell0(L, a), *B = cell0(L, b);
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) { const int q = IDX(L, i, j, k); C[q] = sa * A[q] + sb * B[q]; }
}
static void scale_vec(const olevel *L, int c, double sa, int a)
{
  double *C = cell0(L, c); const double *A = cell0(L, a);
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) { const int q = IDX(L, i, j, k); C[q] = sa * A[q]; }
}
static void mul_vec(const olevel *L, int c, double s, int a, int b)
{
  double *C = cell0(L, c); const double *A = cell0(L, a), *B = cell0(L, b);
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) { const int q = IDX(L, i, j, k); C[q] = s * A[q] * B[q]; }
}
/* dot: per tile (10000 x 8 x 8, level.h:34-44) in k,j,i order, tiles in list order (misc.c:239-282) */
static double dot_vec(const olevel *L, int a, int b)
{
  const double *A = cell0(L, a), *B = cell0(L, b);
  double total = 0.0;
  for (int kk = 0; kk < L->n; kk += 8) for (int jj = 0; jj < L->n; jj += 8) {
    double part = 0.0;
    for (int k = kk; k < kk + 8 && k < L->n; k++) for (int j = jj; j < jj + 8 && j < L->n; j++) for (int i = 0; i < L->n; i++) part += A[IDX(L, i, j, k)] * B[IDX(L, i, j, k)];
    total += part;
  }
  return total;
}
double oracle_norm(const olevel *L, int id)                                             /* max norm, misc.c:287-329 */
{
  const double *A = cell0(L, id);
  double m = 0.0;
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) { double f = fabs(A[IDX(L, i, j, k)]); if (f > m) m = f; }
  return m;
}

/* ---- inter-level transfers ------------------------------------------------------------------------ */
void oracle_restriction(const olevel *Lc, int id_c, const olevel *Lf, int id_f, int type)   /* restriction.c:6-94 */
{
  const double *r = cell0(Lf, id_f);
  double *w = cell0(Lc, id_c);
  const int rj = Lf->jS, rk = Lf->kS, half = Lf->n / 2;
  const int ei = half + (type == OR_RESTRICT_FACE_I), ej = half + (type == OR_RESTRICT_FACE_J), ek = half + (type == OR_RESTRICT_FACE_K);   /* mg.c:574-588 */
  for (int k = 0; k < ek; k++) for (int j = 0; j < ej; j++) for (int i = 0; i < ei; i++) {
    const int q = IDX(Lf, 2 * i, 2 * j, 2 * k);
    double v;
    if (type == OR_RESTRICT_CELL)        v = (r[q] + r[q + 1] + r[q + rj] + r[q + 1 + rj] + r[q + rk] + r[q + 1 + rk] + r[q + rj + rk] + r[q + 1 + rj + rk]) * 0.125;
    else if (type == OR_RESTRICT_FACE_I) v = (r[q] + r[q + rj] + r[q + rk] + r[q + rj + rk]) * 0.25;
    else if (type == OR_RESTRICT_FACE_J) v = (r[q] + r[q + 1] + r[q + rk] + r[q + 1 + rk]) * 0.25;
    else                                 v = (r[q] + r[q + 1] + r[q + rj] + r[q + 1 + rj]) * 0.25;
    w[IDX(Lc, i, j, k)] = v;
  }
}

/* tensor-product prolongation, W=3: interpolation_v2.c:112-172 (c1=1/8); W=5: interpolation_v4.c:149-238 */
static inline void pro(int W, const double *c, int stride, double *lo, double *hi)
{
  if (W == 3) { const double c1 = 1.0 / 8.0; *lo = (c[0] + c1 * (c[-stride] - c[stride])); *hi = (c[0] - c1 * (c[-stride] - c[stride])); }
  else { const double c2 = -3.0 / 128.0, c1 = 22.0 / 128.0;
         *lo = (c[0] + c1 * (c[-stride] - c[stride]) + c2 * (c[-2 * stride] - c[2 * stride]));
         *hi = (c[0] - c1 * (c[-stride] - c[stride]) - c2 * (c[-2 * stride] - c[2 * stride])); }
}
static void interpolation(int W, const olevel *Lf, int id_f, double prescale, const olevel *Lc, int id_c)
{
  const int R = W / 2;
  if (W == 3) oracle_apply_BCs_v2(Lc, id_c, OR_SHAPE_BOX); else oracle_apply_BCs_v4(Lc, id_c, OR_SHAPE_BOX);   /* interpolation_v2.c:211-212, v4.c:277-278 */
  const double *rd = cell0(Lc, id_c);
  double *wr = cell0(Lf, id_f);
  for (int kk = 0; kk < Lc->n; kk++) for (int jj = 0; jj < Lc->n; jj++) for (int ii = 0; ii < Lc->n; ii++) {
    double fi[2][5][5], fj[2][2][5];
    for (int K = 0; K < W; K++) for (int J = 0; J < W; J++)
      pro(W, rd + IDX(Lc, ii, jj + J - R, kk + K - R), 1, &fi[0][J][K], &fi[1][J][K]);
    for (int K = 0; K < W; K++) for (int I = 0; I < 2; I++) {
      double col[5];
      for (int J = 0; J < W; J++) col[J] = fi[I][J][K];
      pro(W, col + R, 1, &fj[I][0][K], &fj[I][1][K]);
    }
    for (int J = 0; J < 2; J++) for (int I = 0; I < 2; I++) {
      double col[5], lo, hi;
      for (int K = 0; K < W; K++) col[K] = fj[I][J][K];
      pro(W, col + R, 1, &lo, &hi);
      double *w0 = wr + IDX(Lf, 2 * ii + I, 2 * jj + J, 2 * kk);
      w0[0] = prescale * w0[0] + lo;
      w0[Lf->kS] = prescale * w0[Lf->kS] + hi;
    }
  }
}
void oracle_interpolation_v2(const olevel *Lf, int id_f, double prescale, const olevel *Lc, int id_c) { interpolation(3, Lf, id_f, prescale, Lc, id_c); }
void oracle_interpolation_v4(const olevel *Lf, int id_f, double prescale, const olevel *Lc, int id_c) { interpolation(5, Lf, id_f, prescale, Lc, id_c); }

/* ---- setup ------------------------------------------------------------------------------------------ */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
static double evalBeta(double x, double y, double z, double h, int xx, int yy, int zz)     /* problem.fv.c:9-27 */
{
  double b = 0.25, a = 2.0 * M_PI;
  double B = 1.0 + b * sin(a * x) * sin(a * y) * sin(a * z);
  double Bxx = -a * a * b * sin(a * x) * sin(a * y) * sin(a * z), Byy = Bxx, Bzz = Bxx;
  if (xx) B += (h * h / 24.0) * Bxx;
  if (yy) B += (h * h / 24.0) * Byy;
  if (zz) B += (h * h / 24.0) * Bzz;
  return B;
}
static double evalF(double x, double y, double z, double h)                                  /* problem.fv.c:31-86, all three corrections */
{
  double a = 2.0 * M_PI, p = 7.0;
  double sx = sin(a * x), sy = sin(a * y), sz = sin(a * z);
  double F = pow(sx, p) * pow(sy, p) * pow(sz, p);
  double Fxx = -a * a * p * pow(sx, p) * pow(sy, p) * pow(sz, p) + a * a * p * (p - 1) * pow(sx, p - 2) * pow(sy, p) * pow(sz, p) * pow(cos(a * x), 2);
  double Fyy = -a * a * p * pow(sx, p) * pow(sy, p) * pow(sz, p) + a * a * p * (p - 1) * pow(sx, p) * pow(sy, p - 2) * pow(sz, p) * pow(cos(a * y), 2);
  double Fzz = -a * a * p * pow(sx, p) * pow(sy, p) * pow(sz, p) + a * a * p * (p - 1) * pow(sx, p) * pow(sy, p) * pow(sz, p - 2) * pow(cos(a * z), 2);
  F += (h * h / 24.0) * Fxx;
  F += (h * h / 24.0) * Fyy;
  F += (h * h / 24.0) * Fzz;
  return F;
}
static void initialize_problem(olevel *L, double h)                                          /* problem.fv.c:90-140 */
{
  L->h = h;
  double *Bi = cell0(L, V_BI), *Bj = cell0(L, V_BJ), *Bk = cell0(L, V_BK), *F = cell0(L, V_F);
  for (int k = 0; k <= L->n; k++) for (int j = 0; j <= L->n; j++) for (int i = 0; i <= L->n; i++) {
    const int q = IDX(L, i, j, k);
    double x = h * ((double)i + 0.5), y = h * ((double)j + 0.5), z = h * ((double)k + 0.5);
    Bi[q] = evalBeta(x - h * 0.5, y, z, h, 0, 1, 1);
    Bj[q] = evalBeta(x, y - h * 0.5, z, h, 1, 0, 1);
    Bk[q] = evalBeta(x, y, z - h * 0.5, h, 1, 1, 0);
    F[q] = evalF(x, y, z, h);
  }
}

/* rebuild_operator_blackbox (rebuild.c:47-208) with 4 colours per dimension; VECTOR_E holds sum|Aij| */
static void rebuild_blackbox(olevel *L, double a, double b)
{
  int colors = 4;
  if (L->n < colors) colors = L->n;
  const double h2inv = 1.0 / (L->h * L->h);
  zero_vec(L, V_DINV);  zero_vec(L, V_E);
  double *x = cell0(L, V_TEMP), *Aii = cell0(L, V_DINV), *sum = cell0(L, V_E);
  const double *bi = cell0(L, V_BI), *bj = cell0(L, V_BJ), *bk = cell0(L, V_BK);
  for (int kc = 0; kc < colors; kc++) for (int jc = 0; jc < colors; jc++) for (int ic = 0; ic < colors; ic++) {
    for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++)         /* color_vector, misc.c:441-471 */
      x[IDX(L, i, j, k)] = (((i + ic) % colors) == 0 ? 1.0 : 0.0) * (((j + jc) % colors) == 0 ? 1.0 : 0.0) * (((k + kc) % colors) == 0 ? 1.0 : 0.0);
    fill_ghosts(L, V_TEMP);
    for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) {
      const int q = IDX(L, i, j, k);
      const double Ax = apply_op_ijk(x, bi, bj, bk, q, L->jS, L->kS, b, h2inv);
      Aii[q] += (x[q]) * Ax;
      sum[q] += fabs((1.0 - x[q]) * Ax);
    }
  }
  double eig = -1e9;
  for (int k = 0; k < L->n; k++) for (int j = 0; j < L->n; j++) for (int i = 0; i < L->n; i++) {
    const int q = IDX(L, i, j, k);
    if (Aii[q] == 0.0) Aii[q] = a + b * h2inv;
    double Di = (Aii[q] + sum[q]) / Aii[q];
    if (Di > eig) eig = Di;
    if (Aii[q] >= 1.5 * sum[q]) sum[q] = 1.0 / (Aii[q]); else sum[q] = 1.0 / (Aii[q] + 0.5 * sum[q]);
    Aii[q] = 1.0 / Aii[q];
  }
  L->eig = eig;
}
static void rebuild_operator(olevel *L, const olevel *from, double a, double b)               /* operators.fv4.c:145-173 */
{
  if (from) {
    oracle_restriction(L, V_BI, from, V_BI, OR_RESTRICT_FACE_I);
    oracle_restriction(L, V_BJ, from, V_BJ, OR_RESTRICT_FACE_J);
    oracle_restriction(L, V_BK, from, V_BK, OR_RESTRICT_FACE_K);
  }
  oracle_extrapolate_betas(L);
  rebuild_blackbox(L, a, b);                                                                   /* the 4 exchanges are empty for one box */
}

/* ---- bottom solver: solvers/bicgstab.c:14-97 --------------------------------------------------------- */
static int bicgstab(olevel *L, int x_id, int R_id, double b, double rtol)
{
  const int r0 = V_RESERVED, r = V_RESERVED + 1, p = V_RESERVED + 2, q = V_RESERVED + 3, s = V_RESERVED + 4, t = V_RESERVED + 5, Ap = V_RESERVED + 6, As = V_RESERVED + 7;
  int j = 0, failed = 0, converged = 0;
  oracle_residual(L, r0, x_id, R_id, b);
  scale_vec(L, r, 1.0, r0);
  scale_vec(L, p, 1.0, r0);
  double r_dot_r0 = dot_vec(L, r, r0), norm_of_r0 = oracle_norm(L, r);
  if (r_dot_r0 == 0.0) converged = 1;
  if (norm_of_r0 == 0.0) converged = 1;
  while (j < 200 && !failed && !converged) {
    j++;
    mul_vec(L, q, 1.0, V_DINV, p);
    oracle_apply_op(L, Ap, q, b);
    double Ap_dot_r0 = dot_vec(L, Ap, r0);
    if (Ap_dot_r0 == 0.0) { failed = 1; break; }
    double alpha = r_dot_r0 / Ap_dot_r0;
    if (isinf(alpha)) { failed = 2; break; }
    add_vec(L, x_id, 1.0, x_id, alpha, q);
    add_vec(L, s, 1.0, r, -alpha, Ap);
    double norm_of_s = oracle_norm(L, s);
    if (norm_of_s == 0.0) { converged = 1; break; }
    if (norm_of_s < rtol * norm_of_r0) { converged = 1; break; }
    mul_vec(L, t, 1.0, V_DINV, s);
    oracle_apply_op(L, As, t, b);
    double As_dot_As = dot_vec(L, As, As), As_dot_s = dot_vec(L, As, s);
    if (As_dot_As == 0.0) { converged = 1; break; }
    double omega = As_dot_s / As_dot_As;
    if (omega == 0.0) { failed = 3; break; }
    if (isinf(omega)) { failed = 4; break; }
    add_vec(L, x_id, 1.0, x_id, omega, t);
    add_vec(L, r, 1.0, s, -omega, As);
    double norm_of_r = oracle_norm(L, r);
    if (norm_of_r == 0.0) { converged = 1; break; }
    if (norm_of_r < rtol * norm_of_r0) { converged = 1; break; }
    double r_dot_r0_new = dot_vec(L, r, r0);
    if (r_dot_r0_new == 0.0) { failed = 5; break; }
    double beta = (r_dot_r0_new / r_dot_r0) * (alpha / omega);
    if (isinf(beta)) { failed = 6; break; }
    add_vec(L, V_TEMP, 1.0, p, -omega, Ap);
    add_vec(L, p, 1.0, r, beta, V_TEMP);
    r_dot_r0 = r_dot_r0_new;
  }
  return j;
}

/* ---- cycles: mg.c:1135-1164 (MGVCycle), :1237-1344 (FMGSolve) -------------------------------------- */
static void smooth(const ohier *H, const olevel *L, int x_id, int rhs_id) { if (H->cheby) oracle_smooth_cheby(L, x_id, rhs_id, H->b); else oracle_smooth_gsrb(L, x_id, rhs_id, H->b); }

static void vcycle(ohier *H, int e_id, int R_id, int l)
{
  if (l == H->nlevels - 1) { H->krylov += bicgstab(&H->L[l], e_id, R_id, H->b, 1e-3); return; }
  smooth(H, &H->L[l], e_id, R_id);
  oracle_residual(&H->L[l], V_TEMP, e_id, R_id, H->b);
  oracle_restriction(&H->L[l + 1], R_id, &H->L[l], V_TEMP, OR_RESTRICT_CELL);
  zero_vec(&H->L[l + 1], e_id);
  vcycle(H, e_id, R_id, l + 1);
  oracle_interpolation_v2(&H->L[l], e_id, 1.0, &H->L[l + 1], e_id);
  smooth(H, &H->L[l], e_id, R_id);
}

double oracle_fmg_solve(ohier *H, int onLevel, double *norm_of_F)
{
  const int e_id = V_U, R_id = V_R, bottom = H->nlevels - 1;
  zero_vec(&H->L[onLevel], V_U);                                                    /* bench_hpgmg, hpgmg-fv.c:78 */
  const double nF = oracle_norm(&H->L[onLevel], V_F);
  scale_vec(&H->L[onLevel], R_id, 1.0, V_F);
  for (int l = onLevel; l < bottom; l++) oracle_restriction(&H->L[l + 1], R_id, &H->L[l], R_id, OR_RESTRICT_CELL);
  if (bottom > onLevel) zero_vec(&H->L[bottom], e_id);
  H->krylov += bicgstab(&H->L[bottom], e_id, R_id, H->b, 1e-3);
  for (int l = bottom - 1; l >= onLevel; l--) {
    oracle_interpolation_v4(&H->L[l], e_id, 0.0, &H->L[l + 1], e_id);
    vcycle(H, e_id, R_id, l);
  }
  oracle_residual(&H->L[onLevel], V_TEMP, e_id, V_F, H->b);
  if (norm_of_F) *norm_of_F = nF;
  return oracle_norm(&H->L[onLevel], V_TEMP);
}

/* hpgmg-fv.c:280-308 + MGBuild (mg.c:842-1022) for a one-box-per-level hierarchy of 2^log2_dim cells */
ohier *oracle_build(int log2_dim, int cheby)
{
  ohier *H = (ohier *)calloc(1, sizeof(ohier));
  H->a = 0.0;  H->b = 1.0;  H->cheby = cheby;
  H->nlevels = log2_dim;                                          /* 2^L ... 2: a box is never smaller than the radius (mg.c:942) */
  H->L = (olevel *)calloc((size_t)H->nlevels, sizeof(olevel));
  for (int l = 0; l < H->nlevels; l++) level_init(&H->L[l], 1 << (log2_dim - l), V_RESERVED + (l == H->nlevels - 1 ? 8 : 0));
  initialize_problem(&H->L[0], 1.0 / (double)(1 << log2_dim));
  rebuild_operator(&H->L[0], NULL, H->a, H->b);
  for (int l = 1; l < H->nlevels; l++) { H->L[l].h = 2.0 * H->L[l - 1].h; rebuild_operator(&H->L[l], &H->L[l - 1], H->a, H->b); }
  return H;
}
void oracle_destroy(ohier *H)
{
  for (int l = 0; l < H->nlevels; l++) level_free(&H->L[l]);
  free(H->L);
  free(H);
}

/* the driver's Richardson pass (hpgmg-fv.c:351-366, mg.c:1113-1131): solves on levels 0,1,2 then
 * ||u2h - R uh|| and the observed order.  norms[3] receives the three F-cycle residual norms. */
void oracle_richardson(ohier *H, double norms[3], double *err, double *order)
{
  for (int l = 0; l < 3; l++) {
    if (l > 0) oracle_restriction(&H->L[l], V_F, &H->L[l - 1], V_F, OR_RESTRICT_CELL);
    norms[l] = oracle_fmg_solve(H, l, NULL);
  }
  oracle_restriction(&H->L[1], V_TEMP, &H->L[0], V_U, OR_RESTRICT_CELL);
  oracle_restriction(&H->L[2], V_TEMP, &H->L[1], V_U, OR_RESTRICT_CELL);
  add_vec(&H->L[1], V_TEMP, 1.0, V_U, -1.0, V_TEMP);
  add_vec(&H->L[2], V_TEMP, 1.0, V_U, -1.0, V_TEMP);
  const double e2h = oracle_norm(&H->L[1], V_TEMP), e4h = oracle_norm(&H->L[2], V_TEMP);
  *err = e2h;
  *order = log(e4h / e2h) / log(2);
}

/* raw access for cell-by-cell comparisons from Python */
double *oracle_vector(ohier *H, int level, int id) { return H->L[level].v[id]; }
int oracle_level_dim(ohier *H, int level) { return H->L[level].n; }
int oracle_level_jstride(ohier *H, int level) { return H->L[level].jS; }
int oracle_level_volume(ohier *H, int level) { return H->L[level].vol; }
double oracle_level_eig(ohier *H, int level) { return H->L[level].eig; }
int oracle_num_levels(ohier *H) { return H->nlevels; }
int oracle_krylov_iterations(ohier *H) { return H->krylov; }
void oracle_set_h(ohier *H, int level, double h) { H->L[level].h = h; }
olevel *oracle_level(ohier *H, int level) { return &H->L[level]; }
