/*
 * problem.c -- the benchmark's analytic test problem (right-hand side and coefficient field).
 *
 * Restates /root/reference/finite-volume/source/operators/problem.fv.c: evaluateBeta :9-27,
 * evaluateF :31-86, initialize_problem :90-140.  Setup only (untimed in the reference).
 *
 * This stays on the HOST on purpose: f and beta are built from sin/cos/pow, and CUDA's libm
 * differs from glibc's in the last ulp, which would move every golden number.  The fields are
 * evaluated with glibc exactly as the reference does and then uploaded; only the entries the
 * reference writes (cells and low faces 0..dim inclusive) are touched.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hpgmg_b200.h"
#include "runtime.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* beta(x,y,z) = 1 + 1/4 sin(2 pi x) sin(2 pi y) sin(2 pi z), optionally converted from a point
 * value to a face average with the h^2/24 second-derivative correction in the chosen directions */
double evaluateBeta(double x, double y, double z, double h, int add_Bxx, int add_Byy, int add_Bzz)
{
  double b = 0.25;
  double a = 2.0 * M_PI;
  double B   = 1.0 + b * sin(a * x) * sin(a * y) * sin(a * z);
  double Bxx = -a * a * b * sin(a * x) * sin(a * y) * sin(a * z);
  double Byy = -a * a * b * sin(a * x) * sin(a * y) * sin(a * z);
  double Bzz = -a * a * b * sin(a * x) * sin(a * y) * sin(a * z);
  if (add_Bxx) B += (h * h / 24.0) * Bxx;
  if (add_Byy) B += (h * h / 24.0) * Byy;
  if (add_Bzz) B += (h * h / 24.0) * Bzz;
  return B;
}

/* f(x,y,z) = sin^7(2 pi x) sin^7(2 pi y) sin^7(2 pi z), converted to a cell average the same way */
double evaluateF(double x, double y, double z, double h, int add_Fxx, int add_Fyy, int add_Fzz)
{
  double a = 2.0 * M_PI;
  double p = 7.0;
  double F   = pow(sin(a * x), p) * pow(sin(a * y), p) * pow(sin(a * z), p);
  double Fxx = -a * a * p * pow(sin(a * x), p) * pow(sin(a * y), p) * pow(sin(a * z), p) + a * a * p * (p - 1) * pow(sin(a * x), p - 2) * pow(sin(a * y), p) * pow(sin(a * z), p) * pow(cos(a * x), 2);
  double Fyy = -a * a * p * pow(sin(a * x), p) * pow(sin(a * y), p) * pow(sin(a * z), p) + a * a * p * (p - 1) * pow(sin(a * x), p) * pow(sin(a * y), p - 2) * pow(sin(a * z), p) * pow(cos(a * y), 2);
  double Fzz = -a * a * p * pow(sin(a * x), p) * pow(sin(a * y), p) * pow(sin(a * z), p) + a * a * p * (p - 1) * pow(sin(a * x), p) * pow(sin(a * y), p) * pow(sin(a * z), p - 2) * pow(cos(a * z), 2);
  if (add_Fxx) F += (h * h / 24.0) * Fxx;
  if (add_Fyy) F += (h * h / 24.0) * Fyy;
  if (add_Fzz) F += (h * h / 24.0) * Fzz;
  return F;
}

void initialize_problem(level_type *L, double hLevel, double a, double b)
{
  (void)a; (void)b;
  L->h = hLevel;
  if (L->num_my_boxes == 0) return;
  const size_t vol = (size_t)L->box_volume;
  double *stage = (double *)malloc(4 * vol * sizeof(double));
  const int ids[4] = { VECTOR_BETA_I, VECTOR_BETA_J, VECTOR_BETA_K, VECTOR_F };

  for (int box = 0; box < L->num_my_boxes; box++) {
    const box_type *B = &L->my_boxes[box];
    const int jS = B->jStride, kS = B->kStride, g = B->ghosts, n = B->dim;
    for (int v = 0; v < 4; v++) hpgmg_download_box_vector(L, box, ids[v], stage + v * vol);
    double *Bi = stage, *Bj = stage + vol, *Bk = stage + 2 * vol, *F = stage + 3 * vol;
    #pragma omp parallel for collapse(2)
    for (int k = 0; k <= n; k++)                     /* 0..dim inclusive: the high faces too */
    for (int j = 0; j <= n; j++)
    for (int i = 0; i <= n; i++) {
      size_t ijk = (size_t)(i + g) + (size_t)(j + g) * jS + (size_t)(k + g) * kS;
      double x = hLevel * ((double)(i + B->low.i) + 0.5);       /* cell centre */
      double y = hLevel * ((double)(j + B->low.j) + 0.5);
      double z = hLevel * ((double)(k + B->low.k) + 0.5);
      Bi[ijk] = evaluateBeta(x - hLevel * 0.5, y, z, hLevel, 0, 1, 1);
      Bj[ijk] = evaluateBeta(x, y - hLevel * 0.5, z, hLevel, 1, 0, 1);
      Bk[ijk] = evaluateBeta(x, y, z - hLevel * 0.5, hLevel, 1, 1, 0);
      F[ijk]  = evaluateF(x, y, z, hLevel, 1, 1, 1);
    }
    for (int v = 0; v < 4; v++) hpgmg_upload_box_vector(L, box, ids[v], stage + v * vol);
#ifdef VECTOR_ALPHA
    /* A = 1.0 on the same index range (problem.fv.c:118,129-131); stage[0] is free again */
    hpgmg_download_box_vector(L, box, VECTOR_ALPHA, stage);
    for (int k = 0; k <= n; k++) for (int j = 0; j <= n; j++) for (int i = 0; i <= n; i++)
      stage[(size_t)(i + g) + (size_t)(j + g) * jS + (size_t)(k + g) * kS] = 1.0;
    hpgmg_upload_box_vector(L, box, VECTOR_ALPHA, stage);
#endif
  }
  hpgmg_rt_sync();
  free(stage);
}
