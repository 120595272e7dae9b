/*
 * ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A 60-line harness around the UNMODIFIED reference (linked from oracle/_ref/libhpgmg_ref.so, built
 * by oracle/Makefile from /root/reference/finite-volume/source): it performs the reference driver's
 * setup (hpgmg-fv.c:280-308) and then W warm-up + K timed `zero_vector(U); FMGSolve(...)` steps
 * (hpgmg-fv.c:78-80) on the host's cores with OpenMP, and prints one line
 *     REF dof=<dim^3> seconds_per_solve=<t> norm=<||r||> rel=<||r||/||f||> threads=<n>
 * bench.py uses it for `cpu_baseline` and for `--impl reference`.  Compiled against the
 * reference's own headers (-I$(REF)), so no struct layout is assumed here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <unistd.h>
#include <omp.h>
#include "defines.h"
#include "level.h"
#include "operators.h"
#include "mg.h"
#include "solvers.h"

int main(int argc, char **argv)
{
  if (argc < 5) { fprintf(stderr, "usage: ref_bench log2_box_dim boxes warmup steps [level]\n"); return 2; }
  int log2_box_dim = atoi(argv[1]), target = atoi(argv[2]), warmup = atoi(argv[3]), steps = atoi(argv[4]);
  int onLevel = argc > 5 ? atoi(argv[5]) : 0;
  int box_dim = 1 << log2_box_dim, boxes_in_i = -1;
  for (int64_t bi = 1; bi < 1000; bi++) if (bi * bi * bi <= target) {
    int64_t c = box_dim * bi; while ((c % 2) == 0) c /= 2;
    if (c <= 11) boxes_in_i = (int)bi;
  }
  if (boxes_in_i < 1) return 3;
  /* the reference prints its progress on stdout: park stdout on /dev/null while it runs */
  fflush(stdout);
  int saved = dup(1);
  FILE *devnull = freopen("/dev/null", "w", stdout);
  (void)devnull;
  level_type level_h;
  create_level(&level_h, boxes_in_i, box_dim, stencil_get_radius(), VECTORS_RESERVED, BC_DIRICHLET, 0, 1);
  double a = 0.0, b = 1.0, h = 1.0 / ((double)boxes_in_i * (double)box_dim);
  initialize_problem(&level_h, h, a, b);
  rebuild_operator(&level_h, NULL, a, b);
  mg_type MG;
  MGBuild(&MG, &level_h, a, b, 1);
  for (int l = 1; l <= onLevel; l++) restriction(MG.levels[l], VECTOR_F, MG.levels[l - 1], VECTOR_F, RESTRICT_CELL);
  for (int n = 0; n < warmup; n++) { zero_vector(MG.levels[onLevel], VECTOR_U); FMGSolve(&MG, onLevel, VECTOR_U, VECTOR_F, a, b, 1e-10); }
  double t0 = omp_get_wtime();
  for (int n = 0; n < steps; n++) { zero_vector(MG.levels[onLevel], VECTOR_U); FMGSolve(&MG, onLevel, VECTOR_U, VECTOR_F, a, b, 1e-10); }
  double t1 = omp_get_wtime();
  residual(MG.levels[onLevel], VECTOR_TEMP, VECTOR_U, VECTOR_F, a, b);
  double nr = norm(MG.levels[onLevel], VECTOR_TEMP), nf = norm(MG.levels[onLevel], VECTOR_F);
  fflush(stdout);
  dup2(saved, 1);
  FILE *out = fdopen(saved, "w");
  double dof = (double)MG.levels[onLevel]->dim.i * MG.levels[onLevel]->dim.j * MG.levels[onLevel]->dim.k;
  fprintf(out, "REF dof=%.0f seconds_per_solve=%.9f norm=%.15e rel=%.15e threads=%d\n", dof, steps > 0 ? (t1 - t0) / steps : 0.0, nr, nr / nf, omp_get_max_threads());
  fflush(out);
  return 0;
}
