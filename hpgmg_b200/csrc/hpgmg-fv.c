/*
 * hpgmg-fv.c -- command-line driver of the B200 build: `hpgmg-fv <log2_box_dim> <target_boxes_per_rank>`.
 *
 * Same arguments, problem sizing, benchmark protocol and output lines as the reference driver
 * (/root/reference/finite-volume/source/hpgmg-fv.c: argument checks :152-204, box-count search
 * :184-197, setup :280-308, 3-size benchmark loop :316-345 with bench_hpgmg :50-99, Richardson
 * check :351-366).  It is a plain C program against include/ headers: it is also the proof that code
 * written for the reference API links against libhpgmg_b200.so unchanged.
 *
 * Extra options (after the two positional ones):  --cheby  --no-graphs  --solves N  --error-only
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include "hpgmg_b200.h"

#ifndef MAX_COARSE_DIM
#define MAX_COARSE_DIM 11
#endif
#define DYNAMIC_RANGE 3

static int g_min_solves = 10;

/* warm up with N solves, then time N solves (hpgmg-fv.c:50-99, no-MPI branch) */
static void bench_hpgmg(mg_type *all_grids, int onLevel, double a, double b, double rtol)
{
  for (int doTiming = 0; doTiming <= 1; doTiming++) {
    if (all_grids->levels[onLevel]->my_rank == 0) {
      if (doTiming == 0) fprintf(stdout, "\n\n===== Warming up by running %d solves ==========================================\n", g_min_solves);
      else               fprintf(stdout, "\n\n===== Running %d solves ========================================================\n", g_min_solves);
      fflush(stdout);
    }
    MGResetTimers(all_grids);
    for (int n = 0; n < g_min_solves; n++) {
      zero_vector(all_grids->levels[onLevel], VECTOR_U);
      FMGSolve(all_grids, onLevel, VECTOR_U, VECTOR_F, a, b, rtol);
    }
  }
}

int main(int argc, char **argv)
{
  int my_rank = 0, num_tasks = 1;
  int error_only = 0;
  if (argc < 3) {
    fprintf(stderr, "usage: ./hpgmg-fv  [log2_box_dim]  [target_boxes_per_rank]  [--cheby] [--no-graphs] [--solves N] [--error-only]\n");
    exit(0);
  }
  int log2_box_dim = atoi(argv[1]);
  int target_boxes_per_rank = atoi(argv[2]);
  for (int i = 3; i < argc; i++) {
    if (!strcmp(argv[i], "--cheby")) hpgmg_b200_set_smoother(HPGMG_SMOOTHER_CHEBY);
    else if (!strcmp(argv[i], "--no-graphs")) hpgmg_b200_use_graphs(0);
    else if (!strcmp(argv[i], "--error-only")) error_only = 1;
    else if (!strcmp(argv[i], "--solves") && i + 1 < argc) g_min_solves = atoi(argv[++i]);
    else { fprintf(stderr, "unrecognized option '%s'\n", argv[i]); exit(0); }
  }
  if (log2_box_dim > 9) { fprintf(stderr, "log2_box_dim must be less than 10\n"); exit(0); }
  if (log2_box_dim < 4) { fprintf(stderr, "log2_box_dim must be at least 4\n"); exit(0); }
  if (target_boxes_per_rank < 1) { fprintf(stderr, "target_boxes_per_rank must be at least 1\n"); exit(0); }

  /* largest cube of boxes that fits the target and whose side has an odd part <= MAX_COARSE_DIM */
  const int box_dim = 1 << log2_box_dim;
  const int64_t target_boxes = (int64_t)target_boxes_per_rank * (int64_t)num_tasks;
  int boxes_in_i = -1;
  for (int64_t bi = 1; bi < 1000; bi++) {
    if (bi * bi * bi > target_boxes) continue;
    int64_t odd = (int64_t)box_dim * bi;
    while ((odd % 2) == 0) odd /= 2;
    if (odd <= MAX_COARSE_DIM) boxes_in_i = (int)bi;
  }
  if (boxes_in_i < 1) { fprintf(stderr, "failed to find an acceptable problem size\n"); exit(0); }

  if (hpgmg_b200_init(0) != 0) { fprintf(stderr, "hpgmg-fv: no usable CUDA device\n"); return 1; }

  fprintf(stdout, "\n\n");
  fprintf(stdout, "********************************************************************************\n");
  fprintf(stdout, "***                            HPGMG-FV Benchmark                            ***\n");
  fprintf(stdout, "********************************************************************************\n");
  fprintf(stdout, "%d GPU tasks (%s)\n", num_tasks, hpgmg_b200_backend());
  fprintf(stdout, "\n\n===== Benchmark setup ==========================================================\n");

  const int bc = BC_DIRICHLET, minCoarseDim = 1;
  level_type level_h;
  create_level(&level_h, boxes_in_i, box_dim, stencil_get_radius(), VECTORS_RESERVED, bc, my_rank, num_tasks);
  const double a = 0.0, b = 1.0;
  fprintf(stdout, "  Creating Poisson (a=%f, b=%f) test problem\n", a, b);
  const double h = 1.0 / ((double)boxes_in_i * (double)box_dim);
  initialize_problem(&level_h, h, a, b);
  rebuild_operator(&level_h, NULL, a, b);

  mg_type MG_h;
  MGBuild(&MG_h, &level_h, a, b, minCoarseDim);

  const double rtol = 1e-10;
  if (!error_only) {
    double AverageSolveTime[DYNAMIC_RANGE];
    int levels_timed = 0;
    for (int l = 0; l < DYNAMIC_RANGE && l < MG_h.num_levels; l++) {
      if (l > 0) restriction(MG_h.levels[l], VECTOR_F, MG_h.levels[l - 1], VECTOR_F, RESTRICT_CELL);
      bench_hpgmg(&MG_h, l, a, b, rtol);
      AverageSolveTime[l] = (double)MG_h.timers.MGSolve / (double)MG_h.MGSolves_performed;
      fprintf(stdout, "\n\n===== Timing Breakdown =========================================================\n");
      MGPrintTiming(&MG_h, l);
      levels_timed++;
    }
    fprintf(stdout, "\n\n===== Performance Summary ======================================================\n");
    for (int l = 0; l < levels_timed; l++) {
      double DOF = (double)MG_h.levels[l]->dim.i * (double)MG_h.levels[l]->dim.j * (double)MG_h.levels[l]->dim.k;
      double seconds = AverageSolveTime[l];
      fprintf(stdout, "  h=%0.15e  DOF=%0.15e  time=%0.6f  DOF/s=%0.3e  MPI=%d  OMP=%d\n", MG_h.levels[l]->h, DOF, seconds, DOF / seconds, num_tasks, 1);
    }
  }

  fprintf(stdout, "\n\n===== Richardson error analysis ================================================\n");
  MGResetTimers(&MG_h);
  for (int l = 0; l < 3 && l < MG_h.num_levels; l++) {
    if (l > 0) restriction(MG_h.levels[l], VECTOR_F, MG_h.levels[l - 1], VECTOR_F, RESTRICT_CELL);
    zero_vector(MG_h.levels[l], VECTOR_U);
    FMGSolve(&MG_h, l, VECTOR_U, VECTOR_F, a, b, rtol);
  }
  richardson_error(&MG_h, 0, VECTOR_U);

  fprintf(stdout, "\n\n===== Deallocating memory ======================================================\n");
  MGDestroy(&MG_h);
  destroy_level(&level_h);
  fprintf(stdout, "\n\n===== Done =====================================================================\n");
  hpgmg_b200_finalize();
  return 0;
}
