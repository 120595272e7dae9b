/*
 * hpgmg_mg.h -- multigrid hierarchy and cycles.
 * Drop-in for /root/reference/finite-volume/source/mg.h:15-45 (same struct, same entry points).
 */
#ifndef HPGMG_B200_MG_H
#define HPGMG_B200_MG_H

#include "hpgmg_level.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MG_AGGLOMERATION_START
#define MG_AGGLOMERATION_START 8      /* boxes stop shrinking and start merging at 8^3 (mg.h:15-17) */
#endif
#ifndef MG_DEFAULT_BOTTOM_NORM
#define MG_DEFAULT_BOTTOM_NORM 1e-3   /* bottom-solver relative tolerance              (mg.h:18-20) */
#endif

typedef struct {
  int my_rank;
  int num_levels;
  level_type **levels;
  struct {
    double MGBuild;
    double MGSolve;
  } timers;
  int MGSolves_performed;
} mg_type;

void MGBuild(mg_type *all_grids, level_type *fine_grid, double a, double b, int minCoarseGridDim); /* mg.c:842  */
void MGSolve(mg_type *all_grids, int onLevel, int u_id, int F_id, double a, double b, double rtol);   /* mg.c:1168 */
void FMGSolve(mg_type *all_grids, int onLevel, int u_id, int F_id, double a, double b, double rtol);  /* mg.c:1237 */
void FMGSolve2(mg_type *all_grids, int onLevel, int u_id, int F_id, double a, double b, double rtol); /* mg.c:1348 */
void MGPCG(mg_type *all_grids, int onLevel, int x_id, int F_id, double a, double b, double rtol);     /* mg.c:1500 */
void MGVCycle(mg_type *all_grids, int e_id, int R_id, double a, double b, int level);                 /* mg.c:1135 */
void MGDestroy(mg_type *all_grids);                                                                   /* mg.c:1027 */
void MGPrintTiming(mg_type *all_grids, int fromLevel);                                                /* mg.c:54   */
void MGResetTimers(mg_type *all_grids);                                                               /* mg.c:166  */
void richardson_error(mg_type *all_grids, int levelh, int u_id);                                      /* mg.c:1113 */

#ifdef __cplusplus
}
#endif
#endif
