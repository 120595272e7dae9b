/*
 * hpgmg_solvers.h -- bottom solver.  Drop-in for /root/reference/finite-volume/source/solvers.h:9-10.
 * The only solver built is the diagonally preconditioned BiCGStab of solvers/bicgstab.c:14-97
 * (the reference's default, -DUSE_BICGSTAB).
 */
#ifndef HPGMG_B200_SOLVERS_H
#define HPGMG_B200_SOLVERS_H
#include "hpgmg_level.h"
#ifdef __cplusplus
extern "C" {
#endif
void IterativeSolver(level_type *level, int u_id, int f_id, double a, double b, double desired_reduction_in_norm);
int  IterativeSolver_NumVectors(void);
#ifdef __cplusplus
}
#endif
#endif
