/*
 * hpgmg_defines.h -- vector-id map of a level (drop-in for the reference's defines.h).
 *
 * Replaces: /root/reference/finite-volume/source/defines.h:28-38 (default, non-Helmholtz map).
 * Every operator addresses a vector by integer id; storage is level->my_boxes[b].vectors[id].
 * Operator: L u = a*alpha*u - b*div( beta grad u ); the benchmark is Poisson (a=0, b=1) so
 * there is no VECTOR_ALPHA / VECTOR_L1INV in the default map; -DUSE_HELMHOLTZ selects the reference's
 * Helmholtz map (defines.h:12-26), as it does in the reference.  They are macros, not an enum,
 * because reference callers test them with #ifdef.
 */
#ifndef HPGMG_B200_DEFINES_H
#define HPGMG_B200_DEFINES_H

#define VECTOR_TEMP      0  /* scratch; clobbered by smooth()                              */
#define VECTOR_U         1  /* numerical solution                                          */
#define VECTOR_F         2  /* right-hand side, cell averaged                              */
#define VECTOR_E         3  /* error / correction                                          */
#define VECTOR_R         4  /* residual f-Av (also the RHS of every coarse level)          */
#define VECTOR_DINV      5  /* inverse of the operator's diagonal                          */
#define VECTOR_BETA_I    6  /* face coefficient; element ijk is the LOW i-face of cell ijk */
#define VECTOR_BETA_J    7
#define VECTOR_BETA_K    8
#if defined(USE_HELMHOLTZ) || defined(USE_L1JACOBI)
/* the reference's research map (defines.h:12-26): the library built with -DUSE_HELMHOLTZ (libhpgmg_b200_helmholtz.so) */
#define VECTOR_ALPHA     9  /* cell-centred coefficient of the a*alpha*u term               */
#define VECTOR_L1INV    10  /* inverse of the L1 norm of each row (written by rebuild_operator_blackbox) */
#define VECTORS_RESERVED 11
#else
#define VECTORS_RESERVED 9  /* first id available to the bottom solver's Krylov vectors    */
#endif

#endif
