/*
 * hpgmg_level.h -- level / box / block-list data model of the B200 build.
 *
 * This is the drop-in boundary for the reference's level.h
 * (/root/reference/finite-volume/source/level.h:65-215): same type names, same field names,
 * same field order and the same sizes for the no-MPI build (blockCopy_type 128 B,
 * communicator_type 104 B, box_type 56 B, level_type 1296 B), so a
 * caller written against the reference header compiles and links unchanged.
 *
 * What differs is where the bytes live:
 *   - box_type.vectors[id] and every communicator buffer are DEVICE pointers (HBM of the GPU
 *     this rank owns).  Host code may do pointer arithmetic on them but must never
 *     dereference them; use hpgmg_b200.h (hpgmg_download_box_vector & co) to move data.
 *   - level_type::fluxes (unused by the reference's default build) carries an opaque handle to the
 *     device-side mirror of the block lists, so the struct is byte-for-byte the reference's.
 */
#ifndef HPGMG_B200_LEVEL_H
#define HPGMG_B200_LEVEL_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boundary conditions of the (cubical) domain -- level.h:28-29 */
#define BC_PERIODIC  0
#define BC_DIRICHLET 1

/* which parts of a box's 26-neighbourhood a ghost exchange / BC touches -- level.h:32-35 */
#define STENCIL_SHAPE_BOX        0   /* faces + edges + corners */
#define STENCIL_SHAPE_STAR       1   /* faces only              */
#define STENCIL_SHAPE_NO_CORNERS 2   /* faces + edges           */
#define STENCIL_MAX_SHAPES       3

/* tiling of the flattened work lists -- level.h:37-47.  The CUDA kernels do not need the
 * tiles (a grid is launched per box) but the lists are kept bit-identical to the
 * reference's because "index mapping must be bit-exact" is part of the contract. */
#ifndef BLOCKCOPY_TILE_I
#define BLOCKCOPY_TILE_I 10000
#endif
#ifndef BLOCKCOPY_TILE_J
#define BLOCKCOPY_TILE_J 8
#endif
#ifndef BLOCKCOPY_TILE_K
#define BLOCKCOPY_TILE_K 8
#endif

/* padding of the unit-stride pencil (doubles) -- level.h:49-57.  Only JSTRIDE is ever applied
 * by the reference (level.c:935-938); KSTRIDE/VOLUME are defined for completeness. */
#ifndef BOX_ALIGN_JSTRIDE
#define BOX_ALIGN_JSTRIDE 4
#endif
#ifndef BOX_ALIGN_KSTRIDE
#define BOX_ALIGN_KSTRIDE 4
#endif
#ifndef BOX_ALIGN_VOLUME
#define BOX_ALIGN_VOLUME  4
#endif

/* memory layout is [box][vector][k][j][i] -- level.h:59-65 */
#define USE_BVKJI_LAYOUT

/* One 3-D copy / tile.  box<0 means "use ptr" (a communication buffer).  -- level.h:65-73 */
typedef struct {
  int subtype;                                   /* BC lists: domain normal 0..26          */
  struct { int i, j, k; } dim;                   /* extent of the block                    */
  struct { int box, i, j, k, jStride, kStride; double *ptr; } read, write;
} __attribute__((aligned(64))) blockCopy_type;

/* A "mini program": pack list, local list, unpack list + neighbour buffers. -- level.h:77-93 */
typedef struct {
  int       num_recvs;
  int       num_sends;
  int      *recv_ranks;
  int      *send_ranks;
  int      *recv_sizes;                          /* in doubles                             */
  int      *send_sizes;
  double  **recv_buffers;                        /* DEVICE pointers                        */
  double  **send_buffers;                        /* DEVICE pointers (may be peer-mapped)   */
  int       allocated_blocks[3];
  int       num_blocks[3];                       /* [0]=pack [1]=local [2]=unpack          */
  blockCopy_type *blocks[3];
} communicator_type;

/* One cubical box of a level. -- level.h:97-108 */
typedef struct {
  int global_box_id;                             /* index into level->rank_of_box          */
  struct { int i, j, k; } low;                   /* global coordinate of cell (0,0,0)      */
  int dim;                                       /* owned cells per side                   */
  int ghosts;                                    /* ghost depth                            */
  int jStride, kStride, volume;                  /* in doubles, ghosts + padding included  */
  int numVectors;
  double **vectors;                              /* host array of DEVICE pointers          */
  double  *fp_base;                              /* DEVICE base of this box's 4-D array    */
} box_type;

/* One multigrid level. -- level.h:112-200 */
typedef struct {
  double h;                                      /* grid spacing                           */
  int active;                                    /* this rank has work here or below       */
  int num_ranks;
  int my_rank;
  int box_dim;
  int box_ghosts;
  int box_jStride, box_kStride, box_volume;
  int numVectors;
  int tag;                                       /* log2(dim): unique per level            */
  struct { int i, j, k; } boxes_in;
  struct { int i, j, k; } dim;

  int      *rank_of_box;                         /* [k][j][i] owner of every box           */
  int       num_my_boxes;
  box_type *my_boxes;

  int             allocated_blocks;
  int             num_my_blocks;
  blockCopy_type *my_blocks;                     /* boxes flattened into compute tiles     */

  struct {
    int type;                                    /* BC_PERIODIC | BC_DIRICHLET             */
    int allocated_blocks[STENCIL_MAX_SHAPES];
    int       num_blocks[STENCIL_MAX_SHAPES];
    blockCopy_type *blocks[STENCIL_MAX_SHAPES];
  } boundary_condition;

  communicator_type exchange_ghosts[STENCIL_MAX_SHAPES];
  communicator_type restriction[4];              /* cell, face-i, face-j, face-k           */
  communicator_type interpolation;

  double dominant_eigenvalue_of_DinvA;
  int    must_subtract_mean;
  double *RedBlack_base;                         /* kept for layout parity; unused on GPU  */
  double *RedBlack_FP;
  double *fluxes;                                /* reference: scratch of the experimental flux smoother;
                                                    here: opaque handle of the device mirror (lists, geometry) */

  int num_threads;

  struct {                                       /* seconds, accumulated per operator class */
    double smooth, apply_op, residual, blas1, blas3, boundary_conditions;
    double restriction_total, restriction_pack, restriction_local, restriction_unpack,
           restriction_recv, restriction_send, restriction_wait;
    double interpolation_total, interpolation_pack, interpolation_local, interpolation_unpack,
           interpolation_recv, interpolation_send, interpolation_wait;
    double ghostZone_total, ghostZone_pack, ghostZone_local, ghostZone_unpack,
           ghostZone_recv, ghostZone_send, ghostZone_wait;
    double collectives;
    double Total;
  } timers;
  int Krylov_iterations;
  int CAKrylov_formations_of_G;
  int vcycles_from_this_level;
} level_type;

/* level.c:1075,1305,929,1265,95,313 */
void create_level(level_type *level, int boxes_in_i, int box_dim, int box_ghosts, int numVectors,
                  int domain_boundary_condition, int my_rank, int num_ranks);
void destroy_level(level_type *level);
void create_vectors(level_type *level, int numVectors);
void reset_level_timers(level_type *level);
int  qsortInt(const void *a, const void *b);
void append_block_to_list(blockCopy_type **blocks, int *allocated_blocks, int *num_blocks,
                          int dim_i, int dim_j, int dim_k,
                          int read_box,  double *read_ptr,  int read_i,  int read_j,  int read_k,
                          int read_jStride,  int read_kStride,  int read_scale,
                          int write_box, double *write_ptr, int write_i, int write_j, int write_k,
                          int write_jStride, int write_kStride, int write_scale,
                          int my_blockcopy_tile_i, int my_blockcopy_tile_j, int my_blockcopy_tile_k,
                          int subtype);

/* Aligned allocation hooks of the reference (level.c:24-40).  Here they hand out DEVICE memory
 * (zero-filled), because every caller in level.c / mg.c uses them for vector or buffer data. */
void *MALLOC(size_t size);
void  FREE(void *ptr);

#ifdef __cplusplus
}
#endif
#endif
