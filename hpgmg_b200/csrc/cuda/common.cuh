/*
 * common.cuh -- what every CUDA translation unit of the library shares: the compute stream, the
 * launch counter, error checking and the by-value description of a level's HBM layout.
 *
 * There is no host fallback: every kernel is a __global__ function for sm_100a launched on the
 * library's stream; if no device is present hpgmg_b200_init() fails and nothing else works.
 */
#ifndef HPGMG_B200_COMMON_CUH
#define HPGMG_B200_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

#include "hpgmg_b200.h"
#include "../runtime.h"

extern cudaStream_t g_stream;                   /* compute stream (runtime.cu)                    */
extern unsigned long long g_launches;           /* kernels enqueued; graph replays add their nodes */
extern int g_capturing;                         /* inside a stream capture?                        */

void hpgmg_cuda_check(cudaError_t e, const char *what, const char *file, int line);
#define CUDA_CHECK(x) hpgmg_cuda_check((x), #x, __FILE__, __LINE__)

/* launch on the compute stream and count it */
void hpgmg_refuse_launch(const char *kernel);
#define LAUNCH(kernel, grid, block, smem, ...)                                  \
  do {                                                                          \
    if (hpgmg_rt_layout_only()) hpgmg_refuse_launch(#kernel);                   \
    kernel<<<(grid), (block), (smem), g_stream>>>(__VA_ARGS__);                 \
    g_launches++;                                                               \
    CUDA_CHECK(cudaGetLastError());                                             \
  } while (0)

/* One level's HBM layout, passed by value to kernels: a single slab [box][vector][k][j][i], every
 * box the same padded cube (reference level.c:935-938).  vec(b,id) points at cell (0,0,0). */
struct DLevel {
  double *base;
  int nboxes, nvec;
  int dim, ghosts;
  int jStride, kStride, volume;
  int origin;                                   /* ghosts*(1+jStride+kStride)                      */
  __host__ __device__ __forceinline__ double *vec(int box, int id) const
  {
    return base + ((size_t)box * (size_t)nvec + (size_t)id) * (size_t)volume + origin;
  }
};

/* a device copy of one blockCopy_type list */
struct DList {
  blockCopy_type *blocks;
  int n;
};

/* Flat work tables of a level's ghost fill (built on the host from the reference-identical block lists,
 * device_level.cu): one record per ghost cell copied inside this GPU and one per BC column.  Offsets
 * are in doubles from (DLevel::base + id*volume), i.e. they already contain box and cell.
 * A BC column whose tangential coordinates lie in a ghost region reads the image of that region in the
 * LOCAL neighbour box (src != dst), so copies and BCs have no mutual dependency and run as one kernel;
 * columns that depend on a box owned by another GPU are "late": they read the box's own ghost cells
 * after the unpack. */
#define COARSE_MAX_CELLS 4096
struct FillCopy { int src, dst; };
struct FillBC   { int dst, src, subtype, pad; };
struct FillTable {
  FillCopy *copies;  int ncopies;
  FillBC   *bc;      int nbc;         /* local-source columns */
  FillBC   *late;    int nlate;       /* columns that need data from another GPU first */
};

/* device mirror hanging off level_type::fluxes */
struct hpgmg_device_level {
  DLevel L;
  int   *low;                                   /* device [nboxes][3]: global coords of cell 0,0,0 */
  DList  bc[STENCIL_MAX_SHAPES];
  DList  exchange[STENCIL_MAX_SHAPES][3];
  DList  restriction[4][3];
  DList  interpolation[3];
  FillTable fill[STENCIL_MAX_SHAPES];
  int fill_nvec;                                /* numVectors the fill offsets were built for */
  double *tile_partials;                        /* scratch for dot/mean: one double per compute tile */
  blockCopy_type *tiles;                        /* device copy of level->my_blocks                  */
  int     ntiles;
};

static inline const DLevel &dl_of(const level_type *level) { return HPGMG_DEV(level)->L; }

/* max over non-negative doubles through their bit pattern (IEEE order == unsigned integer order) */
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
  atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

#endif
