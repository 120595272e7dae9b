/*
 * device_level.cu -- the device mirror of a level: DLevel (slab geometry) and device copies of the
 * block lists the host built in level.c / mg.c (the same blockCopy_type records the reference
 * walks with OpenMP: level.c:367-465 BCs, :498-922 ghost exchange, mg.c:181-831 transfers).
 */
#include <string.h>
#include "common.cuh"

static void upload_list(DList *dst, const blockCopy_type *src, int n)
{
  if (dst->blocks && !hpgmg_rt_layout_only()) { CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaFree(dst->blocks)); }
  dst->blocks = NULL;
  dst->n = n;
  if (n <= 0 || hpgmg_rt_layout_only()) return;
  CUDA_CHECK(cudaMalloc(&dst->blocks, (size_t)n * sizeof(blockCopy_type)));
  CUDA_CHECK(cudaMemcpyAsync(dst->blocks, src, (size_t)n * sizeof(blockCopy_type), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
}

static void free_list(DList *l)
{
  if (l->blocks) CUDA_CHECK(cudaFree(l->blocks));
  l->blocks = NULL;
  l->n = 0;
}

extern "C" void hpgmg_device_level_rebind_vectors(level_type *level)
{
  hpgmg_device_level *D = HPGMG_DEV(level);
  DLevel &L = D->L;
  L.nboxes = level->num_my_boxes;
  L.nvec = level->numVectors;
  L.dim = level->box_dim;
  L.ghosts = level->box_ghosts;
  L.jStride = level->box_jStride;
  L.kStride = level->box_kStride;
  L.volume = level->box_volume;
  L.origin = level->box_ghosts * (1 + level->box_jStride + level->box_kStride);
  L.base = (level->num_my_boxes > 0) ? level->my_boxes[0].vectors[0] : NULL;
}

extern "C" void hpgmg_device_level_create(level_type *level)
{
  hpgmg_device_level *D = (hpgmg_device_level *)calloc(1, sizeof(hpgmg_device_level));
  HPGMG_SET_DEV(level, D);
  hpgmg_device_level_rebind_vectors(level);

  if (hpgmg_rt_layout_only()) return;
  const int nb = level->num_my_boxes;
  if (nb > 0) {
    int *low = (int *)malloc((size_t)nb * 3 * sizeof(int));
    for (int b = 0; b < nb; b++) {
      low[3 * b + 0] = level->my_boxes[b].low.i;
      low[3 * b + 1] = level->my_boxes[b].low.j;
      low[3 * b + 2] = level->my_boxes[b].low.k;
    }
    CUDA_CHECK(cudaMalloc(&D->low, (size_t)nb * 3 * sizeof(int)));
    CUDA_CHECK(cudaMemcpyAsync(D->low, low, (size_t)nb * 3 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    free(low);
  }
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) {
    upload_list(&D->bc[s], level->boundary_condition.blocks[s], level->boundary_condition.num_blocks[s]);
    for (int p = 0; p < 3; p++)
      upload_list(&D->exchange[s][p], level->exchange_ghosts[s].blocks[p], level->exchange_ghosts[s].num_blocks[p]);
  }
  D->ntiles = level->num_my_blocks;
  if (D->ntiles > 0) {
    CUDA_CHECK(cudaMalloc(&D->tiles, (size_t)D->ntiles * sizeof(blockCopy_type)));
    CUDA_CHECK(cudaMemcpyAsync(D->tiles, level->my_blocks, (size_t)D->ntiles * sizeof(blockCopy_type), cudaMemcpyHostToDevice, g_stream));
    CUDA_CHECK(cudaMalloc(&D->tile_partials, (size_t)D->ntiles * sizeof(double)));
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
  }
}

extern "C" void hpgmg_device_level_upload_transfer_lists(level_type *level)
{
  hpgmg_device_level *D = HPGMG_DEV(level);
  for (int t = 0; t < 4; t++)
    for (int p = 0; p < 3; p++)
      upload_list(&D->restriction[t][p], level->restriction[t].blocks[p], level->restriction[t].num_blocks[p]);
  for (int p = 0; p < 3; p++)
    upload_list(&D->interpolation[p], level->interpolation.blocks[p], level->interpolation.num_blocks[p]);
}

extern "C" void hpgmg_device_level_destroy(level_type *level)
{
  hpgmg_device_level *D = HPGMG_DEV(level);
  if (!D) return;
  if (hpgmg_rt_layout_only()) { free(D); HPGMG_SET_DEV(level, NULL); return; }
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) {
    free_list(&D->bc[s]);
    for (int p = 0; p < 3; p++) free_list(&D->exchange[s][p]);
  }
  for (int t = 0; t < 4; t++) for (int p = 0; p < 3; p++) free_list(&D->restriction[t][p]);
  for (int p = 0; p < 3; p++) free_list(&D->interpolation[p]);
  if (D->low) CUDA_CHECK(cudaFree(D->low));
  if (D->tiles) CUDA_CHECK(cudaFree(D->tiles));
  if (D->tile_partials) CUDA_CHECK(cudaFree(D->tile_partials));
  free(D);
  HPGMG_SET_DEV(level, NULL);
}
