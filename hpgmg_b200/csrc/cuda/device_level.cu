/*
 * device_level.cu -- the device mirror of a level: DLevel (slab geometry) and device copies of the
 * block lists the host built in level.c / mg.c (the same blockCopy_type records the reference
 * walks with OpenMP: level.c:367-465 BCs, :498-922 ghost exchange, mg.c:181-831 transfers).
 */
#include <string.h>
#include <vector>
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

static void upload_table(DTable *t, const void *host, int n, size_t item)
{
  t->items = NULL;
  t->n = n;
  if (n <= 0) return;
  CUDA_CHECK(cudaMalloc(&t->items, (size_t)n * item));
  CUDA_CHECK(cudaMemcpyAsync(t->items, host, (size_t)n * item, cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
}

/* expand the local ghost-exchange list and the BC list of every shape into per-cell / per-column records */
static void build_small_level_tables(level_type *level, hpgmg_device_level *D)
{
  const int jS = level->box_jStride, kS = level->box_kStride, n = level->box_dim;
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) {
    std::vector<CopyItem> copies;
    const blockCopy_type *xb = level->exchange_ghosts[s].blocks[1];
    for (int e = 0; e < level->exchange_ghosts[s].num_blocks[1]; e++) {
      const blockCopy_type &B = xb[e];
      for (int k = 0; k < B.dim.k; k++) for (int j = 0; j < B.dim.j; j++) for (int i = 0; i < B.dim.i; i++) {
        CopyItem c = { B.read.box, (B.read.i + i) + (B.read.j + j) * jS + (B.read.k + k) * kS,
                       B.write.box, (B.write.i + i) + (B.write.j + j) * jS + (B.write.k + k) * kS };
        copies.push_back(c);
      }
    }
    std::vector<BCItem> cols;
    std::vector<ZeroItem> zeros;
    const blockCopy_type *bb = level->boundary_condition.blocks[s];
    for (int e = 0; e < level->boundary_condition.num_blocks[s]; e++) {
      const blockCopy_type &B = bb[e];
      const int nrm[3] = { (B.subtype % 3) - 1, ((B.subtype % 9) / 3) - 1, (B.subtype / 9) - 1 };
      const int lo[3] = { B.read.i, B.read.j, B.read.k }, ext[3] = { B.dim.i, B.dim.j, B.dim.k }, st[3] = { 1, jS, kS };
      for (int k = 0; k < ext[2]; k++) for (int j = 0; j < ext[1]; j++) for (int i = 0; i < ext[0]; i++) {
        ZeroItem z = { B.read.box, (lo[0] + i) + (lo[1] + j) * jS + (lo[2] + k) * kS };
        zeros.push_back(z);
      }
      const int e0 = nrm[0] ? 1 : ext[0], e1 = nrm[1] ? 1 : ext[1], e2 = nrm[2] ? 1 : ext[2];
      for (int c = 0; c < e0 * e1 * e2; c++) {
        const int p[3] = { c % e0, (c / e0) % e1, c / (e0 * e1) };
        int ijk = 0;
        for (int a = 0; a < 3; a++) ijk += (nrm[a] ? (nrm[a] < 0 ? -1 : n) : p[a] + lo[a]) * st[a];
        BCItem it = { B.read.box, B.subtype, ijk, 0 };
        cols.push_back(it);
      }
    }
    upload_table(&D->copy_tab[s], copies.data(), (int)copies.size(), sizeof(CopyItem));
    upload_table(&D->bc_tab[s], cols.data(), (int)cols.size(), sizeof(BCItem));
    upload_table(&D->bczero_tab[s], zeros.data(), (int)zeros.size(), sizeof(ZeroItem));
  }
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
  if ((long)level->dim.i * level->dim.j * level->dim.k <= COARSE_MAX_CELLS) build_small_level_tables(level, D);
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
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) {
    if (D->copy_tab[s].items) CUDA_CHECK(cudaFree(D->copy_tab[s].items));
    if (D->bc_tab[s].items) CUDA_CHECK(cudaFree(D->bc_tab[s].items));
    if (D->bczero_tab[s].items) CUDA_CHECK(cudaFree(D->bczero_tab[s].items));
  }
  if (D->low) CUDA_CHECK(cudaFree(D->low));
  if (D->tiles) CUDA_CHECK(cudaFree(D->tiles));
  if (D->tile_partials) CUDA_CHECK(cudaFree(D->tile_partials));
  free(D);
  HPGMG_SET_DEV(level, NULL);
}
