/*
 * level.c -- construction of one multigrid level for the B200 build.
 *
 * Restates the behaviour of /root/reference/finite-volume/source/level.c (create_level :1075,
 * create_vectors :929, decompose_level_zmort :240, append_block_to_list :313,
 * build_boundary_conditions :367, build_exchange_ghosts :498, destroy_level :1305) with the
 * SAME list contents in the SAME order -- the block lists are the index mapping the contract
 * requires to be bit-exact (tests/test_lists.py diffs them against the reference build).
 *
 * Differences by design:
 *   - all floating-point storage (box vectors, exchange buffers) is device memory obtained from
 *     the runtime; the host only does pointer arithmetic on it;
 *   - one slab per level instead of one malloc per box: box b, vector v lives at
 *     slab + (b*numVectors + v)*volume, the reference's [box][vector][k][j][i] order
 *     (level.c:954-995), so a whole level is one contiguous, 128-byte aligned HBM range;
 *   - the finished lists are mirrored to the GPU once (hpgmg_device_level_create).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "hpgmg_b200.h"
#include "runtime.h"

/* ------------------------------------------------------------------------------------------ */
int qsortInt(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* direction d in 0..26 <-> (di,dj,dk) in {-1,0,1}^3, d = 13 + di + 3 dj + 9 dk (level.c:382) */
static inline void dir_to_offsets(int d, int *di, int *dj, int *dk)
{
  *di = (d % 3) - 1;
  *dj = ((d % 9) / 3) - 1;
  *dk = (d / 9) - 1;
}
static inline int nonzero_count(int di, int dj, int dk) { return (di != 0) + (dj != 0) + (dk != 0); }

/* does `shape` communicate / enforce BCs in direction d?  faces always; edges unless STAR;
 * corners only for BOX (level.c:526-531, :420-424). */
static int shape_uses_dir(int shape, int d)
{
  int di, dj, dk;
  dir_to_offsets(d, &di, &dj, &dk);
  int n = nonzero_count(di, dj, dk);
  if (n == 0) return 0;
  if (shape == STENCIL_SHAPE_STAR) return n == 1;
  if (shape == STENCIL_SHAPE_NO_CORNERS) return n <= 2;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Z-Morton assignment over a possibly odd-sized cube of boxes: recursive octant split with the
 * low half = floor(n/2) (level.c:240-275).  Returns the advanced curve offset. */
static int zmort_assign(int *rank_of_box, int bi, int bj, int bk, int ilo, int jlo, int klo,
                        int idim, int jdim, int kdim, int ranks, int offset, int curve_len)
{
  if (idim < 1 || jdim < 1 || kdim < 1) return offset;
  if (ilo < 0 || jlo < 0 || klo < 0) return offset;
  if (idim == 1 && jdim == 1 && kdim == 1) {
    if (ilo < bi && jlo < bj && klo < bk) {
      int b = ilo + jlo * bi + klo * bi * bj;
      rank_of_box[b] = (int)(((uint64_t)ranks * (uint64_t)offset) / (uint64_t)curve_len);
      return offset + 1;
    }
    return offset;
  }
  const int il = idim / 2, jl = jdim / 2, kl = kdim / 2;
  for (int oct = 0; oct < 8; oct++) {          /* i fastest, then j, then k: the Z order */
    int hi_i = oct & 1, hi_j = (oct >> 1) & 1, hi_k = (oct >> 2) & 1;
    offset = zmort_assign(rank_of_box, bi, bj, bk,
                          hi_i ? ilo + il : ilo, hi_j ? jlo + jl : jlo, hi_k ? klo + kl : klo,
                          hi_i ? idim - il : il, hi_j ? jdim - jl : jl, hi_k ? kdim - kl : kl,
                          ranks, offset, curve_len);
  }
  return offset;
}

/* ------------------------------------------------------------------------------------------ */
/* Append a 3-D region to a block list, cut into tiles of at most tile_i x tile_j x tile_k
 * (k outermost, i innermost).  read/write_scale convert tile offsets between index spaces of
 * different resolution: 1/1 ghost copy, 2/1 restriction, 1/2 interpolation (level.c:313-361). */
void append_block_to_list(blockCopy_type **blocks, int *allocated_blocks, int *num_blocks,
                          int dim_i, int dim_j, int dim_k,
                          int read_box, double *read_ptr, int read_i, int read_j, int read_k,
                          int read_jStride, int read_kStride, int read_scale,
                          int write_box, double *write_ptr, int write_i, int write_j, int write_k,
                          int write_jStride, int write_kStride, int write_scale,
                          int tile_i, int tile_j, int tile_k, int subtype)
{
  for (int kk = 0; kk < dim_k; kk += tile_k)
  for (int jj = 0; jj < dim_j; jj += tile_j)
  for (int ii = 0; ii < dim_i; ii += tile_i) {
    if (*num_blocks >= *allocated_blocks) {
      int want = (*allocated_blocks == 0) ? 1000 : 2 * (*allocated_blocks);
      blockCopy_type *grown = NULL;
      /* blockCopy_type is 64-byte aligned: allocate accordingly and move by hand */
      if (posix_memalign((void **)&grown, 64, (size_t)want * sizeof(blockCopy_type)) != 0 || !grown) {
        fprintf(stderr, "append_block_to_list: cannot grow list (%d -> %d)\n", *allocated_blocks, want);
        exit(0);
      }
      if (*num_blocks > 0) memcpy(grown, *blocks, (size_t)(*num_blocks) * sizeof(blockCopy_type));
      free(*blocks);
      *blocks = grown;
      *allocated_blocks = want;
    }
    blockCopy_type *t = &(*blocks)[*num_blocks];
    memset(t, 0, sizeof(*t));
    t->subtype = subtype;
    t->dim.i = (dim_i - ii < tile_i) ? dim_i - ii : tile_i;
    t->dim.j = (dim_j - jj < tile_j) ? dim_j - jj : tile_j;
    t->dim.k = (dim_k - kk < tile_k) ? dim_k - kk : tile_k;
    t->read.box = read_box;    t->read.ptr = read_ptr;
    t->read.i = read_i + read_scale * ii;
    t->read.j = read_j + read_scale * jj;
    t->read.k = read_k + read_scale * kk;
    t->read.jStride = read_jStride;   t->read.kStride = read_kStride;
    t->write.box = write_box;  t->write.ptr = write_ptr;
    t->write.i = write_i + write_scale * ii;
    t->write.j = write_j + write_scale * jj;
    t->write.k = write_k + write_scale * kk;
    t->write.jStride = write_jStride; t->write.kStride = write_kStride;
    (*num_blocks)++;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* box-grid helpers */
static inline void box_grid_coords(const level_type *L, int box, int *bi, int *bj, int *bk)
{
  *bi = L->my_boxes[box].low.i / L->box_dim;
  *bj = L->my_boxes[box].low.j / L->box_dim;
  *bk = L->my_boxes[box].low.k / L->box_dim;
}

/* global id of the box adjacent to (bi,bj,bk) in direction (di,dj,dk); -1 if that is outside a
 * non-periodic domain (level.c:558-573). */
static int adjacent_box_id(const level_type *L, int bi, int bj, int bk, int di, int dj, int dk)
{
  int ni = bi + di, nj = bj + dj, nk = bk + dk;
  if (L->boundary_condition.type == BC_PERIODIC) {
    ni = (ni + L->boxes_in.i) % L->boxes_in.i;
    nj = (nj + L->boxes_in.j) % L->boxes_in.j;
    nk = (nk + L->boxes_in.k) % L->boxes_in.k;
  } else if (ni < 0 || nj < 0 || nk < 0 || ni >= L->boxes_in.i || nj >= L->boxes_in.j || nk >= L->boxes_in.k) {
    return -1;
  }
  return ni + nj * L->boxes_in.i + nk * L->boxes_in.i * L->boxes_in.j;
}

static int local_index_of_box(const level_type *L, int global_id)
{
  for (int b = 0; b < L->num_my_boxes; b++)
    if (L->my_boxes[b].global_box_id == global_id) return b;
  return -1;
}

/* Along one axis, a message travelling in direction d (relative to the SENDER) leaves from
 * `send_lo`, has `extent` cells and lands at `recv_lo` in the receiver (level.c:645-659). */
static inline void axis_geometry(int d, int box_dim, int ghosts, int *send_lo, int *extent, int *recv_lo)
{
  if (d < 0)      { *send_lo = 0;                *extent = ghosts;  *recv_lo = box_dim; }
  else if (d > 0) { *send_lo = box_dim - ghosts; *extent = ghosts;  *recv_lo = -ghosts; }
  else            { *send_lo = 0;                *extent = box_dim; *recv_lo = 0;       }
}

/* ------------------------------------------------------------------------------------------ */
/* Boundary-condition work list for `shape`: every ghost region of one of my boxes that lies
 * outside the domain.  The region is addressed by the box-relative direction, the subtype
 * stores the normal to the DOMAIN (a box corner is not necessarily a domain corner).
 * Tiles are 16x16 in j,k (level.c:367-465). */
static void build_boundary_conditions(level_type *L, int shape)
{
  L->boundary_condition.blocks[shape] = NULL;
  L->boundary_condition.num_blocks[shape] = 0;
  L->boundary_condition.allocated_blocks[shape] = 0;
  if (L->boundary_condition.type == BC_PERIODIC) return;

  const int g = L->box_ghosts, n = L->box_dim;
  const int tile_i = (BLOCKCOPY_TILE_I < g) ? g : BLOCKCOPY_TILE_I;
  const int tile_j = (16 < g) ? g : 16;
  const int tile_k = (16 < g) ? g : 16;

  for (int box = 0; box < L->num_my_boxes; box++) {
    int bi, bj, bk;
    box_grid_coords(L, box, &bi, &bj, &bk);
    for (int d = 0; d < 27; d++) {
      int di, dj, dk;
      dir_to_offsets(d, &di, &dj, &dk);
      int normal = 13, outside = 0;
      if (bi + di < 0)              { outside = 1; normal -= 1; }
      if (bj + dj < 0)              { outside = 1; normal -= 3; }
      if (bk + dk < 0)              { outside = 1; normal -= 9; }
      if (bi + di >= L->boxes_in.i) { outside = 1; normal += 1; }
      if (bj + dj >= L->boxes_in.j) { outside = 1; normal += 3; }
      if (bk + dk >= L->boxes_in.k) { outside = 1; normal += 9; }
      if (!outside) continue;
      if (shape != STENCIL_SHAPE_BOX && !shape_uses_dir(shape, d)) continue;

      int lo[3], ext[3];
      const int dd[3] = { di, dj, dk };
      for (int a = 0; a < 3; a++) {
        if (dd[a] < 0)      { ext[a] = g; lo[a] = -g; }
        else if (dd[a] > 0) { ext[a] = g; lo[a] = n;  }
        else                { ext[a] = n; lo[a] = 0;  }
      }
      append_block_to_list(&L->boundary_condition.blocks[shape], &L->boundary_condition.allocated_blocks[shape],
                           &L->boundary_condition.num_blocks[shape],
                           ext[0], ext[1], ext[2],
                           box, NULL, lo[0], lo[1], lo[2], L->my_boxes[box].jStride, L->my_boxes[box].kStride, 1,
                           box, NULL, lo[0], lo[1], lo[2], L->my_boxes[box].jStride, L->my_boxes[box].kStride, 1,
                           tile_i, tile_j, tile_k, normal);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* One ghost message: a (box, direction) pair and who is on either end. */
typedef struct {
  int sendRank, sendBoxID, sendBox, sendDir;
  int recvRank, recvBoxID, recvBox;
} ghost_msg;

/* buffers are ordered by (sendRank, sendBoxID, sendDir) on both ends so that the k-th double a
 * sender packs is the k-th double the receiver unpacks (level.c:79-92). */
static int cmp_ghost_msg(const void *pa, const void *pb)
{
  const ghost_msg *a = (const ghost_msg *)pa, *b = (const ghost_msg *)pb;
  if (a->sendRank  != b->sendRank)  return (a->sendRank  > b->sendRank)  ? 1 : -1;
  if (a->sendBoxID != b->sendBoxID) return (a->sendBoxID > b->sendBoxID) ? 1 : -1;
  if (a->sendDir   != b->sendDir)   return (a->sendDir   > b->sendDir)   ? 1 : -1;
  return 0;
}

/* sort + unique a list of ranks in place, return the new length */
static int sorted_unique(int *v, int n)
{
  qsort(v, n, sizeof(int), qsortInt);
  int m = 0;
  for (int i = 0; i < n; i++)
    if (m == 0 || v[i] != v[m - 1]) v[m++] = v[i];
  return m;
}

static int find_rank(const int *ranks, int n, int r)
{
  for (int i = 0; i < n; i++) if (ranks[i] == r) return i;
  return -1;
}

static void reset_communicator(communicator_type *c)
{
  memset(c, 0, sizeof(*c));
}

/* Ghost-zone exchange "mini program" for `shape` (level.c:498-922): pack list (my cells ->
 * send buffers), local list (my cells -> my ghosts), unpack list (recv buffers -> my ghosts).
 * Tiles are 16x16 in j,k. */
static void build_exchange_ghosts(level_type *L, int shape)
{
  communicator_type *C = &L->exchange_ghosts[shape];
  reset_communicator(C);

  const int n = L->box_dim, g = L->box_ghosts;
  const int maxMsgs = 26 * (L->num_my_boxes > 0 ? L->num_my_boxes : 1);

  /* ---- what I send ------------------------------------------------------------------ */
  ghost_msg *out = (ghost_msg *)malloc((size_t)maxMsgs * sizeof(ghost_msg));
  int *outRanks  = (int *)malloc((size_t)maxMsgs * sizeof(int));
  int nOut = 0, nOutRemote = 0;
  for (int box = 0; box < L->num_my_boxes; box++) {
    int bi, bj, bk;
    box_grid_coords(L, box, &bi, &bj, &bk);
    for (int d = 0; d < 27; d++) {
      if (!shape_uses_dir(shape, d)) continue;
      int di, dj, dk;
      dir_to_offsets(d, &di, &dj, &dk);
      int nb = adjacent_box_id(L, bi, bj, bk, di, dj, dk);
      if (nb < 0 || L->rank_of_box[nb] == -1) continue;
      ghost_msg *m = &out[nOut++];
      m->sendRank = L->my_rank;            m->sendBoxID = L->my_boxes[box].global_box_id;
      m->sendBox  = box;                   m->sendDir   = d;
      m->recvRank = L->rank_of_box[nb];    m->recvBoxID = nb;
      m->recvBox  = -1;
      if (m->recvRank != L->my_rank) outRanks[nOutRemote++] = m->recvRank;
      else                           m->recvBox = local_index_of_box(L, nb);
    }
  }
  qsort(out, nOut, sizeof(ghost_msg), cmp_ghost_msg);
  const int nSendRanks = sorted_unique(outRanks, nOutRemote);

  C->num_sends    = nSendRanks;
  C->send_ranks   = (int *)malloc((size_t)(nSendRanks > 0 ? nSendRanks : 1) * sizeof(int));
  C->send_sizes   = (int *)calloc((size_t)(nSendRanks > 0 ? nSendRanks : 1), sizeof(int));
  C->send_buffers = (double **)calloc((size_t)(nSendRanks > 0 ? nSendRanks : 1), sizeof(double *));
  for (int r = 0; r < nSendRanks; r++) C->send_ranks[r] = outRanks[r];

  /* pass 1: sizes; then one bulk (zeroed) device allocation carved per neighbour */
  size_t total = 0;
  for (int m = 0; m < nOut; m++) {
    if (out[m].recvRank == L->my_rank) continue;
    int di, dj, dk, s, ei, ej, ek, r;
    dir_to_offsets(out[m].sendDir, &di, &dj, &dk);
    axis_geometry(di, n, g, &s, &ei, &r);
    axis_geometry(dj, n, g, &s, &ej, &r);
    axis_geometry(dk, n, g, &s, &ek, &r);
    C->send_sizes[find_rank(C->send_ranks, nSendRanks, out[m].recvRank)] += ei * ej * ek;
    total += (size_t)ei * ej * ek;
  }
  if (nSendRanks > 0) {
    double *bulk = (double *)MALLOC(total * sizeof(double));
    for (int r = 0; r < nSendRanks; r++) { C->send_buffers[r] = bulk; bulk += C->send_sizes[r]; }
  }
  /* pass 2: populate pack + local lists, tracking the running offset per neighbour */
  int *cursor = (int *)calloc((size_t)(nSendRanks > 0 ? nSendRanks : 1), sizeof(int));
  for (int m = 0; m < nOut; m++) {
    int di, dj, dk, si, sj, sk, ei, ej, ek, ri, rj, rk;
    dir_to_offsets(out[m].sendDir, &di, &dj, &dk);
    axis_geometry(di, n, g, &si, &ei, &ri);
    axis_geometry(dj, n, g, &sj, &ej, &rj);
    axis_geometry(dk, n, g, &sk, &ek, &rk);
    const box_type *sb = &L->my_boxes[out[m].sendBox];
    if (out[m].recvRank == L->my_rank) {
      const box_type *rb = &L->my_boxes[out[m].recvBox];
      append_block_to_list(&C->blocks[1], &C->allocated_blocks[1], &C->num_blocks[1], ei, ej, ek,
                           out[m].sendBox, NULL, si, sj, sk, sb->jStride, sb->kStride, 1,
                           out[m].recvBox, NULL, ri, rj, rk, rb->jStride, rb->kStride, 1,
                           BLOCKCOPY_TILE_I, 16, 16, 0);
    } else {
      int nb = find_rank(C->send_ranks, nSendRanks, out[m].recvRank);
      append_block_to_list(&C->blocks[0], &C->allocated_blocks[0], &C->num_blocks[0], ei, ej, ek,
                           out[m].sendBox, NULL, si, sj, sk, sb->jStride, sb->kStride, 1,
                           -1, C->send_buffers[nb], cursor[nb], 0, 0, ei, ei * ej, 1,
                           BLOCKCOPY_TILE_I, 16, 16, 0);
      cursor[nb] += ei * ej * ek;
    }
  }
  free(cursor);
  free(out);
  free(outRanks);

  /* ---- what I receive --------------------------------------------------------------- */
  ghost_msg *in = (ghost_msg *)malloc((size_t)maxMsgs * sizeof(ghost_msg));
  int *inRanks  = (int *)malloc((size_t)maxMsgs * sizeof(int));
  int nIn = 0;
  for (int box = 0; box < L->num_my_boxes; box++) {
    int bi, bj, bk;
    box_grid_coords(L, box, &bi, &bj, &bk);
    for (int d = 0; d < 27; d++) {
      if (!shape_uses_dir(shape, d)) continue;
      int di, dj, dk;
      dir_to_offsets(d, &di, &dj, &dk);
      int nb = adjacent_box_id(L, bi, bj, bk, di, dj, dk);
      if (nb < 0) continue;
      int owner = L->rank_of_box[nb];
      if (owner == -1 || owner == L->my_rank) continue;
      ghost_msg *m = &in[nIn];
      m->sendRank = owner;        m->sendBoxID = nb;   m->sendBox = -1;
      m->sendDir  = 26 - d;       /* the direction as the sender sees it */
      m->recvRank = L->my_rank;   m->recvBoxID = L->my_boxes[box].global_box_id;  m->recvBox = box;
      inRanks[nIn++] = owner;
    }
  }
  qsort(in, nIn, sizeof(ghost_msg), cmp_ghost_msg);
  const int nRecvRanks = sorted_unique(inRanks, nIn);

  C->num_recvs    = nRecvRanks;
  C->recv_ranks   = (int *)malloc((size_t)(nRecvRanks > 0 ? nRecvRanks : 1) * sizeof(int));
  C->recv_sizes   = (int *)calloc((size_t)(nRecvRanks > 0 ? nRecvRanks : 1), sizeof(int));
  C->recv_buffers = (double **)calloc((size_t)(nRecvRanks > 0 ? nRecvRanks : 1), sizeof(double *));
  for (int r = 0; r < nRecvRanks; r++) C->recv_ranks[r] = inRanks[r];

  total = 0;
  for (int m = 0; m < nIn; m++) {
    int di, dj, dk, s, ei, ej, ek, r;
    dir_to_offsets(in[m].sendDir, &di, &dj, &dk);
    axis_geometry(di, n, g, &s, &ei, &r);
    axis_geometry(dj, n, g, &s, &ej, &r);
    axis_geometry(dk, n, g, &s, &ek, &r);
    C->recv_sizes[find_rank(C->recv_ranks, nRecvRanks, in[m].sendRank)] += ei * ej * ek;
    total += (size_t)ei * ej * ek;
  }
  if (nRecvRanks > 0) {
    double *bulk = (double *)hpgmg_rt_alloc_comm(total * sizeof(double));   /* peer-visible: neighbours store straight into it */
    for (int r = 0; r < nRecvRanks; r++) { C->recv_buffers[r] = bulk; bulk += C->recv_sizes[r]; }
  }
  cursor = (int *)calloc((size_t)(nRecvRanks > 0 ? nRecvRanks : 1), sizeof(int));
  for (int m = 0; m < nIn; m++) {
    int di, dj, dk, s, ei, ej, ek, ri, rj, rk;
    dir_to_offsets(in[m].sendDir, &di, &dj, &dk);
    axis_geometry(di, n, g, &s, &ei, &ri);
    axis_geometry(dj, n, g, &s, &ej, &rj);
    axis_geometry(dk, n, g, &s, &ek, &rk);
    int nb = find_rank(C->recv_ranks, nRecvRanks, in[m].sendRank);
    const box_type *rb = &L->my_boxes[in[m].recvBox];
    append_block_to_list(&C->blocks[2], &C->allocated_blocks[2], &C->num_blocks[2], ei, ej, ek,
                         -1, C->recv_buffers[nb], cursor[nb], 0, 0, ei, ei * ej, 1,
                         in[m].recvBox, NULL, ri, rj, rk, rb->jStride, rb->kStride, 1,
                         BLOCKCOPY_TILE_I, 16, 16, 0);
    cursor[nb] += ei * ej * ek;
  }
  free(cursor);
  free(in);
  free(inRanks);
}

/* ------------------------------------------------------------------------------------------ */
/* (Re)allocate the level's vectors.  Growing keeps the contents of the existing vectors
 * (level.c:929-1068).  Layout arithmetic is the reference's: jStride = dim+2g rounded up to
 * BOX_ALIGN_JSTRIDE, kStride = jStride*(dim+2g), volume = kStride*(dim+2g). */
void create_vectors(level_type *L, int numVectors)
{
  if (numVectors <= L->numVectors) return;
  const int old_n = L->numVectors;
  const int padded = L->box_dim + 2 * L->box_ghosts;

  L->box_jStride = padded;
  while (L->box_jStride % BOX_ALIGN_JSTRIDE) L->box_jStride++;
  L->box_kStride = L->box_jStride * padded;
  L->box_volume  = L->box_kStride * padded;

  /* one zeroed slab; the first interior cell of box 0 / vector 0 is 128-byte aligned */
  const size_t origin = (size_t)L->box_ghosts * (1 + L->box_jStride + L->box_kStride);
  const size_t lead   = (16 - (origin % 16)) % 16;                 /* doubles: 16*8 = 128 B */
  const size_t per_box = (size_t)numVectors * L->box_volume;
  double *old_slab = (L->num_my_boxes > 0 && old_n > 0) ? L->my_boxes[0].fp_base : NULL;
  double *slab = NULL;
  if (L->num_my_boxes > 0) {
    slab = (double *)MALLOC((per_box * L->num_my_boxes + lead + 16) * sizeof(double));
    if (!slab) { fprintf(stderr, "create_vectors: device allocation failed\n"); exit(0); }
  }
  for (int b = 0; b < L->num_my_boxes; b++) {
    box_type *box = &L->my_boxes[b];
    double **old_vectors = (old_n > 0) ? box->vectors : NULL;
    double *base = slab + lead + (size_t)b * per_box;
    if (old_n > 0) hpgmg_rt_copy_d2d(base, old_vectors[0], (size_t)old_n * L->box_volume * sizeof(double));
    box->vectors = (double **)malloc((size_t)numVectors * sizeof(double *));
    for (int v = 0; v < numVectors; v++) box->vectors[v] = base + (size_t)v * L->box_volume;
    free(old_vectors);
    box->fp_base = (b == 0) ? slab : NULL;   /* only box 0 owns the allocation */
  }
  if (old_slab) { hpgmg_rt_sync(); FREE(old_slab); }

  /* (re)describe my boxes in global-id order, i fastest (level.c:1043-1064) */
  int b = 0;
  for (int k = 0; k < L->boxes_in.k; k++)
  for (int j = 0; j < L->boxes_in.j; j++)
  for (int i = 0; i < L->boxes_in.i; i++) {
    int id = i + j * L->boxes_in.i + k * L->boxes_in.i * L->boxes_in.j;
    if (L->rank_of_box[id] != L->my_rank) continue;
    box_type *box = &L->my_boxes[b++];
    box->numVectors = numVectors;
    box->dim = L->box_dim;          box->ghosts = L->box_ghosts;
    box->jStride = L->box_jStride;  box->kStride = L->box_kStride;  box->volume = L->box_volume;
    box->low.i = i * L->box_dim;    box->low.j = j * L->box_dim;    box->low.k = k * L->box_dim;
    box->global_box_id = id;
  }
  L->numVectors = numVectors;
  if (HPGMG_DEV(L)) hpgmg_device_level_rebind_vectors(L);
}

/* ------------------------------------------------------------------------------------------ */
void create_level(level_type *L, int boxes_in_i, int box_dim, int box_ghosts, int numVectors,
                  int domain_boundary_condition, int my_rank, int num_ranks)
{
  const int totalBoxes = boxes_in_i * boxes_in_i * boxes_in_i;
  const int chatty = (my_rank == 0) && hpgmg_rt_verbose();
  if (chatty) {
    fprintf(stdout, "\nattempting to create a %d^3 level from %d x %d^3 boxes distributed among %d tasks...\n",
            box_dim * boxes_in_i, totalBoxes, box_dim, num_ranks);
    fprintf(stdout, domain_boundary_condition == BC_DIRICHLET ? "  boundary condition = BC_DIRICHLET\n"
                                                              : "  boundary condition = BC_PERIODIC\n");
  }
  if (box_ghosts < stencil_get_radius()) {
    if (my_rank == 0) fprintf(stderr, "ghosts(%d) must be >= stencil_get_radius(%d)\n", box_ghosts, stencil_get_radius());
    exit(0);
  }

  hpgmg_comm_recycle_arena();      /* collective: peer-visible buffers of destroyed hierarchies are handed out again */
  memset(L, 0, sizeof(*L));
  L->box_dim = box_dim;           L->box_ghosts = box_ghosts;
  L->boxes_in.i = L->boxes_in.j = L->boxes_in.k = boxes_in_i;
  L->dim.i = L->dim.j = L->dim.k = box_dim * boxes_in_i;
  L->active = 1;                  L->my_rank = my_rank;         L->num_ranks = num_ranks;
  L->boundary_condition.type = domain_boundary_condition;
  L->must_subtract_mean = -1;
  L->num_threads = 1;
  L->tag = (int)log2((double)L->dim.i);

  L->rank_of_box = (int *)malloc((size_t)totalBoxes * sizeof(int));
  if (!L->rank_of_box) { fprintf(stderr, "malloc of level->rank_of_box failed\n"); exit(0); }
  for (int b = 0; b < totalBoxes; b++) L->rank_of_box[b] = -1;

  if (chatty) { fprintf(stdout, "  Decomposing level via Z-mort ordering... "); fflush(stdout); }
  zmort_assign(L->rank_of_box, boxes_in_i, boxes_in_i, boxes_in_i, 0, 0, 0,
               boxes_in_i, boxes_in_i, boxes_in_i, num_ranks, 0, totalBoxes);
  if (chatty) { fprintf(stdout, "done\n"); fflush(stdout); }

  for (int b = 0; b < totalBoxes; b++) if (L->rank_of_box[b] == my_rank) L->num_my_boxes++;
  L->my_boxes = (box_type *)calloc((size_t)(L->num_my_boxes > 0 ? L->num_my_boxes : 1), sizeof(box_type));

  if (chatty) { fprintf(stdout, "  Allocating vectors... "); fflush(stdout); }
  create_vectors(L, numVectors);
  if (chatty) { fprintf(stdout, "done\n"); fflush(stdout); }

  /* flatten boxes into 10000x8x8 compute tiles (level.c:1184-1210) */
  for (int b = 0; b < L->num_my_boxes; b++) {
    const box_type *box = &L->my_boxes[b];
    append_block_to_list(&L->my_blocks, &L->allocated_blocks, &L->num_my_blocks,
                         box->dim, box->dim, box->dim,
                         b, NULL, 0, 0, 0, box->jStride, box->kStride, 1,
                         b, NULL, 0, 0, 0, box->jStride, box->kStride, 1,
                         BLOCKCOPY_TILE_I, BLOCKCOPY_TILE_J, BLOCKCOPY_TILE_K, 0);
  }

  for (int shape = 0; shape < STENCIL_MAX_SHAPES; shape++) build_exchange_ghosts(L, shape);
  for (int shape = 0; shape < STENCIL_MAX_SHAPES; shape++) hpgmg_comm_register_exchange(L, shape);   /* collective: swaps buffer/flag addresses */
  for (int shape = 0; shape < STENCIL_MAX_SHAPES; shape++) build_boundary_conditions(L, shape);
  for (int t = 0; t < 4; t++) reset_communicator(&L->restriction[t]);
  reset_communicator(&L->interpolation);

  hpgmg_device_level_create(L);

  if (chatty) {
    /* every rank holds the full rank_of_box map, so the reference's MPI_Allreduce(MAX)
     * (level.c:1255) is a local count here */
    int *per_rank = (int *)calloc((size_t)num_ranks, sizeof(int)), most = 0;
    for (int b = 0; b < totalBoxes; b++) if (L->rank_of_box[b] >= 0) per_rank[L->rank_of_box[b]]++;
    for (int r = 0; r < num_ranks; r++) if (per_rank[r] > most) most = per_rank[r];
    free(per_rank);
    fprintf(stdout, "  Calculating boxes per process... target=%0.3f, max=%d\n",
            (double)totalBoxes / (double)num_ranks, most);
  }
}

/* ------------------------------------------------------------------------------------------ */
void reset_level_timers(level_type *L)
{
  memset(&L->timers, 0, sizeof(L->timers));
  L->Krylov_iterations = 0;
  L->CAKrylov_formations_of_G = 0;
  L->vcycles_from_this_level = 0;
}

static void free_communicator(communicator_type *c)
{
  if (c->num_recvs > 0 && c->recv_buffers && c->recv_buffers[0]) FREE(c->recv_buffers[0]);
  if (c->num_sends > 0 && c->send_buffers && c->send_buffers[0]) FREE(c->send_buffers[0]);
  free(c->recv_buffers); free(c->recv_ranks); free(c->recv_sizes);
  free(c->send_buffers); free(c->send_ranks); free(c->send_sizes);
  for (int l = 0; l < 3; l++) free(c->blocks[l]);
  memset(c, 0, sizeof(*c));
}

/* exported for mg.c (MGDestroy frees the inter-level programs) */
void hpgmg_free_communicator(communicator_type *c) { free_communicator(c); }

void destroy_level(level_type *L)
{
  const int chatty = (L->my_rank == 0) && hpgmg_rt_verbose();
  if (chatty) { fprintf(stdout, "attempting to free the %5d^3 level... ", L->dim.i); fflush(stdout); }
  hpgmg_rt_sync();
  hpgmg_device_level_destroy(L);
  for (int b = 0; b < L->num_my_boxes; b++) free(L->my_boxes[b].vectors);
  if (L->num_my_boxes > 0 && L->my_boxes[0].fp_base) FREE(L->my_boxes[0].fp_base);
  free(L->rank_of_box);
  free(L->my_boxes);
  free(L->my_blocks);
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) free(L->boundary_condition.blocks[s]);
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++) { hpgmg_comm_unregister(&L->exchange_ghosts[s]); free_communicator(&L->exchange_ghosts[s]); }
  if (chatty) fprintf(stdout, "done\n");
}
