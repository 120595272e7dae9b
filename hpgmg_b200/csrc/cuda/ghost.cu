/*
 * ghost.cu -- everything that fills ghost cells: the intra-GPU part of exchange_boundary, the
 * homogeneous-Dirichlet boundary conditions (linear / quadratic / quartic extrapolation of cell
 * averages) and the extrapolation of the face coefficients.
 *
 *   exchange_boundary   operators/exchange_boundary.c:12-117 + CopyBlock operators/blockCopy.c:6-105
 *   apply_BCs_v1/v2/v4  operators/boundary_fv.c:6-90 / 101-250 / 262-569 (apply_BCs = v4, operators.fv4.c:51)
 *   extrapolate_betas   operators/boundary_fv.c:573-681
 *
 * The kernels walk the device copies of the SAME block lists the reference walks (level.c:367-465,
 * 498-922): one thread block per list entry.  A BC entry's `subtype` is the normal to the DOMAIN
 * (0..26 = 13+di+3dj+9dk); ghost cells along each normal axis are produced by a 1-D extrapolation
 * from the 4 (v4) interior cells along that axis, applied axis by axis in i, j, k order -- the
 * order the reference's face / edge (16-point) / corner (64-point) code uses -- so results are
 * bit-identical.
 */
#include "common.cuh"

/* ---- copy lists ------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128) copy_blocks_kernel(const DLevel L, const int id, const blockCopy_type *__restrict__ blocks)
{
  const blockCopy_type B = blocks[blockIdx.x];
  const double *__restrict__ rd;
  double *__restrict__ wr;
  int rj, rk, wj, wk;
  if (B.read.box >= 0) { rd = L.vec(B.read.box, id); rj = L.jStride; rk = L.kStride; }
  else                 { rd = B.read.ptr;            rj = B.read.jStride; rk = B.read.kStride; }
  if (B.write.box >= 0) { wr = L.vec(B.write.box, id); wj = L.jStride; wk = L.kStride; }
  else                  { wr = B.write.ptr;            wj = B.write.jStride; wk = B.write.kStride; }
  rd += B.read.i + B.read.j * rj + B.read.k * rk;
  wr += B.write.i + B.write.j * wj + B.write.k * wk;
  const int di = B.dim.i, dj = B.dim.j, cells = di * dj * B.dim.k;
  for (int c = threadIdx.x; c < cells; c += blockDim.x) {
    const int i = c % di, j = (c / di) % dj, k = c / (di * dj);
    wr[i + j * wj + k * wk] = rd[i + j * rj + k * rk];
  }
}

void hpgmg_run_copy_list(const DLevel &L, int id, const DList &list)
{
  if (list.n <= 0) return;
  LAUNCH(copy_blocks_kernel, list.n, 128, 0, L, id, list.blocks);
}

extern "C" void exchange_boundary(level_type *level, int id, int shape)
{
  if (shape >= STENCIL_MAX_SHAPES) shape = STENCIL_SHAPE_BOX;
  hpgmg_device_level *D = HPGMG_DEV(level);
  if (level->num_ranks > 1 && (level->exchange_ghosts[shape].num_sends > 0 || level->exchange_ghosts[shape].num_recvs > 0)) {
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][0]);                 /* pack   */
    hpgmg_comm_exchange(level, &level->exchange_ghosts[shape], 0);         /* send / recv */
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][1]);                 /* local  */
    hpgmg_comm_exchange_wait(level, &level->exchange_ghosts[shape]);
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][2]);                 /* unpack */
  } else {
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][1]);
  }
}

/* ---- boundary conditions (device bodies in bc.cuh) ---------------------------------------------- */
#include "bc.cuh"

__global__ void __launch_bounds__(128) bc_v4_kernel(const DLevel L, const int id, const blockCopy_type *__restrict__ blocks)
{
  const blockCopy_type B = blocks[blockIdx.x];
  bc_v4_block(L, id, B, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(128) bc_v2_kernel(const DLevel L, const int id, const blockCopy_type *__restrict__ blocks)
{
  const blockCopy_type B = blocks[blockIdx.x];
  bc_v2_block(L, id, B, threadIdx.x, blockDim.x);
}

/* linear: every ghost cell of a block mirrors (with sign) the cell one step along the inverted
 * normal (boundary_fv.c:6-90).  With two ghost layers a deeper ghost reads a shallower one, possibly
 * of ANOTHER block, so the result depends on the order: the whole list is walked by one thread in
 * list order with k,j,i ascending inside a block -- the reference run on one thread.  The path only
 * exists for box_dim<2, which the radius-2 fv4 stencil never produces; kept for API completeness. */
__global__ void bc_v1_kernel(const DLevel L, const int id, const blockCopy_type *__restrict__ blocks, const int nblocks)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int e = 0; e < nblocks; e++) {
    const blockCopy_type B = blocks[e];
    const BCGeom G = bc_geometry(B, B.subtype, L.dim, L.jStride, L.kStride);
    double *x = L.vec(B.read.box, id);
    int m = 0;
    for (int a = 0; a < 3; a++) m += G.normal[a] ? 1 : 0;
    const double scale = (m == 2) ? 1.0 : -1.0;           /* faces -1, edges +1, corners -1 */
    const int stride = -G.normal[0] - G.normal[1] * L.jStride - G.normal[2] * L.kStride;
    const int cells = G.ext[0] * G.ext[1] * G.ext[2];
    for (int c = 0; c < cells; c++) {
      const int i = c % G.ext[0], j = (c / G.ext[0]) % G.ext[1], k = c / (G.ext[0] * G.ext[1]);
      const int ijk = (i + G.lo[0]) + (j + G.lo[1]) * L.jStride + (k + G.lo[2]) * L.kStride;
      x[ijk] = scale * x[ijk + stride];
    }
  }
}

/* ---- fused ghost fill: exchange_boundary + apply_BCs in ONE kernel ------------------------------- */
/* Threads [0, ncopies) copy one ghost cell each from the neighbouring box on this GPU; threads
 * [ncopies, ncopies+nbc) extrapolate one BC column each, reading interior values straight from the
 * box that owns them (FillBC::src), which makes the two kinds of work independent.  version: 4
 * quartic, 2 quadratic (also zeroes the deeper ghost layer, like the reference). */
__global__ void __launch_bounds__(256) fill_ghosts_kernel(const DLevel L, const int id, const FillCopy *__restrict__ copies, const int ncopies,
                                                          const FillBC *__restrict__ bc, const int nbc, const int version)
{
  double *v = L.base + (size_t)id * (size_t)L.volume;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ncopies) {
    const FillCopy c = copies[t];
    v[c.dst] = v[c.src];
  } else if (t < ncopies + nbc) {
    const FillBC it = bc[t - ncopies];
    const BCNormal N = bc_normal(it.subtype, L.jStride, L.kStride);
    if (version == 4) bc_v4_column(v + it.src, v + it.dst, N);
    else              bc_v2_col_zero_rest(v + it.src, v + it.dst, N.m, N.d[0], N.d[1], N.d[2]);
  }
}

/* exchange_boundary(level,id,shape) followed by apply_BCs_v4 (bc_version 4; v2 below 4^3 like
 * boundary_fv.c:269) or apply_BCs_v2 (bc_version 2).  Identical results to the two separate calls. */
void hpgmg_fill_ghosts(level_type *level, int id, int shape, int bc_version)
{
  hpgmg_device_level *D = HPGMG_DEV(level);
  const FillTable &T = D->fill[shape];
  const bool dirichlet = level->boundary_condition.type == BC_DIRICHLET;
  int version = bc_version;
  if (version == 4 && level->box_dim < 4) version = 2;
  const bool have_tables = D->fill_nvec == level->numVectors && level->box_ghosts == 2 && !(version == 2 && level->box_dim < 2);
  if (!have_tables) {                                   /* unusual geometry: the list kernels */
    exchange_boundary(level, id, shape);
    if (bc_version == 4) apply_BCs_v4(level, id, shape); else apply_BCs_v2(level, id, shape);
    return;
  }
  communicator_type *C = &level->exchange_ghosts[shape];
  const bool remote = level->num_ranks > 1 && (C->num_sends > 0 || C->num_recvs > 0);
  if (remote) {
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][0]);                 /* pack   */
    hpgmg_comm_exchange(level, C, 0);                                      /* send / recv */
  }
  const int nbc = dirichlet ? T.nbc : 0;
  const int work = T.ncopies + nbc;
  if (work > 0) LAUNCH(fill_ghosts_kernel, (work + 255) / 256, 256, 0, D->L, id, T.copies, T.ncopies, T.bc, nbc, version);
  if (remote) {
    hpgmg_comm_exchange_wait(level, C);
    hpgmg_run_copy_list(D->L, id, D->exchange[shape][2]);                 /* unpack */
    if (dirichlet && T.nlate > 0) LAUNCH(fill_ghosts_kernel, (T.nlate + 255) / 256, 256, 0, D->L, id, (const FillCopy *)NULL, 0, T.late, T.nlate, version);
  }
}

extern "C" void apply_BCs_v1(level_type *level, int x_id, int shape)
{
  if (shape >= STENCIL_MAX_SHAPES) shape = STENCIL_SHAPE_BOX;
  if (level->boundary_condition.type == BC_PERIODIC) return;
  const DList &list = HPGMG_DEV(level)->bc[shape];
  if (list.n > 0) LAUNCH(bc_v1_kernel, 1, 32, 0, dl_of(level), x_id, list.blocks, list.n);
}

extern "C" void apply_BCs_v2(level_type *level, int x_id, int shape)
{
  if (shape >= STENCIL_MAX_SHAPES) shape = STENCIL_SHAPE_BOX;
  if (level->boundary_condition.type == BC_PERIODIC) return;
  if (level->box_dim < 2) { apply_BCs_v1(level, x_id, shape); return; }
  const DList &list = HPGMG_DEV(level)->bc[shape];
  if (list.n > 0) LAUNCH(bc_v2_kernel, list.n, 128, 0, dl_of(level), x_id, list.blocks);
}

extern "C" void apply_BCs_v4(level_type *level, int x_id, int shape)
{
  if (shape >= STENCIL_MAX_SHAPES) shape = STENCIL_SHAPE_BOX;
  if (level->boundary_condition.type == BC_PERIODIC) return;
  if (level->box_ghosts < 2) { fprintf(stderr, "called quartic BC's with only 1 ghost zone!!!\n"); exit(0); }
  if (level->box_dim < 4) { apply_BCs_v2(level, x_id, shape); return; }
  const DList &list = HPGMG_DEV(level)->bc[shape];
  if (list.n > 0) LAUNCH(bc_v4_kernel, list.n, 128, 0, dl_of(level), x_id, list.blocks);
}

extern "C" void apply_BCs(level_type *level, int x_id, int shape) { apply_BCs_v4(level, x_id, shape); }

/* ---- face coefficients beyond the domain ------------------------------------------------------- */
/* boundary_fv.c:573-681.  Uses the BOX-shape BC list only to know which ghost regions lie outside
 * the domain; the extrapolation direction is the BOX-relative normal of the region, and beta_d is
 * not extrapolated along d.  Each list entry is processed by ONE thread in the reference's loop
 * order (k, j, i ascending): the second ghost layer reads the first, so the order is part of the
 * result.  Setup only (untimed in the reference). */
__global__ void extrapolate_betas_kernel(const DLevel L, const blockCopy_type *__restrict__ blocks, const int n)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const blockCopy_type B = blocks[e];
  const int ilo = B.read.i, jlo = B.read.j, klo = B.read.k;
  int subtype = 13;
  if (ilo < 0) subtype -= 1;
  if (jlo < 0) subtype -= 3;
  if (klo < 0) subtype -= 9;
  if (ilo >= L.dim) subtype += 1;
  if (jlo >= L.dim) subtype += 3;
  if (klo >= L.dim) subtype += 9;
  const int normal = 26 - subtype;
  const int di = (normal % 3) - 1, dj = ((normal % 9) / 3) - 1, dk = (normal / 9) - 1;
  const int jS = L.jStride, kS = L.kStride;
  double *beta_i = L.vec(B.read.box, VECTOR_BETA_I);
  double *beta_j = L.vec(B.read.box, VECTOR_BETA_J);
  double *beta_k = L.vec(B.read.box, VECTOR_BETA_K);
  const int bi = dj * jS + dk * kS;
  const int bj = di + dk * kS;
  const int bk = di + dj * jS;
  const bool do_i = (subtype != 14) && (subtype != 12);
  const bool do_j = (subtype != 16) && (subtype != 10);
  const bool do_k = (subtype != 22) && (subtype != 4);
  for (int k = 0; k < B.dim.k; k++)
  for (int j = 0; j < B.dim.j; j++)
  for (int i = 0; i < B.dim.i; i++) {
    const int ijk = (i + ilo) + (j + jlo) * jS + (k + klo) * kS;
    if (L.dim >= 5) {
      if (do_i) beta_i[ijk] = 5.0 * beta_i[ijk + bi] - 10.0 * beta_i[ijk + 2 * bi] + 10.0 * beta_i[ijk + 3 * bi] - 5.0 * beta_i[ijk + 4 * bi] + beta_i[ijk + 5 * bi];
      if (do_j) beta_j[ijk] = 5.0 * beta_j[ijk + bj] - 10.0 * beta_j[ijk + 2 * bj] + 10.0 * beta_j[ijk + 3 * bj] - 5.0 * beta_j[ijk + 4 * bj] + beta_j[ijk + 5 * bj];
      if (do_k) beta_k[ijk] = 5.0 * beta_k[ijk + bk] - 10.0 * beta_k[ijk + 2 * bk] + 10.0 * beta_k[ijk + 3 * bk] - 5.0 * beta_k[ijk + 4 * bk] + beta_k[ijk + 5 * bk];
    } else if (L.dim >= 4) {
      if (do_i) beta_i[ijk] = 4.0 * beta_i[ijk + bi] - 6.0 * beta_i[ijk + 2 * bi] + 4.0 * beta_i[ijk + 3 * bi] - beta_i[ijk + 4 * bi];
      if (do_j) beta_j[ijk] = 4.0 * beta_j[ijk + bj] - 6.0 * beta_j[ijk + 2 * bj] + 4.0 * beta_j[ijk + 3 * bj] - beta_j[ijk + 4 * bj];
      if (do_k) beta_k[ijk] = 4.0 * beta_k[ijk + bk] - 6.0 * beta_k[ijk + 2 * bk] + 4.0 * beta_k[ijk + 3 * bk] - beta_k[ijk + 4 * bk];
    } else if (L.dim >= 2) {
      if (do_i) beta_i[ijk] = 2.0 * beta_i[ijk + bi] - beta_i[ijk + 2 * bi];
      if (do_j) beta_j[ijk] = 2.0 * beta_j[ijk + bj] - beta_j[ijk + 2 * bj];
      if (do_k) beta_k[ijk] = 2.0 * beta_k[ijk + bk] - beta_k[ijk + 2 * bk];
    }
  }
}

extern "C" void extrapolate_betas(level_type *level)
{
  if (level->boundary_condition.type == BC_PERIODIC) return;
  const DList &list = HPGMG_DEV(level)->bc[STENCIL_SHAPE_BOX];
  if (list.n > 0) LAUNCH(extrapolate_betas_kernel, (list.n + 63) / 64, 64, 0, dl_of(level), list.blocks, list.n);
}
