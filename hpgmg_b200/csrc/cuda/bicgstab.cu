/*
 * bicgstab.cu -- the bottom solver as ONE single-thread-block kernel.
 *
 * Restates solvers/bicgstab.c:14-97 (diagonally right-preconditioned BiCGStab, <=200 iterations,
 * stop when ||s||inf or ||r||inf < rtol*||r0||inf, exact-zero breakdown tests).  In the reference
 * every scalar (6 per iteration) steers a host branch; here the coarsest grid is one box of at
 * most 11^3 cells, so a single CTA runs the whole solve with the scalars in registers: no host
 * round trip, no launch per BLAS1 op, and the kernel can sit inside a captured CUDA graph.
 *
 * Bit-exactness: dot products are accumulated in the reference's order (k, j, i; one tile) by one
 * thread over products staged in shared memory; the max-norm is order-free; apply_op / residual use
 * the same device bodies (stencil.cuh, bc.cuh) as the level-wide kernels.
 */
#include <math.h>
#include "common.cuh"
#include "stencil.cuh"
#include "bc.cuh"

#define BOTTOM_MAX_DIM   11
#define BOTTOM_MAX_CELLS (BOTTOM_MAX_DIM * BOTTOM_MAX_DIM * BOTTOM_MAX_DIM)
#define BOTTOM_THREADS   256

struct BottomArgs {
  DLevel L;
  const blockCopy_type *bc;     /* NO_CORNERS list of the (single) box */
  int nbc;
  int x_id, R_id;
  double a, b, h2inv, rtol;
  double *iters;                /* device scalar slot: iterations are added to it */
};

struct BottomCtx {
  const BottomArgs &A;
  double *prod;                 /* shared: BOTTOM_MAX_CELLS products / scratch */
  double *red;                  /* shared: per-warp partials + broadcast slot  */
  int n, cells, jS, kS;
  __device__ int cell_offset(int c) const { return (c % n) + ((c / n) % n) * jS + (c / (n * n)) * kS; }
};

__device__ static void b_fill_ghosts(const BottomCtx &C, const int id)
{
  /* exchange_boundary is empty for a single box with Dirichlet BCs; apply_BCs = v4 (v2 if dim<4) */
  for (int e = 0; e < C.A.nbc; e++) {
    const blockCopy_type B = C.A.bc[e];
    if (C.n < 4) bc_v2_block(C.A.L, id, B, threadIdx.x, blockDim.x);
    else         bc_v4_block(C.A.L, id, B, threadIdx.x, blockDim.x);
  }
  __syncthreads();
}

/* out = A in   (mode 0)   or   out = rhs - A in   (mode 1) */
__device__ static void b_apply(const BottomCtx &C, const int out_id, const int in_id, const int rhs_id, const int mode)
{
  __syncthreads();
  b_fill_ghosts(C, in_id);
  const DLevel &L = C.A.L;
  const double *x = L.vec(0, in_id), *bi = L.vec(0, VECTOR_BETA_I), *bj = L.vec(0, VECTOR_BETA_J), *bk = L.vec(0, VECTOR_BETA_K);
  double *out = L.vec(0, out_id);
  const double *rhs = L.vec(0, rhs_id);
  for (int c = threadIdx.x; c < C.cells; c += blockDim.x) {
    const int ijk = C.cell_offset(c);
    const double Ax = fv4_apply_op(x + ijk, bi + ijk, bj + ijk, bk + ijk, C.jS, C.kS, C.A.b, C.A.h2inv);
    out[ijk] = mode ? rhs[ijk] - Ax : Ax;
  }
  __syncthreads();
}

/* c = sa*a + sb*b */
__device__ static void b_add(const BottomCtx &C, const int c_id, const double sa, const int a_id, const double sb, const int b_id)
{
  const DLevel &L = C.A.L;
  double *c = L.vec(0, c_id);
  const double *a = L.vec(0, a_id), *b = L.vec(0, b_id);
  for (int q = threadIdx.x; q < C.cells; q += blockDim.x) {
    const int ijk = C.cell_offset(q);
    c[ijk] = sa * a[ijk] + sb * b[ijk];
  }
  __syncthreads();
}
__device__ static void b_scale(const BottomCtx &C, const int c_id, const double sa, const int a_id)
{
  const DLevel &L = C.A.L;
  double *c = L.vec(0, c_id);
  const double *a = L.vec(0, a_id);
  for (int q = threadIdx.x; q < C.cells; q += blockDim.x) {
    const int ijk = C.cell_offset(q);
    c[ijk] = sa * a[ijk];
  }
  __syncthreads();
}
__device__ static void b_mul(const BottomCtx &C, const int c_id, const double s, const int a_id, const int b_id)
{
  const DLevel &L = C.A.L;
  double *c = L.vec(0, c_id);
  const double *a = L.vec(0, a_id), *b = L.vec(0, b_id);
  for (int q = threadIdx.x; q < C.cells; q += blockDim.x) {
    const int ijk = C.cell_offset(q);
    c[ijk] = s * a[ijk] * b[ijk];
  }
  __syncthreads();
}

__device__ static double b_dot(const BottomCtx &C, const int a_id, const int b_id)
{
  const DLevel &L = C.A.L;
  const double *a = L.vec(0, a_id), *b = L.vec(0, b_id);
  for (int q = threadIdx.x; q < C.cells; q += blockDim.x) {
    const int ijk = C.cell_offset(q);
    C.prod[q] = a[ijk] * b[ijk];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int q = 0; q < C.cells; q++) s += C.prod[q];      /* k,j,i order == linear cell order */
    C.red[32] = s;
  }
  __syncthreads();
  const double r = C.red[32];
  __syncthreads();
  return r;
}

__device__ static double b_norm(const BottomCtx &C, const int a_id)
{
  const DLevel &L = C.A.L;
  const double *a = L.vec(0, a_id);
  double m = 0.0;
  for (int q = threadIdx.x; q < C.cells; q += blockDim.x) {
    const double f = fabs(a[C.cell_offset(q)]);
    if (f > m) m = f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_down_sync(0xffffffffu, m, o);
    if (other > m) m = other;
  }
  if ((threadIdx.x & 31) == 0) C.red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) if (C.red[w] > m) m = C.red[w];
    C.red[32] = m;
  }
  __syncthreads();
  const double r = C.red[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(BOTTOM_THREADS) bicgstab_kernel(const BottomArgs A)
{
  __shared__ double prod[BOTTOM_MAX_CELLS];
  __shared__ double red[33];
  BottomCtx C = { A, prod, red, A.L.dim, A.L.dim * A.L.dim * A.L.dim, A.L.jStride, A.L.kStride };

  const int r0 = VECTORS_RESERVED + 0, r = VECTORS_RESERVED + 1, p = VECTORS_RESERVED + 2, q = VECTORS_RESERVED + 3;
  const int s = VECTORS_RESERVED + 4, t = VECTORS_RESERVED + 5, Ap = VECTORS_RESERVED + 6, As = VECTORS_RESERVED + 7;
  const int x_id = A.x_id;
  const int jMax = 200;
  int j = 0;
  bool failed = false, converged = false;

  b_apply(C, r0, x_id, A.R_id, 1);                        /* r0 = R - A x */
  b_scale(C, r, 1.0, r0);
  b_scale(C, p, 1.0, r0);
  double r_dot_r0 = b_dot(C, r, r0);
  const double norm_of_r0 = b_norm(C, r);
  if (r_dot_r0 == 0.0) converged = true;
  if (norm_of_r0 == 0.0) converged = true;
  while ((j < jMax) && !failed && !converged) {
    j++;
    b_mul(C, q, 1.0, VECTOR_DINV, p);                     /* q = D^-1 p */
    b_apply(C, Ap, q, 0, 0);                              /* Ap = A q   */
    const double Ap_dot_r0 = b_dot(C, Ap, r0);
    if (Ap_dot_r0 == 0.0) { failed = true; break; }
    const double alpha = r_dot_r0 / Ap_dot_r0;
    if (isinf(alpha)) { failed = true; break; }
    b_add(C, x_id, 1.0, x_id, alpha, q);
    b_add(C, s, 1.0, r, -alpha, Ap);
    const double norm_of_s = b_norm(C, s);
    if (norm_of_s == 0.0) { converged = true; break; }
    if (norm_of_s < A.rtol * norm_of_r0) { converged = true; break; }
    b_mul(C, t, 1.0, VECTOR_DINV, s);                     /* t = D^-1 s */
    b_apply(C, As, t, 0, 0);                              /* As = A t   */
    const double As_dot_As = b_dot(C, As, As);
    const double As_dot_s = b_dot(C, As, s);
    if (As_dot_As == 0.0) { converged = true; break; }
    const double omega = As_dot_s / As_dot_As;
    if (omega == 0.0) { failed = true; break; }
    if (isinf(omega)) { failed = true; break; }
    b_add(C, x_id, 1.0, x_id, omega, t);
    b_add(C, r, 1.0, s, -omega, As);
    const double norm_of_r = b_norm(C, r);
    if (norm_of_r == 0.0) { converged = true; break; }
    if (norm_of_r < A.rtol * norm_of_r0) { converged = true; break; }
    const double r_dot_r0_new = b_dot(C, r, r0);
    if (r_dot_r0_new == 0.0) { failed = true; break; }
    const double beta = (r_dot_r0_new / r_dot_r0) * (alpha / omega);
    if (isinf(beta)) { failed = true; break; }
    b_add(C, VECTOR_TEMP, 1.0, p, -omega, Ap);
    b_add(C, p, 1.0, r, beta, VECTOR_TEMP);
    r_dot_r0 = r_dot_r0_new;
  }
  if (threadIdx.x == 0) atomicAdd(A.iters, (double)j);
}

/* Returns 1 if the solve was enqueued on the device, 0 if the level is not eligible (more than one
 * box on this rank, other ranks involved, a periodic problem that needs mean subtraction, or a box
 * larger than the kernel's shared-memory staging) and the caller must use the host-driven loop. */
extern "C" int hpgmg_bicgstab_device(level_type *level, int x_id, int R_id, double a, double b, double rtol)
{
  if (level->must_subtract_mean == 1) return 0;
  if (level->boundary_condition.type != BC_DIRICHLET) return 0;
  if (level->num_my_boxes != 1) return 0;
  if (level->boxes_in.i != 1 || level->boxes_in.j != 1 || level->boxes_in.k != 1) return 0;
  if (level->box_dim > BOTTOM_MAX_DIM || level->box_dim < 2) return 0;
  hpgmg_device_level *D = HPGMG_DEV(level);
  BottomArgs A;
  A.L = D->L;
  A.bc = D->bc[STENCIL_SHAPE_NO_CORNERS].blocks;
  A.nbc = D->bc[STENCIL_SHAPE_NO_CORNERS].n;
  A.x_id = x_id;  A.R_id = R_id;  A.a = a;  A.b = b;
  A.h2inv = 1.0 / (level->h * level->h);
  A.rtol = rtol;
  A.iters = hpgmg_rt_scalar_slots() + HPGMG_SLOT_KRYLOV;
  LAUNCH(bicgstab_kernel, 1, BOTTOM_THREADS, 0, A);
  return 1;
}

