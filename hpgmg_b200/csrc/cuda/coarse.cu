/*
 * coarse.cu -- the coarse end of the multigrid cycle as ONE single-thread-block kernel.
 *
 * Levels of at most COARSE_MAX_CELLS cells (16^3 and coarser in the benchmark) carry ~4 % of the
 * flops but ~75 % of the kernel launches of an F-cycle: every operator there
 * is a few microseconds of launch latency around nanoseconds of work.  This kernel runs, for a
 * chain of such levels c..bottom that live entirely on this GPU,
 *
 *   MODE_VCYCLE : MGVCycle(c)                                             (mg.c:1135-1164)
 *   MODE_FTAIL  : the coarse tail of FMGSolve (mg.c:1285-1301): zero(e_bottom); bottom solve;
 *                 for l = bottom-1 .. c: interpolation_fcycle(l <- l+1); MGVCycle(l)
 *
 * with __syncthreads() where the stream version has kernel boundaries.  Every step calls the SAME
 * device bodies as the level-wide kernels (stencil.cuh, bc.cuh, bicgstab.cuh) or re-states their
 * loops over the same block lists, so the bits are identical to the multi-launch path
 * (tests/test_gpu_parity.py::test_coarse_kernel_equals_multilaunch_path).
 */
#include <math.h>
#include "common.cuh"
#include "stencil.cuh"
#include "bc.cuh"
#include "bicgstab.cuh"

#define COARSE_MAX_LEVELS 8
#define COARSE_THREADS    512

enum { MODE_VCYCLE = 0, MODE_FTAIL = 1 };

struct CoarseLevel {
  DLevel L;
  const int *low;
  const FillCopy *xch[2];                         /* [0] NO_CORNERS, [1] BOX: one record per ghost cell copied   */
  const FillBC *bc[2];                            /*                          one record per BC column            */
  int n_xch[2], n_bc[2];
  const blockCopy_type *restr, *interp;           /* local transfer lists */
  int n_restr, n_interp;
  double h2inv;
  double c1[6], c2[6];                            /* Chebyshev coefficients of this level */
  int smem_offset;                                /* >=0: the level's whole slab lives in shared memory during the kernel (doubles) */
  int slab_doubles;
};

/* phase clocks of the last profiled launch (hpgmg_b200_coarse_profile): SM cycles per category */
enum { CP_LOAD = 0, CP_FILL, CP_STENCIL, CP_RESTRICT, CP_ZERO, CP_INTERP, CP_BOTTOM, CP_STORE, CP_TOTAL, CP_N };
__device__ long long g_coarse_prof[CP_N];
__shared__ long long s_prof[CP_N + 1];
#define CPROF(cat) do { if (s_prof_on && threadIdx.x == 0) { const long long t_ = clock64(); s_prof[cat] += t_ - s_prof[CP_N]; s_prof[CP_N] = t_; } } while (0)
__shared__ int s_prof_on;

/* The cycle is handed to the kernel as a PROGRAM of phases (built on the host by the same recursion the stream
 * version runs, p_vcycle / p_ftail below) and executed by one loop with one switch, so that every operator body
 * exists once in the binary: written as nested inlined calls the kernel was 29 k SASS instructions of mostly
 * straight-line code and its 16 warps spent a third of their issue slots waiting for instruction fetches. */
enum { PH_FILL = 0, PH_STENCIL, PH_RESTRICT, PH_ZERO, PH_INTERP3, PH_INTERP5, PH_BOTTOM };
struct Phase { unsigned char op, lv, a, b, c, d, e, pad; };
#define COARSE_MAX_PHASES 384

struct CoarseArgs {
  int nlevels, mode, smoother, zero_bottom, profile, nphases;
  int e_id, R_id;
  double a, b, rtol;
  double *krylov;
  CoarseLevel lv[COARSE_MAX_LEVELS];
  Phase prog[COARSE_MAX_PHASES];
};

/* ---- cooperative (whole thread block) versions of the level operators ---------------------------- */
/* exchange_boundary + apply_BCs_v4 (or v2) for one shape, from the flat tables: copies and BC columns are
 * independent of each other (FillBC::src), so this is a single phase */
__device__ static void c_fill_ghosts(const CoarseLevel &V, const int id, const bool box_shape, const bool force_v2)
{
  const DLevel &L = V.L;
  double *v = L.base + (size_t)id * (size_t)L.volume;
  const int w = box_shape ? 1 : 0;
  const bool v2 = force_v2 || L.dim < 4;
  const int ncopies = V.n_xch[w], work = ncopies + V.n_bc[w];
  for (int e = threadIdx.x; e < work; e += blockDim.x) {
    if (e < ncopies) {
      const FillCopy c = V.xch[w][e];
      v[c.dst] = v[c.src];
    } else {
      const FillBC it = V.bc[w][e - ncopies];
      const BCNormal N = bc_normal(it.subtype, L.jStride, L.kStride);
      if (v2) bc_v2_col_zero_rest(v + it.src, v + it.dst, N.m, N.d[0], N.d[1], N.d[2]);
      else    bc_v4_column(v + it.src, v + it.dst, N);
    }
  }
  __syncthreads();
  CPROF(CP_FILL);
}

/* one sweep / residual over every cell of every box.  mode: 0 GSRB sweep s, 1 Chebyshev sweep s, 2 residual.
 * Even box sizes: a thread owns an i-pair (GSRB: exactly one active cell per pair, so no lane idles). */
__device__ static void c_stencil_pairs(const CoarseLevel &V, const int mode, const int src, const int dst, const int rhs_id, const int s, const double b)
{
  const DLevel &L = V.L;
  const int n = L.dim, jS = L.jStride, kS = L.kStride;
  const int hn = n >> 1, per_box = hn * n * n, total = per_box * L.nboxes;
  const bool pow2 = (n & (n - 1)) == 0;                            /* 8, 4, 2: shifts instead of five ~80-cycle integer divisions */
  const int lg = 31 - __clz(n);
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    int box, p, j, k;
    if (pow2) {
      box = q >> (3 * lg - 1);
      const int c = q & (per_box - 1);
      if (n == 8) {
        /* a warp = 4 pairs x 4 rows x 2 planes: with jStride 12 and kStride 144 the 16 lanes of a plane fall into
         * 16 distinct 8-byte banks and the two planes (opposite colour) into the same 16 again -- 2 wavefronts per
         * 64-bit load, the minimum; 4 pairs x 8 rows of one plane would be 4 (rows j and j+4 share banks) */
        p = c & 3;  j = ((c >> 2) & 3) | (((c >> 5) & 1) << 2);  k = ((c >> 4) & 1) | (((c >> 6) & 3) << 1);
      } else {
        p = c & (hn - 1);  j = (c >> (lg - 1)) & (n - 1);  k = c >> (2 * lg - 1);
      }
    } else {
      box = q / per_box;
      const int c = q - box * per_box;
      p = c % hn;  j = (c / hn) % n;  k = c / (hn * n);
    }
    const int ijk = 2 * p + j * jS + k * kS;
    const double *x = L.vec(box, src) + ijk;
    const double *bi = L.vec(box, VECTOR_BETA_I) + ijk, *bj = L.vec(box, VECTOR_BETA_J) + ijk, *bk = L.vec(box, VECTOR_BETA_K) + ijk;
    const double *rhs = L.vec(box, rhs_id) + ijk;
    double *out = L.vec(box, dst) + ijk;
    if (mode == 0) {
      const int color000 = (V.low[3 * box] ^ V.low[3 * box + 1] ^ V.low[3 * box + 2] ^ s) & 1;
      const int a = (j ^ k ^ color000) & 1;
      const double Ax = fv4_apply_op(x + a, bi + a, bj + a, bk + a, jS, kS, b, V.h2inv);
      const double xnew = x[a] + L.vec(box, VECTOR_DINV)[ijk + a] * (rhs[a] - Ax);
      out[1 - a] = x[1 - a];
      out[a] = xnew;
    } else {
      const double Ax0 = fv4_apply_op(x, bi, bj, bk, jS, kS, b, V.h2inv);
      const double Ax1 = fv4_apply_op(x + 1, bi + 1, bj + 1, bk + 1, jS, kS, b, V.h2inv);
      if (mode == 2) { out[0] = rhs[0] - Ax0; out[1] = rhs[1] - Ax1; }
      else {
        const double *dinv = L.vec(box, VECTOR_DINV) + ijk;
        const double r0 = x[0] + V.c1[s] * (x[0] - out[0]) + V.c2[s] * dinv[0] * (rhs[0] - Ax0);
        const double r1 = x[1] + V.c1[s] * (x[1] - out[1]) + V.c2[s] * dinv[1] * (rhs[1] - Ax1);
        out[0] = r0;  out[1] = r1;
      }
    }
  }
  __syncthreads();
  CPROF(CP_STENCIL);
}

/* the same, one thread per cell (odd box sizes) */
__device__ static void c_stencil(const CoarseLevel &V, const int mode, const int src, const int dst, const int rhs_id, const int s, const double b)
{
  if ((V.L.dim & 1) == 0) { c_stencil_pairs(V, mode, src, dst, rhs_id, s, b); return; }
  const DLevel &L = V.L;
  const int n = L.dim, cells = n * n * n, total = cells * L.nboxes;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int box = q / cells, c = q - box * cells;
    const int i = c % n, j = (c / n) % n, k = c / (n * n);
    const int ijk = i + j * L.jStride + k * L.kStride;
    const double *x = L.vec(box, src) + ijk;
    double *out = L.vec(box, dst) + ijk;
    if (mode == 0) {
      const int color000 = (V.low[3 * box] ^ V.low[3 * box + 1] ^ V.low[3 * box + 2] ^ s) & 1;
      if ((i ^ j ^ k ^ color000) & 1) { out[0] = x[0]; continue; }
    }
    const double Ax = fv4_apply_op(x, L.vec(box, VECTOR_BETA_I) + ijk, L.vec(box, VECTOR_BETA_J) + ijk, L.vec(box, VECTOR_BETA_K) + ijk, L.jStride, L.kStride, b, V.h2inv);
    const double rhs = L.vec(box, rhs_id)[ijk];
    if (mode == 2) { out[0] = rhs - Ax; continue; }
    const double dinv = L.vec(box, VECTOR_DINV)[ijk];
    if (mode == 0) out[0] = x[0] + dinv * (rhs - Ax);
    else { const double xn = x[0]; out[0] = xn + V.c1[s] * (xn - out[0]) + V.c2[s] * dinv * (rhs - Ax); }
  }
  __syncthreads();
  CPROF(CP_STENCIL);
}

__device__ static void c_zero(const DLevel &L, const int id)
{
  const int m = L.dim + 2 * L.ghosts, cells = m * m * m, total = cells * L.nboxes;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int box = q / cells, c = q - box * cells;
    const int i = c % m - L.ghosts, j = (c / m) % m - L.ghosts, k = c / (m * m) - L.ghosts;
    L.vec(box, id)[i + j * L.jStride + k * L.kStride] = 0.0;
  }
  __syncthreads();
  CPROF(CP_ZERO);
}

/* restriction.c:54-57 over the local list of the fine level */
__device__ static void c_restrict_cell(const DLevel &Lc, const int id_c, const DLevel &Lf, const int id_f, const blockCopy_type *blocks, const int n)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int e = warp; e < n; e += nwarps) {                        /* one warp per list entry */
    const blockCopy_type B = blocks[e];
    const int rj = Lf.jStride, rk = Lf.kStride;
    const double *rd = Lf.vec(B.read.box, id_f) + B.read.i + B.read.j * rj + B.read.k * rk;
    double *wr = Lc.vec(B.write.box, id_c) + B.write.i + B.write.j * Lc.jStride + B.write.k * Lc.kStride;
    const int di = B.dim.i, dj = B.dim.j, cells = di * dj * B.dim.k;
    for (int c = lane; c < cells; c += 32) {
      const int i = c % di, j = (c / di) % dj, k = c / (di * dj);
      const double *r = rd + 2 * i + 2 * j * rj + 2 * k * rk;
      wr[i + j * Lc.jStride + k * Lc.kStride] = (r[0] + r[1] + r[rj] + r[1 + rj] + r[rk] + r[1 + rk] + r[rj + rk] + r[1 + rj + rk]) * 0.125;
    }
  }
  __syncthreads();
  CPROF(CP_RESTRICT);
}

__device__ __forceinline__ void c_pro3(const double cm, const double c0, const double cp, double &lo, double &hi)
{
  const double c1 = 1.0 / 8.0;
  lo = (c0 + c1 * (cm - cp));
  hi = (c0 - c1 * (cm - cp));
}
__device__ __forceinline__ void c_pro5(const double cmm, const double cm, const double c0, const double cp, const double cpp, double &lo, double &hi)
{
  const double c2 = -3.0 / 128.0, c1 = 22.0 / 128.0;
  lo = (c0 + c1 * (cm - cp) + c2 * (cmm - cpp));
  hi = (c0 - c1 * (cm - cp) - c2 * (cmm - cpp));
}

/* interpolation_v2.c:112-172 (W=3) / interpolation_v4.c:149-238 (W=5) over the coarse level's local list */
template <int W>
__device__ static void c_interpolate(const DLevel &Lf, const int id_f, const double prescale, const DLevel &Lc, const int id_c, const blockCopy_type *blocks, const int n)
{
  constexpr int R = W / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int e = warp; e < n; e += nwarps) {                        /* one warp per list entry */
    const blockCopy_type B = blocks[e];
    const int rj = Lc.jStride, rk = Lc.kStride, wj = Lf.jStride, wk = Lf.kStride;
    const double *rd = Lc.vec(B.read.box, id_c);
    double *wr = Lf.vec(B.write.box, id_f);
    const int di = B.dim.i, dj = B.dim.j, cells = di * dj * B.dim.k;
    for (int c = lane; c < cells; c += 32) {
      const int ii = c % di, jj = (c / di) % dj, kk = c / (di * dj);
      const double *r = rd + (ii + B.read.i) + (jj + B.read.j) * rj + (kk + B.read.k) * rk;
      double fi[2][W][W], fj[2][2][W];
#pragma unroll
      for (int K = 0; K < W; K++)
#pragma unroll
      for (int J = 0; J < W; J++) {
        const double *p = r + (J - R) * rj + (K - R) * rk;
        if constexpr (W == 3) c_pro3(p[-1], p[0], p[1], fi[0][J][K], fi[1][J][K]);
        else                  c_pro5(p[-2], p[-1], p[0], p[1], p[2], fi[0][J][K], fi[1][J][K]);
      }
#pragma unroll
      for (int K = 0; K < W; K++)
#pragma unroll
      for (int I = 0; I < 2; I++) {
        if constexpr (W == 3) c_pro3(fi[I][0][K], fi[I][1][K], fi[I][2][K], fj[I][0][K], fj[I][1][K]);
        else                  c_pro5(fi[I][0][K], fi[I][1][K], fi[I][2][K], fi[I][3][K], fi[I][W - 1][K], fj[I][0][K], fj[I][1][K]);
      }
      double *w = wr + (2 * ii + B.write.i) + (2 * jj + B.write.j) * wj + (2 * kk + B.write.k) * wk;
#pragma unroll
      for (int J = 0; J < 2; J++)
#pragma unroll
      for (int I = 0; I < 2; I++) {
        double lo, hi;
        if constexpr (W == 3) c_pro3(fj[I][J][0], fj[I][J][1], fj[I][J][2], lo, hi);
        else                  c_pro5(fj[I][J][0], fj[I][J][1], fj[I][J][2], fj[I][J][3], fj[I][J][W - 1], lo, hi);
        double *w0 = w + I + J * wj;
        w0[0] = prescale * w0[0] + lo;
        w0[wk] = prescale * w0[wk] + hi;
      }
    }
  }
  __syncthreads();
  CPROF(CP_INTERP);
}

__device__ static void c_bottom_solve(const CoarseArgs &A, double *prod, double *red)
{
  const CoarseLevel &V = A.lv[A.nlevels - 1];
  BottomArgs B;
  B.L = V.L;  B.bc = V.bc[0];  B.nbc = V.n_bc[0];
  B.x_id = A.e_id;  B.R_id = A.R_id;  B.a = A.a;  B.b = A.b;  B.h2inv = V.h2inv;  B.rtol = A.rtol;  B.iters = A.krylov;
  bicgstab_solve(B, prod, red);
  __syncthreads();
  CPROF(CP_BOTTOM);
}

/* The coarsest levels (8^3, 4^3, 2^3 in the benchmark: 200 KB with all their vectors) are copied into
 * shared memory for the duration of the kernel: DLevel::base is simply pointed at the copy, so every
 * operator body works on it unchanged, at shared-memory instead of L2 latency.  Everything except the
 * read-only operator data (Dinv, betas) is written back at the end. */
__global__ void __launch_bounds__(COARSE_THREADS, 1) coarse_cycle_kernel(const __grid_constant__ CoarseArgs Ain)
{
  PDL_WAIT();
  extern __shared__ __align__(16) double dyn[];
  CoarseArgs &A = *reinterpret_cast<CoarseArgs *>(dyn);
  constexpr int ARGS_DOUBLES = (int)((sizeof(CoarseArgs) + 15) / 16) * 2;
  double *prod = dyn + ARGS_DOUBLES;
  double *red = prod + BOTTOM_MAX_CELLS + 1;
  double *pool = red + 34;
  if (threadIdx.x == 0) { s_prof_on = Ain.profile; for (int c = 0; c < CP_N; c++) s_prof[c] = 0; s_prof[CP_N] = clock64(); s_prof[CP_TOTAL] = -s_prof[CP_N]; }

  {                                                                /* stage the arguments, then patch the resident levels */
    const int *src = reinterpret_cast<const int *>(&Ain);
    int *dst = reinterpret_cast<int *>(dyn);
    for (int w = threadIdx.x; w < (int)(sizeof(CoarseArgs) / sizeof(int)); w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  for (int l = 0; l < Ain.nlevels; l++) {
    const CoarseLevel &G = Ain.lv[l];
    if (G.smem_offset < 0) continue;
    double *copy = pool + G.smem_offset;
    const double2 *g2 = reinterpret_cast<const double2 *>(G.L.base);
    double2 *c2 = reinterpret_cast<double2 *>(copy);
    for (int q = threadIdx.x; q < G.slab_doubles / 2; q += blockDim.x) c2[q] = g2[q];
    if (threadIdx.x == 0) A.lv[l].L.base = copy;
  }
  __syncthreads();
  CPROF(CP_LOAD);

#pragma unroll 1
  for (int ph = 0; ph < A.nphases; ph++) {
    const Phase P = A.prog[ph];
    const CoarseLevel &V = A.lv[P.lv];
    switch (P.op) {
      case PH_FILL:     c_fill_ghosts(V, P.a, P.b != 0, P.c != 0); break;
      case PH_STENCIL:  c_stencil(V, P.a, P.b, P.c, P.d, P.e, A.b); break;
      case PH_RESTRICT: c_restrict_cell(A.lv[P.lv + 1].L, P.a, V.L, P.b, V.restr, V.n_restr); break;
      case PH_ZERO:     c_zero(V.L, P.a); break;
      case PH_INTERP3:  c_interpolate<3>(V.L, P.a, 1.0, A.lv[P.lv + 1].L, P.b, A.lv[P.lv + 1].interp, A.lv[P.lv + 1].n_interp); break;
      case PH_INTERP5:  c_interpolate<5>(V.L, P.a, 0.0, A.lv[P.lv + 1].L, P.b, A.lv[P.lv + 1].interp, A.lv[P.lv + 1].n_interp); break;
      default:          c_bottom_solve(A, prod, red); break;
    }
  }

  __syncthreads();
  for (int l = 0; l < Ain.nlevels; l++) {                           /* write the resident levels back */
    const CoarseLevel &G = Ain.lv[l];
    if (G.smem_offset < 0) continue;
    const DLevel &L = G.L;
    const double *copy = pool + G.smem_offset;
    const int per_vec = L.volume / 2;                               /* volume is a multiple of 4 doubles */
    for (int bv = 0; bv < L.nboxes * L.nvec; bv++) {
      const int v = bv % L.nvec;
      if (v == VECTOR_DINV || v == VECTOR_BETA_I || v == VECTOR_BETA_J || v == VECTOR_BETA_K) continue;
      const double2 *c2 = reinterpret_cast<const double2 *>(copy + (size_t)bv * L.volume);
      double2 *g2 = reinterpret_cast<double2 *>(L.base + (size_t)bv * L.volume);
      for (int q = threadIdx.x; q < per_vec; q += blockDim.x) g2[q] = c2[q];
    }
  }
  __syncthreads();
  CPROF(CP_STORE);
  if (s_prof_on && threadIdx.x == 0) { s_prof[CP_TOTAL] += clock64(); for (int c = 0; c < CP_N; c++) g_coarse_prof[c] = s_prof[c]; }
}

/* ---- host side ------------------------------------------------------------------------------------ */
static int coarse_program_length(int nlevels, int ftail);
static int g_coarse_enabled = -1;
static int g_coarse_smem = 1;
static int g_coarse_profile = 0;
/* cycles per phase category of the next/last coarse kernel launch: load, fill, stencil, restrict, zero, interp, bottom, store, total */
extern "C" void hpgmg_b200_coarse_profile(int on, long long *out9)
{
  g_coarse_profile = on;
  if (out9) {
    CUDA_CHECK(cudaStreamSynchronize(g_stream));
    CUDA_CHECK(cudaMemcpyFromSymbol(out9, g_coarse_prof, sizeof(long long) * CP_N));
  }
}
/* single-block cycles pay off up to 8^3 (one cell per thread, latency-bound); 16^3 is faster as separate launches */
static long g_coarse_max_cells = 512;
extern "C" void hpgmg_b200_coarse_levels_in_smem(int on) { if ((on ? 1 : 0) != g_coarse_smem) hpgmg_graph_drop_all(NULL); g_coarse_smem = on ? 1 : 0; }
extern "C" void hpgmg_b200_use_coarse_kernel(int on) { if ((on ? 1 : 0) != g_coarse_enabled) hpgmg_graph_drop_all(NULL); g_coarse_enabled = on ? 1 : 0; }

static int level_is_coarse_eligible(const level_type *level, int is_top, int is_bottom)
{
  if (level->num_my_boxes != level->boxes_in.i * level->boxes_in.j * level->boxes_in.k) return 0;   /* every box of the level is mine */
  if (level->boundary_condition.type != BC_DIRICHLET || level->must_subtract_mean == 1) return 0;
  if (HPGMG_DEV(level)->fill_nvec != level->numVectors) return 0;
  if ((long)level->dim.i * level->dim.j * level->dim.k > g_coarse_max_cells) return 0;
  for (int s = 0; s < STENCIL_MAX_SHAPES; s++)
    if (level->exchange_ghosts[s].num_sends || level->exchange_ghosts[s].num_recvs) return 0;
  /* transfers BETWEEN chain levels must be local; those across the top of the chain run outside the kernel */
  if (!is_bottom && (level->restriction[RESTRICT_CELL].num_sends || level->interpolation.num_recvs)) return 0;
  if (!is_top && (level->restriction[RESTRICT_CELL].num_recvs || level->interpolation.num_sends)) return 0;
  return 1;
}

/* Can levels `from`..bottom of this hierarchy run in the single-block kernel?  (all of them small,
 * entirely local to this rank, Dirichlet, and a single-box bottom the BiCGStab body can solve) */
extern "C" int hpgmg_coarse_chain_eligible(mg_type *MG, int from)
{
  if (g_coarse_enabled < 0) {
    const char *e = getenv("HPGMG_B200_NO_COARSE_KERNEL");
    g_coarse_enabled = (e && atoi(e)) ? 0 : 1;
    const char *m = getenv("HPGMG_B200_COARSE_MAX_CELLS");
    if (m && atol(m) > 0) g_coarse_max_cells = atol(m) > COARSE_MAX_CELLS ? COARSE_MAX_CELLS : atol(m);
  }
  if (!g_coarse_enabled) return 0;
  const int bottom = MG->num_levels - 1;
  if (from > bottom || bottom - from + 1 > COARSE_MAX_LEVELS) return 0;
  if (coarse_program_length(bottom - from + 1, 1) > COARSE_MAX_PHASES) return 0;
  for (int l = from; l <= bottom; l++) if (!level_is_coarse_eligible(MG->levels[l], l == from, l == bottom)) return 0;
  const level_type *B = MG->levels[bottom];
  if (B->num_my_boxes != 1 || B->boxes_in.i != 1 || B->box_dim > BOTTOM_MAX_DIM || B->box_dim < 2) return 0;
  if (B->numVectors < VECTORS_RESERVED + 8) return 0;
  return 1;
}

/* ---- the phase program: the recursion of the stream version, recorded instead of executed ---- */
static void p_add(CoarseArgs &A, int op, int lv, int a = 0, int b = 0, int c = 0, int d = 0, int e = 0)
{
  if (A.nphases >= COARSE_MAX_PHASES) { fprintf(stderr, "hpgmg_b200: coarse-cycle program too long\n"); exit(1); }
  Phase P = { (unsigned char)op, (unsigned char)lv, (unsigned char)a, (unsigned char)b, (unsigned char)c, (unsigned char)d, (unsigned char)e, 0 };
  A.prog[A.nphases++] = P;
  static int repeat = -1;                       /* timing experiment: run the idempotent phases twice (warm instruction / data caches?) */
  if (repeat < 0) { const char *r = getenv("HPGMG_B200_COARSE_REPEAT"); repeat = r ? atoi(r) : 0; }
  if (repeat && (op == PH_FILL || op == PH_STENCIL) && A.nphases < COARSE_MAX_PHASES) A.prog[A.nphases++] = P;
}
static void p_smooth(CoarseArgs &A, int l, int x_id, int rhs_id)            /* smooth(): gsrb.c:24-132 / chebyshev.c:8-100 */
{
  for (int s = 0; s < 6; s++) {
    const int src = (s & 1) ? VECTOR_TEMP : x_id, dst = (s & 1) ? x_id : VECTOR_TEMP;
    p_add(A, PH_FILL, l, src, 0, 0);
    p_add(A, PH_STENCIL, l, A.smoother == HPGMG_SMOOTHER_CHEBY ? 1 : 0, src, dst, rhs_id, s);
  }
}
static void p_vcycle(CoarseArgs &A, int c)                                 /* MGVCycle: mg.c:1135-1164 */
{
  const int bottom = A.nlevels - 1;
  for (int l = c; l < bottom; l++) {
    p_smooth(A, l, A.e_id, A.R_id);
    p_add(A, PH_FILL, l, A.e_id, 0, 0);                                    /* residual(): exchange + BC on x */
    p_add(A, PH_STENCIL, l, 2, A.e_id, VECTOR_TEMP, A.R_id, 0);
    p_add(A, PH_RESTRICT, l, A.R_id, VECTOR_TEMP);
    p_add(A, PH_ZERO, l + 1, A.e_id);
  }
  p_add(A, PH_BOTTOM, bottom);
  for (int l = bottom - 1; l >= c; l--) {
    p_add(A, PH_FILL, l + 1, A.e_id, 1, 1);                                /* interpolation_v2: exchange(BOX) + apply_BCs_v2 on the coarse level */
    p_add(A, PH_INTERP3, l, A.e_id, A.e_id);
    p_smooth(A, l, A.e_id, A.R_id);
  }
}
static int coarse_program_length(int nlevels, int ftail)
{
  int v = 0, total = 0;                                                     /* phases of MGVCycle(c): 30 per non-bottom level + 1 */
  for (int c = nlevels - 1; c >= 0; c--) {
    v = 30 * (nlevels - 1 - c) + 1;
    if (ftail && c < nlevels - 1) total += 2 + v;
  }
  return ftail ? total + 2 : v;
}

extern "C" void hpgmg_coarse_cycle(mg_type *MG, int from, int mode_ftail, int zero_bottom, int e_id, int R_id, double a, double b)
{
  static CoarseArgs A;                        /* 1.3 KB: passed by value as a __grid_constant__ parameter */
  const int bottom = MG->num_levels - 1;
  A.nlevels = bottom - from + 1;
  A.mode = mode_ftail ? MODE_FTAIL : MODE_VCYCLE;
  A.smoother = hpgmg_rt_smoother();
  A.zero_bottom = zero_bottom;
  A.profile = g_coarse_profile;
  A.e_id = e_id;  A.R_id = R_id;  A.a = a;  A.b = b;  A.rtol = MG_DEFAULT_BOTTOM_NORM;
  A.krylov = hpgmg_rt_scalar_slots() + HPGMG_SLOT_KRYLOV;
  for (int l = from; l <= bottom; l++) {
    level_type *level = MG->levels[l];
    hpgmg_device_level *D = HPGMG_DEV(level);
    CoarseLevel &V = A.lv[l - from];
    V.L = D->L;  V.low = D->low;
    const int shapes[2] = { STENCIL_SHAPE_NO_CORNERS, STENCIL_SHAPE_BOX };
    for (int w = 0; w < 2; w++) {
      V.xch[w] = D->fill[shapes[w]].copies;  V.n_xch[w] = D->fill[shapes[w]].ncopies;
      V.bc[w] = D->fill[shapes[w]].bc;       V.n_bc[w] = D->fill[shapes[w]].nbc;
    }
    V.restr = D->restriction[RESTRICT_CELL][1].blocks;           V.n_restr = D->restriction[RESTRICT_CELL][1].n;
    V.interp = D->interpolation[1].blocks;                       V.n_interp = D->interpolation[1].n;
    V.h2inv = 1.0 / (level->h * level->h);
    /* Chebyshev coefficients exactly as chebyshev.c:22-40 */
    double beta = 1.000 * level->dominant_eigenvalue_of_DinvA, alpha = 0.125000 * beta;
    double theta = 0.5 * (beta + alpha), delta = 0.5 * (beta - alpha), sigma = theta / delta, rho_n = 1 / sigma;
    V.c1[0] = 0.0;  V.c2[0] = 1 / theta;
    for (int s = 1; s < 6; s++) { double rho_nm1 = rho_n; rho_n = 1.0 / (2.0 * sigma - rho_nm1); V.c1[s] = rho_n * rho_nm1; V.c2[s] = rho_n * 2.0 / delta; }
  }
  /* residency: from the bottom up while the slabs fit in the 227 KB of one SM */
  const size_t fixed = sizeof(double) * (size_t)(((sizeof(CoarseArgs) + 15) / 16) * 2 + BOTTOM_MAX_CELLS + 1 + 34);
  const size_t budget = 232320 - fixed;
  size_t used = 0;
  for (int l = bottom; l >= from; l--) {
    CoarseLevel &V = A.lv[l - from];
    const size_t doubles = (size_t)V.L.nboxes * V.L.nvec * V.L.volume;
    V.slab_doubles = (int)doubles;
    V.smem_offset = -1;
    if (g_coarse_smem && (used + doubles) * sizeof(double) <= budget && (doubles % 2) == 0 && ((uintptr_t)V.L.base % 16) == 0) {
      V.smem_offset = (int)used;
      used += doubles;
    } else break;                                                    /* keep the resident set contiguous from the bottom */
  }
  for (int l = from; l <= bottom; l++) if (A.lv[l - from].smem_offset < 0) { A.lv[l - from].smem_offset = -1; }
  A.nphases = 0;
  if (A.mode == MODE_VCYCLE) p_vcycle(A, 0);
  else {                                                                    /* MODE_FTAIL: mg.c:1285-1301 restricted to the chain */
    const int cb = A.nlevels - 1;
    if (A.zero_bottom) p_add(A, PH_ZERO, cb, A.e_id);                       /* mg.c:1285: only if the bottom is not the solve level */
    p_add(A, PH_BOTTOM, cb);
    for (int l = cb - 1; l >= 0; l--) {
      p_add(A, PH_FILL, l + 1, A.e_id, 1, 0);                               /* interpolation_v4: exchange(BOX) + apply_BCs_v4 */
      p_add(A, PH_INTERP5, l, A.e_id, A.e_id);
      p_vcycle(A, l);
    }
  }
  const size_t smem = fixed + used * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(coarse_cycle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232320));
    configured = 232320;
  }
  if (hpgmg_ablate(4)) return;
  LAUNCH(coarse_cycle_kernel, 1, COARSE_THREADS, smem, A);
}
