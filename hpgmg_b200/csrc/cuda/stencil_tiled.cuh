/*
 * stencil_tiled.cuh -- the operator kernels for boxes >= 32^3: 2.5-D blocking with shared-memory halo
 * staging and a software-pipelined march along k.
 *
 * One thread block owns a TI x TJ column of cells of one box and marches over a range of k planes.
 * Per step it needs x on planes k-2..k+2 and the face coefficients on k-1..k+1 (beta_i, beta_j) and
 * k..k+1 (beta_k); these live in shared-memory ring buffers (one spare slot each) that are refilled
 * one plane ahead: global loads for the next plane are issued into registers BEFORE the current
 * plane is computed and parked in the spare slot afterwards, so there is exactly one __syncthreads
 * per plane and every HBM byte of x/beta is fetched once per tile (in-plane halo overlap only).
 *
 * Red-black without divergence: a thread owns a PAIR of cells (2p, 2p+1) of a row.  On a GSRB sweep
 * exactly one of them is active ((i^j^k^color000)&1 == 0), so every lane evaluates one stencil; the
 * other cell is copied (out-of-place GSRB, gsrb.c:65-71).  Rows are stored parity-split in shared
 * memory -- [even cells | odd cells] -- so that the stride-2 accesses of one colour are contiguous
 * across the warp (no bank conflicts), and results leave as coalesced 16-byte stores.
 * Non-coloured operators (Chebyshev, residual, apply_op) evaluate both cells of the pair.
 *
 * Arithmetic: fv4_apply_op_at (stencil.cuh) over shared-memory loaders -- the same expression tree as
 * the generic kernel, hence the same bits.
 */
#ifndef HPGMG_B200_STENCIL_TILED_CUH
#define HPGMG_B200_STENCIL_TILED_CUH

#include "stencil.cuh"

template <int TI, int TJ>
struct TileCfg {
  static constexpr int W = TI + 4;                 /* cells i0-2 .. i0+TI+1 of a row                      */
  static constexpr int HW = W / 2;                 /* pairs per row = size of each parity half            */
  static constexpr int XR = TJ + 4;                /* x rows    j0-2 .. j0+TJ+1                           */
  static constexpr int BR = TJ + 2;                /* beta rows j0-1 .. j0+TJ                             */
  static constexpr int XP = 6, BP = 4, KP = 3;     /* ring depths: 5+1, 3+1, 2+1 planes                   */
  static constexpr int NT = (TI / 2) * TJ;         /* threads                                             */
  static constexpr int XPL = XR * W, BPL = BR * W; /* doubles per plane                                   */
  static constexpr int XE = XR * HW, BE = BR * HW; /* pairs per plane                                     */
  static constexpr int XN = (XE + NT - 1) / NT, BN = (BE + NT - 1) / NT;   /* pairs per thread per plane  */
  static constexpr size_t SMEM = sizeof(double) * (size_t)(XP * XPL + 2 * BP * BPL + KP * BPL);
};

/* global -> registers: NPT pairs of one plane (rows r0.., HW pairs per row, starting at cell i0-2) */
template <int NT, int HW, int NE, int NPT>
__device__ __forceinline__ void plane_fetch(double2 (&reg)[NPT], const double *__restrict__ src, const int jS, const int tid)
{
#pragma unroll
  for (int n = 0; n < NPT; n++) {
    const int e = tid + n * NT;
    if (e < NE) {
      const int r = e / HW, q = e - r * HW;
      reg[n] = *reinterpret_cast<const double2 *>(src + r * jS + 2 * q);
    }
  }
}
/* registers -> shared, parity split: row = [HW even cells | HW odd cells] */
template <int NT, int HW, int NE, int NPT>
__device__ __forceinline__ void plane_park(const double2 (&reg)[NPT], double *__restrict__ dst, const int tid)
{
#pragma unroll
  for (int n = 0; n < NPT; n++) {
    const int e = tid + n * NT;
    if (e < NE) {
      const int r = e / HW, q = e - r * HW;
      dst[r * 2 * HW + q] = reg[n].x;
      dst[r * 2 * HW + HW + q] = reg[n].y;
    }
  }
}

/* asynchronous variant: 8-byte cp.async straight from global into the two parity halves (no register
 * staging; completion is awaited with cp.async.wait_all right before the step's barrier) */
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

/* per-thread prefetch items of one plane: global offset (doubles, from the tile origin) and shared offset
 * (doubles, from the slot base) of the even cell; the odd cell is at +1 / +HW */
template <int NT, int HW, int NE, int NPT>
struct PlaneItems {
  int goff[NPT], soff[NPT];
  __device__ __forceinline__ void init(const int jS, const int tid)
  {
#pragma unroll
    for (int n = 0; n < NPT; n++) {
      const int e = tid + n * NT;
      const int r = e / HW, q = e - r * HW;
      goff[n] = (e < NE) ? r * jS + 2 * q : -1;
      soff[n] = r * 2 * HW + q;
    }
  }
  __device__ __forceinline__ void fetch_async(const double *__restrict__ src, double *__restrict__ dst) const
  {
#pragma unroll
    for (int n = 0; n < NPT; n++)
      if (goff[n] >= 0) {
        cp_async8(dst + soff[n], src + goff[n]);
        cp_async8(dst + soff[n] + HW, src + goff[n] + 1);
      }
  }
};

/* shared-memory loader: value of the array at offset (di,dj,dk) from the thread's active cell.
 * plane[dk - DK0] points at the ring slot of plane k+dk; `row` is the thread's row inside the tile of
 * this array; same/other are the in-row offsets of the cell's own parity half and of the other half. */
template <int W, int DK0, int NPLANES>
struct TileLoader {
  const double *plane[NPLANES];
  int row, same, other;
  __device__ __forceinline__ double operator()(const int di, const int dj, const int dk) const
  {
    const double *p = plane[dk - DK0] + (row + dj) * W;
    return (di & 1) ? p[other + ((di + 1) >> 1)] : p[same + (di >> 1)];
  }
};

template <int OP, int TI, int TJ, bool ASYNC>
__global__ void __launch_bounds__((TI / 2) * TJ, 2) stencil_tiled_kernel(const StencilArgs A, const int kchunk)
{
  PDL_WAIT();
  typedef TileCfg<TI, TJ> C;
  extern __shared__ __align__(16) double smem[];
  double *xs = smem;                              /* [XP][XR][W] */
  double *bis = xs + C::XP * C::XPL;              /* [BP][BR][W] */
  double *bjs = bis + C::BP * C::BPL;             /* [BP][BR][W] */
  double *bks = bjs + C::BP * C::BPL;             /* [KP][BR][W] */

  const DLevel &L = A.L;
  const int n = L.dim, jS = L.jStride, kS = L.kStride;
  const int tiles_i = n / TI;
  const int i0 = (blockIdx.x % tiles_i) * TI, j0 = (blockIdx.x / tiles_i) * TJ;
  const int k0 = blockIdx.y * kchunk;
  const int k1 = min(n, k0 + kchunk);
  const int box = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * (TI / 2) + tx;

  /* tile origins in global memory: x tile starts at (i0-2, j0-2), beta tiles at (i0-2, j0-1) */
  const double *gx  = L.vec(box, A.x_id) + (i0 - 2) + (j0 - 2) * jS;
  const double *gbi = L.vec(box, VECTOR_BETA_I) + (i0 - 2) + (j0 - 1) * jS;
  const double *gbj = L.vec(box, VECTOR_BETA_J) + (i0 - 2) + (j0 - 1) * jS;
  const double *gbk = L.vec(box, VECTOR_BETA_K) + (i0 - 2) + (j0 - 1) * jS;

  double2 px[C::XN], pbi[C::BN], pbj[C::BN], pbk[C::BN];        /* register staging (synchronous variant only) */
  PlaneItems<C::NT, C::HW, C::XE, C::XN> ix;
  PlaneItems<C::NT, C::HW, C::BE, C::BN> ib;
  if (ASYNC) { ix.init(jS, tid); ib.init(jS, tid); }

  /* prologue: x planes k0-2..k0+2, beta_i/j planes k0-1..k0+1, beta_k planes k0, k0+1 */
  if (ASYNC) {                                        /* all 13 planes in flight at once: one memory latency, not thirteen */
    for (int kk = k0 - 2; kk <= k0 + 2; kk++) ix.fetch_async(gx + kk * kS, xs + ((kk + 2) % C::XP) * C::XPL);
    for (int kk = k0 - 1; kk <= k0 + 1; kk++) {
      ib.fetch_async(gbi + kk * kS, bis + ((kk + 1) % C::BP) * C::BPL);
      ib.fetch_async(gbj + kk * kS, bjs + ((kk + 1) % C::BP) * C::BPL);
    }
    for (int kk = k0; kk <= k0 + 1; kk++) ib.fetch_async(gbk + kk * kS, bks + (kk % C::KP) * C::BPL);
    cp_async_wait_all();
  } else {
    for (int kk = k0 - 2; kk <= k0 + 2; kk++) {
      plane_fetch<C::NT, C::HW, C::XE, C::XN>(px, gx + kk * kS, jS, tid);
      plane_park<C::NT, C::HW, C::XE, C::XN>(px, xs + ((kk + 2) % C::XP) * C::XPL, tid);
    }
    for (int kk = k0 - 1; kk <= k0 + 1; kk++) {
      plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbi, gbi + kk * kS, jS, tid);
      plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbj, gbj + kk * kS, jS, tid);
      plane_park<C::NT, C::HW, C::BE, C::BN>(pbi, bis + ((kk + 1) % C::BP) * C::BPL, tid);
      plane_park<C::NT, C::HW, C::BE, C::BN>(pbj, bjs + ((kk + 1) % C::BP) * C::BPL, tid);
    }
    for (int kk = k0; kk <= k0 + 1; kk++) {
      plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbk, gbk + kk * kS, jS, tid);
      plane_park<C::NT, C::HW, C::BE, C::BN>(pbk, bks + (kk % C::KP) * C::BPL, tid);
    }
  }
  __syncthreads();

  const int j = j0 + ty;
  const int color000 = (OP == OP_GSRB) ? ((A.low[3 * box] ^ A.low[3 * box + 1] ^ A.low[3 * box + 2] ^ A.sweep) & 1) : 0;
  const int cell = (i0 + 2 * tx) + j * jS;                       /* pair (2p, 2p+1) of row j, plane 0 */
  const double *g_rhs = (OP == OP_APPLY) ? nullptr : L.vec(box, A.rhs_id) + cell;
  const double *g_dinv = (OP == OP_GSRB || OP == OP_CHEBY) ? L.vec(box, VECTOR_DINV) + cell : nullptr;
  const double *g_xm1 = (OP == OP_CHEBY) ? L.vec(box, A.xm1_id) + cell : nullptr;
  double *g_out = L.vec(box, A.out_id) + cell;

  /* point-wise operands (rhs, Dinv, x_{n-1}) are read one plane ahead into registers */
  double2 rhs_n = make_double2(0.0, 0.0), dinv_n = make_double2(0.0, 0.0), xm_n = make_double2(0.0, 0.0);
  if (OP != OP_APPLY) rhs_n = *reinterpret_cast<const double2 *>(g_rhs + k0 * kS);
  if (OP == OP_GSRB || OP == OP_CHEBY) dinv_n = *reinterpret_cast<const double2 *>(g_dinv + k0 * kS);
  if (OP == OP_CHEBY) xm_n = *reinterpret_cast<const double2 *>(g_xm1 + k0 * kS);

  for (int k = k0; k < k1; k++) {
    /* ---- issue the loads of the next planes and of the next plane's point-wise operands ---- */
    const bool more_x = (k + 3 <= k1 + 1), more_b = (k + 2 <= k1);
    if (ASYNC) {                                      /* straight into the spare ring slots, no registers */
      if (more_x) ix.fetch_async(gx + (k + 3) * kS, xs + ((k + 5) % C::XP) * C::XPL);
      if (more_b) {
        ib.fetch_async(gbi + (k + 2) * kS, bis + ((k + 3) % C::BP) * C::BPL);
        ib.fetch_async(gbj + (k + 2) * kS, bjs + ((k + 3) % C::BP) * C::BPL);
        ib.fetch_async(gbk + (k + 2) * kS, bks + ((k + 2) % C::KP) * C::BPL);
      }
    } else {
      if (more_x) plane_fetch<C::NT, C::HW, C::XE, C::XN>(px, gx + (k + 3) * kS, jS, tid);
      if (more_b) {
        plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbi, gbi + (k + 2) * kS, jS, tid);
        plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbj, gbj + (k + 2) * kS, jS, tid);
        plane_fetch<C::NT, C::HW, C::BE, C::BN>(pbk, gbk + (k + 2) * kS, jS, tid);
      }
    }
    const double2 rhs2 = rhs_n, dinv2 = dinv_n, xm2 = xm_n;
    if (k + 1 < k1) {
      if (OP != OP_APPLY) rhs_n = *reinterpret_cast<const double2 *>(g_rhs + (k + 1) * kS);
      if (OP == OP_GSRB || OP == OP_CHEBY) dinv_n = *reinterpret_cast<const double2 *>(g_dinv + (k + 1) * kS);
      if (OP == OP_CHEBY) xm_n = *reinterpret_cast<const double2 *>(g_xm1 + (k + 1) * kS);
    }

    /* ---- compute plane k from shared memory ---- */
    TileLoader<C::W, -2, 5> X;
    TileLoader<C::W, -1, 3> BI, BJ;
    TileLoader<C::W, 0, 2> BK;
#pragma unroll
    for (int d = 0; d < 5; d++) X.plane[d] = xs + ((k + d) % C::XP) * C::XPL;            /* plane k-2+d -> slot (k-2+d+2)%XP */
#pragma unroll
    for (int d = 0; d < 3; d++) { BI.plane[d] = bis + ((k + d) % C::BP) * C::BPL; BJ.plane[d] = bjs + ((k + d) % C::BP) * C::BPL; }
#pragma unroll
    for (int d = 0; d < 2; d++) BK.plane[d] = bks + ((k + d) % C::KP) * C::BPL;
    X.row = ty + 2;  BI.row = BJ.row = BK.row = ty + 1;

    double2 out2;
    if (OP == OP_GSRB) {
      const int s = (j ^ k ^ color000) & 1;                       /* which cell of the pair is updated on this sweep */
      X.same = s * C::HW + tx + 1;  X.other = (1 - s) * C::HW + tx + s;
      BI.same = BJ.same = BK.same = X.same;  BI.other = BJ.other = BK.other = X.other;
      const double Ax = fv4_apply_op_at(X, BI, BJ, BK, A.b, A.h2inv);
      const double xc = X(0, 0, 0);
      const double xo = X.plane[2][X.row * C::W + (1 - s) * C::HW + tx + 1];   /* the pair's other cell: copied */
      const double xnew = xc + (s ? dinv2.y : dinv2.x) * ((s ? rhs2.y : rhs2.x) - Ax);
      out2 = s ? make_double2(xo, xnew) : make_double2(xnew, xo);
    } else {
      double r[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        X.same = s * C::HW + tx + 1;  X.other = (1 - s) * C::HW + tx + s;
        BI.same = BJ.same = BK.same = X.same;  BI.other = BJ.other = BK.other = X.other;
        const double Ax = fv4_apply_op_at(X, BI, BJ, BK, A.b, A.h2inv);
        if (OP == OP_APPLY) r[s] = Ax;
        else if (OP == OP_RESIDUAL) r[s] = (s ? rhs2.y : rhs2.x) - Ax;
        else {                                                    /* OP_CHEBY, chebyshev.c:90 */
          const double xn = X(0, 0, 0);
          r[s] = xn + A.c1 * (xn - (s ? xm2.y : xm2.x)) + A.c2 * (s ? dinv2.y : dinv2.x) * ((s ? rhs2.y : rhs2.x) - Ax);
        }
      }
      out2 = make_double2(r[0], r[1]);
    }
    *reinterpret_cast<double2 *>(g_out + k * kS) = out2;

    /* ---- park the prefetched planes in the spare ring slots (nobody reads them during this step) ---- */
    if (ASYNC) {
      cp_async_wait_all();
    } else {
      if (more_x) plane_park<C::NT, C::HW, C::XE, C::XN>(px, xs + ((k + 5) % C::XP) * C::XPL, tid);
      if (more_b) {
        plane_park<C::NT, C::HW, C::BE, C::BN>(pbi, bis + ((k + 3) % C::BP) * C::BPL, tid);
        plane_park<C::NT, C::HW, C::BE, C::BN>(pbj, bjs + ((k + 3) % C::BP) * C::BPL, tid);
        plane_park<C::NT, C::HW, C::BE, C::BN>(pbk, bks + ((k + 2) % C::KP) * C::BPL, tid);
      }
    }
    __syncthreads();
  }
}

#endif
