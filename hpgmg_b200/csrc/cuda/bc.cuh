/*
 * bc.cuh -- device bodies of the homogeneous-Dirichlet ghost-cell extrapolations, shared by the
 * list-walking kernels in ghost.cu and the single-block bottom solver in bicgstab.cu.
 * Reference: operators/boundary_fv.c:101-250 (quadratic), :262-569 (quartic).
 */
#ifndef HPGMG_B200_BC_CUH
#define HPGMG_B200_BC_CUH
#include "common.cuh"

/* ---- boundary conditions ----------------------------------------------------------------------- */
/* 1-D quartic extrapolation of cell averages through a zero Dirichlet face (boundary_fv.c:339-340):
 * x1..x4 are the four cells nearest the boundary; near/far are the first/second ghost cell. */
__device__ __forceinline__ void quartic_pair(const double x1, const double x2, const double x3, const double x4, double &near, double &far)
{
  const double OneTwelfth = 1.0 / 12.0;
  near = OneTwelfth * (-77.0 * x1 + 43.0 * x2 - 17.0 * x3 + 3.0 * x4);
  far  = OneTwelfth * (-505.0 * x1 + 335.0 * x2 - 145.0 * x3 + 27.0 * x4);
}

/* geometry of a BC entry: for each axis, is it normal to the domain boundary, where is the nearest
 * ghost cell and which way is inward */
struct BCGeom {
  int normal[3];     /* -1 low side, +1 high side, 0 tangential   (domain normal)       */
  int t[3];          /* coordinate of the nearest ghost cell along a normal axis          */
  int inward[3];     /* +stride or -stride: one cell towards the interior                */
  int stride[3];
  int lo[3], ext[3]; /* block origin and extent                                           */
};

__device__ __forceinline__ BCGeom bc_geometry(const blockCopy_type &B, const int subtype, const int dim, const int jS, const int kS)
{
  BCGeom G;
  G.normal[0] = (subtype % 3) - 1;
  G.normal[1] = ((subtype % 9) / 3) - 1;
  G.normal[2] = (subtype / 9) - 1;
  G.stride[0] = 1;  G.stride[1] = jS;  G.stride[2] = kS;
  G.lo[0] = B.read.i;  G.lo[1] = B.read.j;  G.lo[2] = B.read.k;
  G.ext[0] = B.dim.i;  G.ext[1] = B.dim.j;  G.ext[2] = B.dim.k;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    G.t[a] = (G.normal[a] < 0) ? -1 : dim;
    G.inward[a] = (G.normal[a] < 0) ? G.stride[a] : -G.stride[a];
  }
  return G;
}

/* one BC list entry, worked on cooperatively by the nthreads threads of a thread block */
__device__ __forceinline__ void bc_v4_block(const DLevel &L, const int id, const blockCopy_type &B, const int tid, const int nthreads)
{
  const BCGeom G = bc_geometry(B, B.subtype, L.dim, L.jStride, L.kStride);
  double *__restrict__ x = L.vec(B.read.box, id);

  if (L.ghosts > 2) {                                    /* boundary_fv.c:299-306 */
    const int cells = G.ext[0] * G.ext[1] * G.ext[2];
    for (int c = tid; c < cells; c += nthreads) {
      const int i = c % G.ext[0], j = (c / G.ext[0]) % G.ext[1], k = c / (G.ext[0] * G.ext[1]);
      x[(i + G.lo[0]) + (j + G.lo[1]) * L.jStride + (k + G.lo[2]) * L.kStride] = 0.0;
    }
    __syncthreads();
  }

  /* tangential extent (axes that are not normal): one thread per tangential position */
  int text[3], cols = 1;
#pragma unroll
  for (int a = 0; a < 3; a++) { text[a] = G.normal[a] ? 1 : G.ext[a]; cols *= text[a]; }

  for (int c = tid; c < cols; c += nthreads) {
    int p[3];
    p[0] = c % text[0];  p[1] = (c / text[0]) % text[1];  p[2] = c / (text[0] * text[1]);
    int ijk = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) ijk += (G.normal[a] ? G.t[a] : (p[a] + G.lo[a])) * G.stride[a];

    /* gather the 4^m interior values: v[I][J][K], index 0..3 <-> 1..4 cells inward */
    double v[4][4][4];
    const int ni = G.normal[0] ? 4 : 1, nj = G.normal[1] ? 4 : 1, nk = G.normal[2] ? 4 : 1;
    for (int K = 0; K < nk; K++)
    for (int J = 0; J < nj; J++)
    for (int I = 0; I < ni; I++) {
      int off = ijk;
      if (G.normal[0]) off += (I + 1) * G.inward[0];
      if (G.normal[1]) off += (J + 1) * G.inward[1];
      if (G.normal[2]) off += (K + 1) * G.inward[2];
      v[I][J][K] = x[off];
    }
    /* extrapolate along i, then j, then k; after a pass the axis holds {near, far} in slots 0,1 */
    int ci = ni, cj = nj, ck = nk;
    if (G.normal[0]) {
      for (int K = 0; K < ck; K++) for (int J = 0; J < cj; J++) {
        double n, f;
        quartic_pair(v[0][J][K], v[1][J][K], v[2][J][K], v[3][J][K], n, f);
        v[0][J][K] = n;  v[1][J][K] = f;
      }
      ci = 2;
    }
    if (G.normal[1]) {
      for (int K = 0; K < ck; K++) for (int I = 0; I < ci; I++) {
        double n, f;
        quartic_pair(v[I][0][K], v[I][1][K], v[I][2][K], v[I][3][K], n, f);
        v[I][0][K] = n;  v[I][1][K] = f;
      }
      cj = 2;
    }
    if (G.normal[2]) {
      for (int J = 0; J < cj; J++) for (int I = 0; I < ci; I++) {
        double n, f;
        quartic_pair(v[I][J][0], v[I][J][1], v[I][J][2], v[I][J][3], n, f);
        v[I][J][0] = n;  v[I][J][1] = f;
      }
      ck = 2;
    }
    /* commit: slot 0 = nearest ghost, slot 1 = one further out (away from the interior) */
    for (int K = 0; K < ck; K++)
    for (int J = 0; J < cj; J++)
    for (int I = 0; I < ci; I++) {
      int off = ijk;
      if (G.normal[0]) off -= I * G.inward[0];
      if (G.normal[1]) off -= J * G.inward[1];
      if (G.normal[2]) off -= K * G.inward[2];
      x[off] = v[I][J][K];
    }
  }
}

/* quadratic: only the first ghost layer is extrapolated, deeper layers are zeroed (boundary_fv.c:101-250) */
__device__ __forceinline__ void bc_v2_block(const DLevel &L, const int id, const blockCopy_type &B, const int tid, const int nthreads)
{
  const BCGeom G = bc_geometry(B, B.subtype, L.dim, L.jStride, L.kStride);
  double *__restrict__ x = L.vec(B.read.box, id);

  if (L.ghosts > 1) {
    const int cells = G.ext[0] * G.ext[1] * G.ext[2];
    for (int c = tid; c < cells; c += nthreads) {
      const int i = c % G.ext[0], j = (c / G.ext[0]) % G.ext[1], k = c / (G.ext[0] * G.ext[1]);
      x[(i + G.lo[0]) + (j + G.lo[1]) * L.jStride + (k + G.lo[2]) * L.kStride] = 0.0;
    }
    __syncthreads();
  }
  int text[3], cols = 1, m = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) { text[a] = G.normal[a] ? 1 : G.ext[a]; cols *= text[a]; m += G.normal[a] ? 1 : 0; }
  /* the inward strides of the normal axes in ascending axis order: (dt) | (ds,dt) | (di,dj,dk) */
  int d[3] = { 0, 0, 0 }, nd = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) if (G.normal[a]) d[nd++] = G.inward[a];

  for (int c = tid; c < cols; c += nthreads) {
    int p[3];
    p[0] = c % text[0];  p[1] = (c / text[0]) % text[1];  p[2] = c / (text[0] * text[1]);
    int ijk = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) ijk += (G.normal[a] ? G.t[a] : (p[a] + G.lo[a])) * G.stride[a];
    if (m == 1) {
      const int dt = d[0];
      x[ijk] = -2.5 * x[ijk + dt] + 0.5 * x[ijk + 2 * dt];
    } else if (m == 2) {
      const int ds = d[0], dt = d[1];
      x[ijk] = 6.25 * x[ijk + ds + dt]
             - 1.25 * x[ijk + 2 * ds + dt]
             - 1.25 * x[ijk + ds + 2 * dt]
             + 0.25 * x[ijk + 2 * ds + 2 * dt];
    } else {
      const int di = d[0], dj = d[1], dk = d[2];
      x[ijk] = -15.625 * x[ijk + di + dj + dk]
              + 3.125 * x[ijk + 2 * di + dj + dk]
              + 3.125 * x[ijk + di + 2 * dj + dk]
              + 3.125 * x[ijk + di + dj + 2 * dk]
              - 0.625 * x[ijk + 2 * di + 2 * dj + dk]
              - 0.625 * x[ijk + di + 2 * dj + 2 * dk]
              - 0.625 * x[ijk + 2 * di + dj + 2 * dk]
              + 0.125 * x[ijk + 2 * di + 2 * dj + 2 * dk];
    }
  }
}


#endif
