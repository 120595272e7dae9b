// latency micro-benchmarks on one warp: dependent DADD / DMUL chains, integer division, generic loads from shared memory
// build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/micro/lat tools/micro/lat.cu && tools/micro/lat
// measured on B200: DADD 9.0, DMUL 8.4 cycles dependent-issue latency; runtime integer division 79-118; generic LD (shared) + DADD + cvt 67
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int n, int div)
{
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = a + b;          // dependent DADD
  long long t1 = clock64();
  for (int i = 0; i < n; i++) a = a * b;          // dependent DMUL
  long long t2 = clock64();
  int q = (int)a + 1000000007, acc = 0;
  for (int i = 0; i < n; i++) { q = q / div + 1000000007; acc += q; }   // dependent integer division by a runtime value
  long long t3 = clock64();
  int idx = threadIdx.x;
  const double *gp = sm;                           // generic pointer to shared
  double s = 0;
  for (int i = 0; i < n; i++) { s += gp[idx]; idx = ((int)s + idx) & 1023; }   // dependent generic load + DADD + cvt
  long long t4 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
  out[2 + threadIdx.x] = a + acc + s;
}
int main()
{
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
  double h[2] = {1.0, 1.0000001}; cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
  for (int warps = 1; warps <= 16; warps *= 4) {
    k<<<1, 32 * warps>>>(out, cyc, 1000, 7); cudaDeviceSynchronize();
    long long c[4]; cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
    printf("warps=%2d  cycles per dependent op: DADD %.1f  DMUL %.1f  IDIV %.1f  genericLD+DADD+cvt %.1f\n", warps, c[0] / 1000.0, c[1] / 1000.0, c[2] / 1000.0, c[3] / 1000.0);
  }
  return 0;
}
