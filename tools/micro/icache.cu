// I-cache microbenchmark (developer tool): cycles per instruction of a loop whose body is N_INSTR independent-ish FFMAs,
// for one warp on one SM, and for W warps in lockstep.  nvcc -arch=sm_100a -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void body(float *out, int iters, long long *cyc) {
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, b = 1.0001f, c = 0.5f;
  long long t0 = 0;
  for (int it = 0; it < iters + 1; it++) {
    if (it == 1) t0 = clock64();     // first pass warms L2
#pragma unroll
    for (int k = 0; k < N / 4; k++) { a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int N> void run(int threads, int blocks) {
  float *out; long long *cyc; cudaMalloc(&out, 4 * threads * blocks); cudaMalloc(&cyc, 8 * blocks);
  int iters = 20;
  body<N><<<blocks, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  body<N><<<blocks, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long h[1024]; cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = (double)h[i];
  printf("body %7d instr (%5d KB)  threads %4d blocks %3d : %.3f cycles/instr/warp\n", N, N * 16 / 1024, threads, blocks, mx / ((double)N * iters));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int blocks : {1, 148}) for (int threads : {32, 448}) {
    run<256>(threads, blocks); run<1024>(threads, blocks); run<1536>(threads, blocks); run<2048>(threads, blocks); run<3072>(threads, blocks); run<4096>(threads, blocks);
    run<6144>(threads, blocks); run<8192>(threads, blocks); run<12288>(threads, blocks); run<16384>(threads, blocks); run<32768>(threads, blocks);
  }
  return 0;
}
