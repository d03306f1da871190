/* rsb_ktable.h -- entry points of one lane-width build of the env kernels (rsb_kernels.inl) */
#ifndef RSB_KTABLE_H
#define RSB_KTABLE_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "rsb_devmodel.h"
#define RSB_MAX_THREADS 512    /* CTA size bound for the env kernels (register budget 128/thread): 16 groups of 32 lanes or 32 of 16 */
struct RsbKernelTable {
  int lanes, max_epb;
  cudaError_t (*bind)(const DevModel *, cudaStream_t);
  cudaError_t (*prepare)(size_t smem_bytes, int epb, int *regs, int *blocks_per_sm);
  void (*step)(int blocks, int epb, size_t smem, cudaStream_t, float *state, const float *a, float *o, float *r, unsigned char *d, int n);
  void (*reset)(int blocks, int epb, size_t smem, cudaStream_t, float *state, const unsigned char *mask, float *o, uint64_t seed, uint64_t base, int n);
  void (*debug)(int blocks, int epb, size_t smem, cudaStream_t, float *state, const float *a, int ps, float *dbg, int words, int n);
  void (*random)(cudaStream_t, uint64_t seed, uint64_t base, uint64_t step, int act_dim, float *a, int n);
};
extern const RsbKernelTable rsb_table_32, rsb_table_16;
#endif
