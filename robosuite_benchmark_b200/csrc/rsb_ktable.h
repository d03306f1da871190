/* rsb_ktable.h -- entry points of one lane-width build of the env kernels (rsb_kernels.inl) */
#ifndef RSB_KTABLE_H
#define RSB_KTABLE_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "rsb_devmodel.h"
#define RSB_MAX_THREADS 512    /* CTA size bound for the env kernels (register budget 128/thread): 16 groups of 32 lanes or 32 of 16 */
/* arguments of one k_step launch.  cap == 0: plain [n, .] arrays, row = env.  cap > 0 (ring mode): the arrays are the replay ring's
   (actions, next_obs, rewards, terminals, and `obs2` = observations), row = (slot0 + env) mod cap, obs2 row = (slot1 + env) mod cap. */
struct RsbStepArgs {
  float *state; const float *actions; float *obs; float *obs2; float *rew; unsigned char *done; unsigned int *iters;
  long slot0, slot1, cap; int n;
};
struct RsbKernelTable {
  int lanes, max_epb;
  cudaError_t (*bind)(const DevModel *, cudaStream_t);
  cudaError_t (*prepare)(size_t smem_bytes, int epb, int *regs, int *blocks_per_sm);
  void (*step)(int blocks, int epb, size_t smem, cudaStream_t, const RsbStepArgs *a);
  void (*reset)(int blocks, int epb, size_t smem, cudaStream_t, float *state, const unsigned char *mask, float *o, long slot0, long cap, uint64_t seed, uint64_t base, int n);
  void (*debug)(int blocks, int epb, size_t smem, cudaStream_t, float *state, const float *a, int ps, float *dbg, int words, int n);
  void (*random)(cudaStream_t, uint64_t seed, uint64_t base, uint64_t step, int act_dim, float *a, int n);
};
extern const RsbKernelTable rsb_table_32, rsb_table_16;
#endif
