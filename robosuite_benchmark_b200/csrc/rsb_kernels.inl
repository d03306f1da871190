/*
 * rsb_kernels.inl -- the env kernels for ONE lane-group width (RSB_LANES = 32: one env per warp; 16: two envs per warp).
 * Included by rsb_cuda.cu (32) and rsb_cuda16.cu (16); everything is TU-local and exported through an RsbKernelTable.
 * The 16-lane variant halves the instruction stream per environment (the kernel is instruction-fetch bound, profiles/) and is
 * used for models with nv <= 16; the lane-cooperative code in rsb_dev.h is written against RSB_LANES throughout.
 */
#include <cuda_runtime.h>
#include "rsb_ktable.h"
#define RSB_LOCKSTEP 1
#include "rsb_dev.h"

namespace {
#define GROUPS_PER_WARP (32 / RSB_LANES)
__device__ __forceinline__ Grp make_group(int &slot) {
  const int lane = threadIdx.x & (RSB_LANES - 1); slot = threadIdx.x / RSB_LANES;
  const unsigned full = (RSB_LANES == 32) ? 0xffffffffu : ((1u << RSB_LANES) - 1u);
  Grp g{lane, full << (RSB_LANES * (slot % GROUPS_PER_WARP) % 32)};
  return g;
}

/* Row of env `e` in the caller's arrays: plain batches use row e; in ring mode (cap > 0) the arrays are the replay ring's and the
   row is slot (slot0 + e) mod cap -- the transition lands where rlkit's EnvReplayBuffer.add_sample would have put it, with no copy. */
__device__ __forceinline__ long ring_row(long slot0, long cap, int e) { if (cap <= 0) return e; long r = slot0 + e; return r >= cap ? r - cap : r; }

__global__ void __launch_bounds__(RSB_MAX_THREADS)
k_step(const RsbStepArgs a) {
  int w; Grp g = make_group(w); const int env = blockIdx.x * (blockDim.x / RSB_LANES) + w;
  const bool commit = env < a.n; const int e = commit ? env : a.n - 1;          /* padding groups shadow the last env (no stores) */
  const long r = ring_row(a.slot0, a.cap, e);
  float *obs2 = a.obs2 ? a.obs2 + (size_t)ring_row(a.slot1, a.cap, e) * c_model.obs_dim : nullptr;
  env_step(w * c_model.smem_words, g, a.state + (size_t)e * c_model.st_words, a.actions + (size_t)r * c_model.act_dim,
           a.obs + (size_t)r * c_model.obs_dim, obs2, a.rew + r, a.done + r, a.iters ? a.iters + e : nullptr, commit);
}

__global__ void __launch_bounds__(RSB_MAX_THREADS)
k_reset(float *__restrict__ state, const unsigned char *__restrict__ mask, float *__restrict__ obs, long slot0, long cap, uint64_t seed, uint64_t env_id_base, int n) {
  int w; Grp g = make_group(w); const int env = blockIdx.x * (blockDim.x / RSB_LANES) + w;
  const int e = env < n ? env : n - 1; const bool commit = env < n && (!mask || mask[e]);
  if (!__any_sync(0xffffffffu, commit)) return;       /* whole warp idle; otherwise an idle group shadows its env without storing */
  env_reset(w * c_model.smem_words, g, state + (size_t)e * c_model.st_words, seed, env_id_base + (uint64_t)e, obs + (size_t)ring_row(slot0, cap, e) * c_model.obs_dim, commit);
}

__global__ void __launch_bounds__(RSB_MAX_THREADS)
k_debug_substep(float *__restrict__ state, const float *__restrict__ actions, int policy_step, float *__restrict__ dbg, int dbg_words, int n) {
  int w; Grp g = make_group(w); const int env = blockIdx.x * (blockDim.x / RSB_LANES) + w;
  const bool commit = env < n; const int e = commit ? env : n - 1;
  const DevModel &m = c_model;
  const int so = w * m.smem_words; float *s = rsb_smem + so; float *st = state + (size_t)e * m.st_words;
  load_state(so, st, g);
  for (int i = g.lane; i < m.act_dim; i += RSB_LANES) s[m.o_act + i] = actions[(size_t)e * m.act_dim + i];
  gsync(g);
  PROF_DECL;
  substep(so, g, policy_step != 0, pt_);
  if (!commit) return;
  dump_debug(so, g, dbg + (size_t)e * dbg_words);
  store_state(so, st, g);
}

__global__ void k_random_actions(uint64_t seed, uint64_t env_id_base, uint64_t step, int act_dim, float *__restrict__ actions, int n) {
  const int nblk = (act_dim + 3) / 4; int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nblk) return;
  int env = i / nblk, blk = i - env * nblk;
  random_action_block(seed, env_id_base + (uint64_t)env, step, blk, act_dim, actions + (size_t)env * act_dim);
}

cudaError_t t_bind(const DevModel *dm, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_model, dm, sizeof(DevModel), 0, cudaMemcpyHostToDevice, st); }
cudaError_t t_prepare(size_t smem_bytes, int epb, int *regs, int *blocks_per_sm) {
  cudaError_t e;
  /* the dynamic-shared-memory limit is a property of the FUNCTION (per device), shared by every batch of the process: it only ever grows
     (an evaluation batch of a few envs must not shrink the limit under a 4096-env exploration batch) */
  static size_t s_limit[64]; int dev = 0; cudaGetDevice(&dev); const size_t want = smem_bytes;
  if (dev >= 0 && dev < 64) { if (s_limit[dev] > smem_bytes) smem_bytes = s_limit[dev]; else s_limit[dev] = smem_bytes; }
  if ((e = cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_debug_substep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  cudaFuncAttributes fa; if ((e = cudaFuncGetAttributes(&fa, k_step)) != cudaSuccess) return e; *regs = fa.numRegs;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_step, epb * RSB_LANES, want);
}
void t_step(int blocks, int epb, size_t smem, cudaStream_t st, const RsbStepArgs *a) {
  k_step<<<blocks, epb * RSB_LANES, smem, st>>>(*a); }
void t_reset(int blocks, int epb, size_t smem, cudaStream_t st, float *state, const unsigned char *mask, float *o, long slot0, long cap, uint64_t seed, uint64_t base, int n) {
  k_reset<<<blocks, epb * RSB_LANES, smem, st>>>(state, mask, o, slot0, cap, seed, base, n); }
void t_debug(int blocks, int epb, size_t smem, cudaStream_t st, float *state, const float *a, int ps, float *dbg, int words, int n) {
  k_debug_substep<<<blocks, epb * RSB_LANES, smem, st>>>(state, a, ps, dbg, words, n); }
void t_random(cudaStream_t st, uint64_t seed, uint64_t base, uint64_t step, int act_dim, float *a, int n) {
  int total = n * ((act_dim + 3) / 4); k_random_actions<<<(total + 127) / 128, 128, 0, st>>>(seed, base, step, act_dim, a, n); }
}  // namespace

extern const RsbKernelTable RSB_TABLE_NAME = {RSB_LANES, RSB_MAX_THREADS / RSB_LANES, t_bind, t_prepare, t_step, t_reset, t_debug, t_random};

#if defined(RSB_PROFILE) && RSB_LANES == 16
/* developer build only: read / clear the per-warp stage cycle counters */
extern "C" int rsb_prof_read(unsigned long long *host_out, int nwords) { return (int)cudaMemcpyFromSymbol(host_out, g_prof, (size_t)nwords * 8); }
extern "C" int rsb_prof_reset(void) { void *p; cudaGetSymbolAddress(&p, g_prof); return (int)cudaMemset(p, 0, sizeof(g_prof)); }
#endif
