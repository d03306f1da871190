/*
 * rsb_cuda.cu -- sm_100a kernels + C-ABI (include/rsb.h) for the batched env.step hot path.
 *
 * Mapping: one warp per environment, RSB_EPB environments per CTA, each warp's working set (poses, inertias, M, J,
 * constraint rows, solver vectors: DevModel::smem_words floats) in its own slice of dynamic shared memory for all
 * 25 substeps of the control step.  HBM traffic per env and control step is the state record in, the state record +
 * observation row + reward/done out (DESIGN.md "algorithmic bytes").  No tensor cores: this path is CUDA-core fp32.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rsb.h"

#define RSB_MAX_EPB 16         /* envs (warps) per CTA, chosen at create(): as many as fit shared memory, so that with lockstep
                                  stages ONE CTA per SM shares a single instruction stream */
#define RSB_LOCKSTEP 1
#include "rsb_dev.h"

static thread_local std::string g_err;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)

struct rsb_batch {
  DevModel dm;                 /* device pointers fixed up */
  void *d_arena = nullptr;
  float *d_state = nullptr;
  int n = 0, device = 0, epb = 1;
  uint64_t seed = 0, env_id_base = 0;
  size_t smem_bytes = 0;
  int64_t launches = 0;
  int regs_step = 0, blocks_per_sm = 0;
  /* staging for the host-buffer entry points */
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr; uint8_t *d_done = nullptr, *d_mask = nullptr;
  float *p_act = nullptr, *p_obs = nullptr, *p_rew = nullptr; uint8_t *p_done = nullptr;
  cudaStream_t stream = nullptr;
};

/* ------------------------------------------------------------------ kernels */
__global__ void __launch_bounds__(RSB_MAX_EPB * 32)
k_step(float *__restrict__ state, const float *__restrict__ actions, float *__restrict__ obs,
       float *__restrict__ rew, unsigned char *__restrict__ done, int n) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, env = blockIdx.x * (blockDim.x >> 5) + w;
  const bool commit = env < n; const int e = commit ? env : n - 1;          /* padding warps shadow the last env (no stores) */
  Grp g{lane, 0xffffffffu};
  env_step(w * c_model.smem_words, g, state + (size_t)e * c_model.st_words, actions + (size_t)e * c_model.act_dim,
           obs + (size_t)e * c_model.obs_dim, rew + e, done + e, commit);
}

__global__ void __launch_bounds__(RSB_MAX_EPB * 32)
k_reset(float *__restrict__ state, const unsigned char *__restrict__ mask, float *__restrict__ obs,
        uint64_t seed, uint64_t env_id_base, int n) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, env = blockIdx.x * (blockDim.x >> 5) + w;
  if (env >= n) return;
  if (mask && !mask[env]) return;
  Grp g{lane, 0xffffffffu};
  env_reset(w * c_model.smem_words, g, state + (size_t)env * c_model.st_words, seed, env_id_base + (uint64_t)env, obs + (size_t)env * c_model.obs_dim);
}

__global__ void __launch_bounds__(RSB_MAX_EPB * 32)
k_debug_substep(float *__restrict__ state, const float *__restrict__ actions, int policy_step,
                float *__restrict__ dbg, int dbg_words, int n) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, env = blockIdx.x * (blockDim.x >> 5) + w;
  const bool commit = env < n; const int e = commit ? env : n - 1;
  const DevModel &m = c_model;
  Grp g{lane, 0xffffffffu}; const int so = w * m.smem_words; float *s = rsb_smem + so; float *st = state + (size_t)e * m.st_words;
  load_state(so, st, g);
  for (int i = lane; i < m.act_dim; i += 32) s[m.o_act + i] = actions[(size_t)e * m.act_dim + i];
  gsync(g);
  substep(so, g, policy_step != 0);
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

static const rsb_batch *g_const_owner[64];

/* ------------------------------------------------------------------ C-ABI */
extern "C" {

const char *rsb_last_error(void) { return g_err.c_str(); }
int rsb_sizeof_model(void) { return (int)sizeof(rsb_model); }
int rsb_sizeof_task(void) { return (int)sizeof(rsb_task); }

void rsb_destroy(rsb_batch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  cudaDeviceSynchronize();
  if (b->device < 64 && g_const_owner[b->device] == b) g_const_owner[b->device] = nullptr;
  cudaFree(b->d_arena); cudaFree(b->d_state); cudaFree(b->d_act); cudaFree(b->d_obs); cudaFree(b->d_rew); cudaFree(b->d_done); cudaFree(b->d_mask);
  cudaFreeHost(b->p_act); cudaFreeHost(b->p_obs); cudaFreeHost(b->p_rew); cudaFreeHost(b->p_done);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int rsb_create(const rsb_model *model, const rsb_task *task, int n_envs, int device, uint64_t seed, uint64_t env_id_base,
               int ncon_max, int nefc_max, rsb_batch **out) {
  *out = nullptr;
  if (n_envs <= 0) { g_err = "n_envs must be positive"; return 2; }
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device: the batched env.step path has no CPU fallback"; return 3; }
  CK(cudaSetDevice(device));
  if (ncon_max <= 0) ncon_max = 16;
  if (nefc_max <= 0) nefc_max = 64;
  RsbHostModel hm;
  if (!rsb_build_host_model(model, task, ncon_max, nefc_max, hm)) { g_err = hm.error; return 4; }
  rsb_batch *b = new rsb_batch(); b->n = n_envs; b->device = device; b->seed = seed; b->env_id_base = env_id_base;
  b->dm = hm.dm;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  size_t per_env = (size_t)hm.dm.smem_words * 4;
  int epb = (int)((size_t)prop.sharedMemPerBlockOptin / per_env); if (epb > RSB_MAX_EPB) epb = RSB_MAX_EPB;
  /* two CTAs per SM (each keeps its warps in lockstep) balance instruction-fetch sharing against barrier stalls and tail waves */
  { int half = (int)(((size_t)prop.sharedMemPerMultiprocessor / 2 - 1024) / per_env); if (half >= 4 && half < epb) epb = half; }
  if (const char *e = getenv("RSB_EPB")) { int v = atoi(e); if (v > 0 && v < epb) epb = v; }
  if (epb < 1) { g_err = "per-env working set does not fit shared memory"; delete b; return 5; }
  b->epb = epb; b->smem_bytes = per_env * (size_t)epb;
  CK(cudaMalloc(&b->d_arena, hm.arena.size()));
  CK(cudaMemcpy(b->d_arena, hm.arena.data(), hm.arena.size(), cudaMemcpyHostToDevice));
  rsb_fixup_pointers(b->dm, b->d_arena);
  CK(cudaMalloc(&b->d_state, (size_t)n_envs * b->dm.st_words * 4));
  CK(cudaMemset(b->d_state, 0, (size_t)n_envs * b->dm.st_words * 4));
  CK(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  CK(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  CK(cudaFuncSetAttribute(k_debug_substep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_step)); b->regs_step = fa.numRegs;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->blocks_per_sm, k_step, b->epb * 32, b->smem_bytes));
  CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  *out = b; return 0;
}

int64_t rsb_info(const rsb_batch *b, int what) {
  switch (what) {
    case RSB_INFO_NENVS: return b->n; case RSB_INFO_OBS_DIM: return b->dm.obs_dim; case RSB_INFO_ACT_DIM: return b->dm.act_dim;
    case RSB_INFO_STATE_WORDS: return b->dm.st_words; case RSB_INFO_SMEM_BYTES: return (int64_t)b->dm.smem_words * 4;
    case RSB_INFO_DBG_WORDS: return RSB_DBG_WORDS(b->dm.nv, b->dm.ncon_max, b->dm.nefc_max); case RSB_INFO_NQ: return b->dm.nq; case RSB_INFO_NV: return b->dm.nv;
    case RSB_INFO_ENVS_PER_BLOCK: return b->epb; case RSB_INFO_LAUNCHES: return b->launches;
    case RSB_INFO_NCON_MAX: return b->dm.ncon_max; case RSB_INFO_NEFC_MAX: return b->dm.nefc_max;
    case RSB_INFO_REGS_STEP: return b->regs_step; case RSB_INFO_BLOCKS_PER_SM: return b->blocks_per_sm;
  }
  return -1;
}

static inline int nblocks(const rsb_batch *b) { return (b->n + b->epb - 1) / b->epb; }

/* The kernels read the model from constant memory.  Several batches (e.g. exploration and evaluation envs) may coexist in
   one process: the constant copy is re-uploaded, stream-ordered, whenever another batch used it last. */
static int bind_model(rsb_batch *b, cudaStream_t st) {
  if (b->device < 64 && g_const_owner[b->device] == b) return 0;
  if (b->device < 64 && g_const_owner[b->device] != nullptr) CK(cudaDeviceSynchronize());     /* kernels of the previous owner may be in flight */
  CK(cudaMemcpyToSymbolAsync(c_model, &b->dm, sizeof(DevModel), 0, cudaMemcpyHostToDevice, st));
  if (b->device < 64) g_const_owner[b->device] = b;
  return 0;
}

int rsb_reset(rsb_batch *b, const uint8_t *d_mask, float *d_obs, void *stream) {
  CK(cudaSetDevice(b->device));
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  k_reset<<<nblocks(b), b->epb * 32, b->smem_bytes, (cudaStream_t)stream>>>(b->d_state, d_mask, d_obs, b->seed, b->env_id_base, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_step(rsb_batch *b, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done, void *stream) {
  CK(cudaSetDevice(b->device));
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  k_step<<<nblocks(b), b->epb * 32, b->smem_bytes, (cudaStream_t)stream>>>(b->d_state, d_actions, d_obs, d_reward, d_done, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_debug_substep(rsb_batch *b, const float *d_actions, int policy_step, float *d_dbg, void *stream) {
  CK(cudaSetDevice(b->device));
  int words = RSB_DBG_WORDS(b->dm.nv, b->dm.ncon_max, b->dm.nefc_max);
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  k_debug_substep<<<nblocks(b), b->epb * 32, b->smem_bytes, (cudaStream_t)stream>>>(b->d_state, d_actions, policy_step, d_dbg, words, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_random_actions(rsb_batch *b, uint64_t step, float *d_actions, void *stream) {
  CK(cudaSetDevice(b->device));
  int total = b->n * ((b->dm.act_dim + 3) / 4);
  k_random_actions<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b->seed, b->env_id_base, step, b->dm.act_dim, d_actions, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_get_state(rsb_batch *b, float *d_state, void *stream) {
  CK(cudaSetDevice(b->device));
  CK(cudaMemcpyAsync(d_state, b->d_state, (size_t)b->n * b->dm.st_words * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0;
}
int rsb_set_state(rsb_batch *b, const float *d_state, void *stream) {
  CK(cudaSetDevice(b->device));
  CK(cudaMemcpyAsync(b->d_state, d_state, (size_t)b->n * b->dm.st_words * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0;
}

static int ensure_staging(rsb_batch *b) {
  if (b->d_act) return 0;
  size_t n = (size_t)b->n;
  CK(cudaMalloc(&b->d_act, n * b->dm.act_dim * 4)); CK(cudaMalloc(&b->d_obs, n * b->dm.obs_dim * 4)); CK(cudaMalloc(&b->d_rew, n * 4));
  CK(cudaMalloc(&b->d_done, n)); CK(cudaMalloc(&b->d_mask, n));
  CK(cudaMallocHost(&b->p_act, n * b->dm.act_dim * 4)); CK(cudaMallocHost(&b->p_obs, n * b->dm.obs_dim * 4)); CK(cudaMallocHost(&b->p_rew, n * 4)); CK(cudaMallocHost(&b->p_done, n));
  return 0;
}

int rsb_step_host(rsb_batch *b, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done) {
  CK(cudaSetDevice(b->device));
  if (ensure_staging(b)) return 1;
  size_t n = (size_t)b->n;
  memcpy(b->p_act, h_actions, n * b->dm.act_dim * 4);
  CK(cudaMemcpyAsync(b->d_act, b->p_act, n * b->dm.act_dim * 4, cudaMemcpyHostToDevice, b->stream));
  if (rsb_step(b, b->d_act, b->d_obs, b->d_rew, b->d_done, b->stream)) return 1;
  CK(cudaMemcpyAsync(b->p_obs, b->d_obs, n * b->dm.obs_dim * 4, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaMemcpyAsync(b->p_rew, b->d_rew, n * 4, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaMemcpyAsync(b->p_done, b->d_done, n, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  memcpy(h_obs, b->p_obs, n * b->dm.obs_dim * 4); memcpy(h_reward, b->p_rew, n * 4); memcpy(h_done, b->p_done, n);
  return 0;
}

int rsb_reset_host(rsb_batch *b, const uint8_t *h_mask, float *h_obs) {
  CK(cudaSetDevice(b->device));
  if (ensure_staging(b)) return 1;
  size_t n = (size_t)b->n;
  if (h_mask) CK(cudaMemcpyAsync(b->d_mask, h_mask, n, cudaMemcpyHostToDevice, b->stream));
  /* rows of envs that are not reset must keep the caller's values: stage the caller's buffer first */
  memcpy(b->p_obs, h_obs, n * b->dm.obs_dim * 4);
  CK(cudaMemcpyAsync(b->d_obs, b->p_obs, n * b->dm.obs_dim * 4, cudaMemcpyHostToDevice, b->stream));
  if (rsb_reset(b, h_mask ? b->d_mask : nullptr, b->d_obs, b->stream)) return 1;
  CK(cudaMemcpyAsync(b->p_obs, b->d_obs, n * b->dm.obs_dim * 4, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  memcpy(h_obs, b->p_obs, n * b->dm.obs_dim * 4);
  return 0;
}

}  /* extern "C" */
