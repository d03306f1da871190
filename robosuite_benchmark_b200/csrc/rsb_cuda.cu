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

#define RSB_LANES 32
#define RSB_TABLE_NAME rsb_table_32
#include "rsb_kernels.inl"     /* the 32-lane build of the kernels lives in this TU; rsb_cuda16.cu holds the 16-lane build */

static thread_local std::string g_err;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)

struct rsb_batch {
  DevModel dm;                 /* device pointers fixed up */
  void *d_arena = nullptr;
  float *d_state = nullptr;
  int n = 0, device = 0, epb = 1;
  const RsbKernelTable *kt = nullptr;
  uint64_t seed = 0, env_id_base = 0;
  size_t smem_bytes = 0;
  int64_t launches = 0;
  int regs_step = 0, blocks_per_sm = 0;
  /* staging for the host-buffer entry points */
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr; uint8_t *d_done = nullptr, *d_mask = nullptr;
  float *p_act = nullptr, *p_obs = nullptr, *p_rew = nullptr; uint8_t *p_done = nullptr;
  cudaStream_t stream = nullptr;
};

static const rsb_batch *g_const_owner[64][2];     /* [device][lane-width variant]: which batch's model sits in constant memory */
#define OWNER(b) g_const_owner[(b)->device][(b)->kt->lanes == 16]

/* ------------------------------------------------------------------ C-ABI */
extern "C" {

const char *rsb_last_error(void) { return g_err.c_str(); }
int rsb_sizeof_model(void) { return (int)sizeof(rsb_model); }
int rsb_sizeof_task(void) { return (int)sizeof(rsb_task); }

void rsb_destroy(rsb_batch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  cudaDeviceSynchronize();
  if (b->device < 64 && b->kt && OWNER(b) == b) OWNER(b) = nullptr;
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
  if (const char *e = getenv("RSB_LOCKSTEP")) b->dm.lockstep = (int)strtol(e, nullptr, 0);
  if (const char *e = getenv("RSB_SOLVER_ITERS")) b->dm.solver_iters = atoi(e);        /* developer knobs (tools/): timing experiments only */
  if (const char *e = getenv("RSB_LS_ITERS")) b->dm.ls_iters = atoi(e);
  /* lane-group width: 16 lanes (two envs per warp) when the model fits (nv <= 16), else one warp per env; RSB_LANES=32|16 overrides */
  b->kt = (hm.dm.nv <= 16) ? &rsb_table_16 : &rsb_table_32;
  if (const char *e = getenv("RSB_LANES")) { int v = atoi(e); if (v == 32) b->kt = &rsb_table_32; if (v == 16 && hm.dm.nv <= 16) b->kt = &rsb_table_16; }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  size_t per_env = (size_t)hm.dm.smem_words * 4;
  /* Envs per CTA.  The kernel is latency-bound per warp, so what counts is (1) the number of WAVES the batch needs on this GPU and
     (2) even work per SM.  Try two CTAs per SM and two; take the fewest waves, then the
     smallest envs-per-CTA that still needs that many waves (4096 Lift envs on 148 SMs -> one CTA of 28 per SM, one wave). */
  int epb = 0; long best_waves = 0;
  const int gpw = 32 / b->kt->lanes;                   /* groups per warp: CTAs are whole warps (padding groups shadow the last env) */
  for (int bps = 1; bps <= 2; bps++) {
    size_t budget = (size_t)prop.sharedMemPerMultiprocessor / bps - (size_t)prop.reservedSharedMemPerBlock;
    if (budget > (size_t)prop.sharedMemPerBlockOptin) budget = (size_t)prop.sharedMemPerBlockOptin;
    int cap = (int)(budget / per_env); if (cap > b->kt->max_epb) cap = b->kt->max_epb;
    if (cap < 1) continue;
    long slots = (long)prop.multiProcessorCount * bps, waves = ((long)n_envs + slots * cap - 1) / (slots * cap);
    int e = (int)(((long)n_envs + slots * waves - 1) / (slots * waves));            /* balanced: ceil(n / (waves * CTAs per wave)) */
    e = (e + gpw - 1) / gpw * gpw; if (e > cap) e = cap / gpw * gpw; if (e < 1) continue;
    if (epb == 0 || waves < best_waves) { epb = e; best_waves = waves; }          /* ties: ONE CTA per SM = one instruction stream per SM (measured 1.27x faster than two) */
  }
  if (const char *e = getenv("RSB_EPB")) { int v = atoi(e); int cap = (int)((size_t)prop.sharedMemPerBlockOptin / per_env); if (v > 0) epb = v < cap ? v : cap; if (epb > b->kt->max_epb) epb = b->kt->max_epb; }
  if (epb < 1) { g_err = "per-env working set does not fit shared memory"; delete b; return 5; }
  b->epb = epb; b->smem_bytes = per_env * (size_t)epb;
  CK(cudaMalloc(&b->d_arena, hm.arena.size()));
  CK(cudaMemcpy(b->d_arena, hm.arena.data(), hm.arena.size(), cudaMemcpyHostToDevice));
  rsb_fixup_pointers(b->dm, b->d_arena);
  CK(cudaMalloc(&b->d_state, (size_t)n_envs * b->dm.st_words * 4));
  CK(cudaMemset(b->d_state, 0, (size_t)n_envs * b->dm.st_words * 4));
  CK(b->kt->prepare(b->smem_bytes, b->epb, &b->regs_step, &b->blocks_per_sm));
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
    case RSB_INFO_REGS_STEP: return b->regs_step; case RSB_INFO_BLOCKS_PER_SM: return b->blocks_per_sm; case RSB_INFO_LANES: return b->kt->lanes;
  }
  return -1;
}

static inline int nblocks(const rsb_batch *b) { return (b->n + b->epb - 1) / b->epb; }

/* The kernels read the model from constant memory.  Several batches (e.g. exploration and evaluation envs) may coexist in
   one process: the constant copy is re-uploaded, stream-ordered, whenever another batch used it last. */
static int bind_model(rsb_batch *b, cudaStream_t st) {
  if (b->device >= 64) { g_err = "device index >= 64"; return 1; }
  if (OWNER(b) == b) return 0;
  if (OWNER(b) != nullptr) CK(cudaDeviceSynchronize());     /* kernels of the previous owner may be in flight */
  CK(b->kt->bind(&b->dm, st));
  OWNER(b) = b;
  return 0;
}

int rsb_reset(rsb_batch *b, const uint8_t *d_mask, float *d_obs, void *stream) {
  CK(cudaSetDevice(b->device));
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  b->kt->reset(nblocks(b), b->epb, b->smem_bytes, (cudaStream_t)stream, b->d_state, d_mask, d_obs, b->seed, b->env_id_base, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_step(rsb_batch *b, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done, void *stream) {
  CK(cudaSetDevice(b->device));
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  b->kt->step(nblocks(b), b->epb, b->smem_bytes, (cudaStream_t)stream, b->d_state, d_actions, d_obs, d_reward, d_done, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_debug_substep(rsb_batch *b, const float *d_actions, int policy_step, float *d_dbg, void *stream) {
  CK(cudaSetDevice(b->device));
  int words = RSB_DBG_WORDS(b->dm.nv, b->dm.ncon_max, b->dm.nefc_max);
  if (bind_model(b, (cudaStream_t)stream)) return 1;
  b->kt->debug(nblocks(b), b->epb, b->smem_bytes, (cudaStream_t)stream, b->d_state, d_actions, policy_step, d_dbg, words, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_random_actions(rsb_batch *b, uint64_t step, float *d_actions, void *stream) {
  CK(cudaSetDevice(b->device));
  b->kt->random((cudaStream_t)stream, b->seed, b->env_id_base, step, b->dm.act_dim, d_actions, b->n);
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
