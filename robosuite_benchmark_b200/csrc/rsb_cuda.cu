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
#include <mutex>
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
  unsigned int *d_counters = nullptr;   /* [8] event counters (DevModel::counters) */
  unsigned int *d_iters = nullptr;      /* [n] Newton iterations of every env in its last control step */
  cudaStream_t last_stream = nullptr; bool launched = false;   /* stream of the last launch that touched d_state (see order_after_previous) */
  /* staging for the host-buffer entry points */
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr; uint8_t *d_done = nullptr, *d_mask = nullptr;
  float *p_act = nullptr, *p_obs = nullptr, *p_rew = nullptr; uint8_t *p_done = nullptr;
  cudaStream_t stream = nullptr;
};

static const rsb_batch *g_const_owner[64][2];     /* [device][lane-width variant]: which batch's model sits in constant memory */
static std::mutex g_owner_mutex;                   /* the owner table is process-wide: bind + launch of one call are atomic against other host threads */
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
  cudaFree(b->d_arena); cudaFree(b->d_state); cudaFree(b->d_counters); cudaFree(b->d_iters); cudaFree(b->d_act); cudaFree(b->d_obs); cudaFree(b->d_rew); cudaFree(b->d_done); cudaFree(b->d_mask);
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
  CK(cudaMalloc(&b->d_counters, 8 * sizeof(unsigned int))); CK(cudaMemset(b->d_counters, 0, 8 * sizeof(unsigned int)));
  CK(cudaMalloc(&b->d_iters, (size_t)n_envs * sizeof(unsigned int))); CK(cudaMemset(b->d_iters, 0, (size_t)n_envs * sizeof(unsigned int)));
  b->dm.counters = b->d_counters;
  CK(b->kt->prepare(b->smem_bytes, b->epb, &b->regs_step, &b->blocks_per_sm));
  CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  *out = b; return 0;
}

static int64_t read_counter(const rsb_batch *b, int k) {
  unsigned int v = 0; cudaSetDevice(b->device);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&v, b->d_counters + k, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

int64_t rsb_info(const rsb_batch *b, int what) {
  switch (what) {
    case RSB_INFO_NCON_OVERFLOW: return read_counter(b, 0); case RSB_INFO_NEFC_OVERFLOW: return read_counter(b, 1); case RSB_INFO_STEPS_AFTER_DONE: return read_counter(b, 2);
    case RSB_INFO_SOLVER_ITERATIONS: return b->dm.solver_iters; case RSB_INFO_LS_ITERATIONS: return b->dm.ls_iters;
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
/* Launches of one batch are ordered by their stream.  The device-pointer entry points run on the caller's stream, the host-buffer entry
   points on the batch's private stream; both touch the state records (and the constant-memory model), so when a call arrives on another
   stream than the previous one, the previous stream is drained first (rare: only when the two families are mixed on one batch). */
static int order_after_previous(rsb_batch *b, cudaStream_t st) {
  if (b->launched && b->last_stream != st) CK(cudaStreamSynchronize(b->last_stream));
  b->last_stream = st; b->launched = true; return 0;
}

static int launch_reset(rsb_batch *b, const uint8_t *d_mask, float *d_obs, long slot0, long cap, cudaStream_t st) {
  CK(cudaSetDevice(b->device));
  std::lock_guard<std::mutex> lock(g_owner_mutex);
  if (order_after_previous(b, st) || bind_model(b, st)) return 1;
  b->kt->reset(nblocks(b), b->epb, b->smem_bytes, st, b->d_state, d_mask, d_obs, slot0, cap, b->seed, b->env_id_base, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}
static int launch_step(rsb_batch *b, RsbStepArgs &a, cudaStream_t st) {
  CK(cudaSetDevice(b->device));
  std::lock_guard<std::mutex> lock(g_owner_mutex);
  if (order_after_previous(b, st) || bind_model(b, st)) return 1;
  a.state = b->d_state; a.iters = b->d_iters; a.n = b->n;
  b->kt->step(nblocks(b), b->epb, b->smem_bytes, st, &a);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_reset(rsb_batch *b, const uint8_t *d_mask, float *d_obs, void *stream) { return launch_reset(b, d_mask, d_obs, 0, 0, (cudaStream_t)stream); }

int rsb_step(rsb_batch *b, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done, void *stream) {
  RsbStepArgs a{}; a.actions = d_actions; a.obs = d_obs; a.obs2 = nullptr; a.rew = d_reward; a.done = d_done; a.slot0 = a.slot1 = a.cap = 0;
  return launch_step(b, a, (cudaStream_t)stream);
}

static int check_ring(const rsb_batch *b, const rsb_ring *r, int64_t slot0) {
  if (!r || !r->observations || !r->actions || !r->rewards || !r->terminals || !r->next_obs) { g_err = "ring: null array"; return 1; }
  if (r->capacity < b->n) { g_err = "ring: capacity smaller than the env batch"; return 1; }
  if (slot0 < 0 || slot0 >= r->capacity) { g_err = "ring: slot0 out of range"; return 1; }
  return 0;
}
int rsb_reset_ring(rsb_batch *b, const rsb_ring *ring, int64_t slot0, void *stream) {
  if (check_ring(b, ring, slot0)) return 6;
  return launch_reset(b, nullptr, ring->observations, (long)slot0, (long)ring->capacity, (cudaStream_t)stream);
}
int rsb_step_ring(rsb_batch *b, const rsb_ring *ring, int64_t slot0, int write_next_row, void *stream) {
  if (check_ring(b, ring, slot0)) return 6;
  RsbStepArgs a{}; a.actions = ring->actions; a.obs = ring->next_obs; a.rew = ring->rewards; a.done = ring->terminals;
  a.obs2 = write_next_row ? ring->observations : nullptr;
  a.slot0 = (long)slot0; a.slot1 = (long)((slot0 + b->n) % ring->capacity); a.cap = (long)ring->capacity;
  return launch_step(b, a, (cudaStream_t)stream);
}
int rsb_get_iters(rsb_batch *b, uint32_t *d_iters, void *stream) {
  CK(cudaSetDevice(b->device));
  CK(cudaMemcpyAsync(d_iters, b->d_iters, (size_t)b->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0;
}
int rsb_get_option(const rsb_batch *b, double *out4) {
  out4[0] = b->dm.solver_iters; out4[1] = b->dm.solver_tol; out4[2] = b->dm.ls_iters; out4[3] = b->dm.ls_tol; return 0;
}
int rsb_clear_counters(rsb_batch *b, void *stream) {
  CK(cudaSetDevice(b->device)); CK(cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned int), (cudaStream_t)stream)); return 0;
}

int rsb_debug_substep(rsb_batch *b, const float *d_actions, int policy_step, float *d_dbg, void *stream) {
  CK(cudaSetDevice(b->device));
  int words = RSB_DBG_WORDS(b->dm.nv, b->dm.ncon_max, b->dm.nefc_max);
  std::lock_guard<std::mutex> lock(g_owner_mutex);
  if (order_after_previous(b, (cudaStream_t)stream) || bind_model(b, (cudaStream_t)stream)) return 1;
  b->kt->debug(nblocks(b), b->epb, b->smem_bytes, (cudaStream_t)stream, b->d_state, d_actions, policy_step, d_dbg, words, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_random_actions(rsb_batch *b, uint64_t step, float *d_actions, void *stream) {
  CK(cudaSetDevice(b->device));
  b->kt->random((cudaStream_t)stream, b->seed, b->env_id_base, step, b->dm.act_dim, d_actions, b->n);
  b->launches++; CK(cudaGetLastError()); return 0;
}

int rsb_get_state(rsb_batch *b, float *d_state, void *stream) {
  CK(cudaSetDevice(b->device)); if (order_after_previous(b, (cudaStream_t)stream)) return 1;
  CK(cudaMemcpyAsync(d_state, b->d_state, (size_t)b->n * b->dm.st_words * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0;
}
int rsb_set_state(rsb_batch *b, const float *d_state, void *stream) {
  CK(cudaSetDevice(b->device)); if (order_after_previous(b, (cudaStream_t)stream)) return 1;
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
