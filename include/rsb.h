/*
 * rsb.h -- C-ABI of the B200 batched env.step / replay / SAC library (librsb_cuda.so).
 *
 * The reference has no FFI of its own: its hot path is reached through the Python API of robosuite / rlkit
 * (SURVEY.md 8b).  Each entry point below names the reference call it stands behind.  All `d_*` pointers are
 * caller-owned DEVICE memory (e.g. torch tensor .data_ptr()), all `h_*` pointers are HOST memory; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Calls are stream-ordered and do not block unless stated.
 * Return value: 0 on success, non-zero error code with text in rsb_last_error().  One host thread per batch.
 */
#ifndef RSB_H
#define RSB_H

#include <stdint.h>
#include "rsb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsb_batch rsb_batch;

enum { RSB_INFO_NENVS = 0, RSB_INFO_OBS_DIM = 1, RSB_INFO_ACT_DIM = 2, RSB_INFO_STATE_WORDS = 3, RSB_INFO_SMEM_BYTES = 4,
       RSB_INFO_DBG_WORDS = 5, RSB_INFO_NQ = 6, RSB_INFO_NV = 7, RSB_INFO_ENVS_PER_BLOCK = 8, RSB_INFO_LAUNCHES = 9,
       RSB_INFO_NCON_MAX = 10, RSB_INFO_NEFC_MAX = 11, RSB_INFO_REGS_STEP = 12, RSB_INFO_BLOCKS_PER_SM = 13, RSB_INFO_LANES = 14,
       /* event counters since rsb_create / rsb_clear_counters (reading one SYNCHRONISES the device): env control steps in which a contact
          beyond ncon_max / a constraint row beyond nefc_max was dropped (truncation is never silent: tests and bench assert 0), and
          steps asked of an already terminated episode (done == 2; the host wrapper raises ValueError like robosuite) */
       RSB_INFO_NCON_OVERFLOW = 15, RSB_INFO_NEFC_OVERFLOW = 16, RSB_INFO_STEPS_AFTER_DONE = 17,
       RSB_INFO_SOLVER_ITERATIONS = 18, RSB_INFO_LS_ITERATIONS = 19 };

const char *rsb_last_error(void);
int rsb_sizeof_model(void);
int rsb_sizeof_task(void);

/* suite.make(...) -> env (util/rlkit_utils.py:49-56): upload the compiled model + task, allocate n_envs state records
   on `device`.  Env i draws its randomness from Philox key (seed, env_id_base + i), so results do not depend on how
   envs are sharded over GPUs.  ncon_max / nefc_max bound the per-env contact and constraint-row lists (0 = defaults). */
int rsb_create(const rsb_model *model, const rsb_task *task, int n_envs, int device, uint64_t seed, uint64_t env_id_base,
               int ncon_max, int nefc_max, rsb_batch **out);
void rsb_destroy(rsb_batch *b);
int64_t rsb_info(const rsb_batch *b, int what);

/* env.reset() (util/rlkit_custom.py:418; rlkit MdpPathCollector): envs with d_mask[i] != 0 (NULL = all) are re-initialised
   and their first observation written to row i of d_obs [n_envs, obs_dim]; other rows are left untouched. */
int rsb_reset(rsb_batch *b, const uint8_t *d_mask, float *d_obs, void *stream);

/* env.step(action) (util/rlkit_custom.py:438): one 20 Hz control step = substeps x (controller + mj_step), reward, obs.
   d_actions [n, act_dim] -> d_obs [n, obs_dim], d_reward [n], d_done [n] (1 = horizon reached and not ignore_done;
   2 = the env was already done and was NOT stepped: the host wrapper raises ValueError like robosuite). */
int rsb_step(rsb_batch *b, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done, void *stream);

/* The replay ring of rlkit's EnvReplayBuffer (util/rlkit_utils.py:139-142: observations, actions, rewards, terminals, next_obs; `capacity`
   rows each), resident in HBM and owned by the caller.  The ring entry points below let the env kernels be the producer of the ring: what
   rlkit does as  path = rollout(env, policy); replay_buffer.add_paths([path])  (util/rlkit_custom.py:223-230) happens with no copy. */
typedef struct rsb_ring {
  float *observations;      /* [capacity, obs_dim] */
  float *actions;           /* [capacity, act_dim] */
  float *rewards;           /* [capacity]          */
  uint8_t *terminals;       /* [capacity]          */
  float *next_obs;          /* [capacity, obs_dim] */
  int64_t capacity;
} rsb_ring;

/* env.reset() of every env with the first observation of env i written to ring->observations[(slot0 + i) mod capacity]. */
int rsb_reset_ring(rsb_batch *b, const rsb_ring *ring, int64_t slot0, void *stream);

/* env.step for all envs with the ring as source and sink (one transition per env, row s_i = (slot0 + i) mod capacity):
   reads ring->actions[s_i] (put there by rsb_policy_act, include/rsb_sac.h), writes ring->next_obs[s_i], ring->rewards[s_i],
   ring->terminals[s_i] (= done) and, when write_next_row != 0, the same observation to ring->observations[(slot0 + n_envs + i) mod capacity]
   -- the row the next control step's policy forward reads.  A collection round of T steps from slot `top` is
   rsb_reset_ring(top); for t < T: rsb_policy_act(row top + t n); rsb_step_ring(top + t n, t + 1 < T);  then the host advances its
   ring pointer by T n (EnvReplayBuffer._advance): transition (i, t) sits in row top + t n + i. */
int rsb_step_ring(rsb_batch *b, const rsb_ring *ring, int64_t slot0, int write_next_row, void *stream);

/* Newton iterations every env spent in its last control step -> d_iters [n_envs] (device, uint32): solver-load diagnostics */
int rsb_get_iters(rsb_batch *b, uint32_t *d_iters, void *stream);
/* the solver option the kernels run with, as read from the model's <option>: out4 = {iterations, tolerance, ls_iterations, ls_tolerance} */
int rsb_get_option(const rsb_batch *b, double *out4);
int rsb_clear_counters(rsb_batch *b, void *stream);

/* same call with HOST buffers (pinned or pageable): copies actions in, steps, copies results out, synchronises. */
int rsb_step_host(rsb_batch *b, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done);
int rsb_reset_host(rsb_batch *b, const uint8_t *h_mask, float *h_obs);

/* synthetic workload of BASELINE.json configs[1]: a = tanh(N(0,1)) keyed (seed, env, step, dim) -> d_actions [n, act_dim] */
int rsb_random_actions(rsb_batch *b, uint64_t step, float *d_actions, void *stream);

/* sim.get_state()/set_state() (needed for parity from identical states): records of RSB_INFO_STATE_WORDS 32-bit words:
   qpos[nq] qvel[nv] qacc_warmstart[nv] controller_state[nrobot*80] timestep(int) episode(int) */
int rsb_get_state(rsb_batch *b, float *d_state, void *stream);
int rsb_set_state(rsb_batch *b, const float *d_state, void *stream);

/* test hook: ONE physics substep from the stored state (policy_step != 0 also runs set_goal), state written back, internals
   (M, bias, torques, contact list, constraint rows ...) dumped to d_dbg [n, RSB_INFO_DBG_WORDS] */
int rsb_debug_substep(rsb_batch *b, const float *d_actions, int policy_step, float *d_dbg, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RSB_H */
