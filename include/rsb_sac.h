/*
 * rsb_sac.h -- C-ABI of the replay-sampling and SAC-update kernels (part of librsb_cuda.so).
 *
 * Reference interfaces replaced (un-vendored rlkit @ b7f97b2, reached from util/rlkit_custom.py:235-238):
 *   EnvReplayBuffer.random_batch(batch_size)          -> rsb_replay_sample
 *   TanhGaussianPolicy.forward / TanhNormal.rsample   -> rsb_normal + rsb_head_fwd / rsb_head_bwd
 *   SACTrainer.train_from_torch (losses)              -> rsb_sac_losses
 *   torch.optim.Adam x4 + soft_update_from_to         -> rsb_adam_polyak
 * The dense products between these calls are rsb_gemm_tf32 (include/rsb_gemm.h: tcgen05 TF32, bias / ReLU / ReLU-backward fused).  All pointers are DEVICE pointers, `stream` is a
 * cudaStream_t as void*; every call is stream-ordered, non-blocking and CUDA-graph capturable.  Returns 0 or an error code.
 */
#ifndef RSB_SAC_H
#define RSB_SAC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char *rsb_sac_last_error(void);
/* ring arrays in rlkit's layout: obs[cap,O] act[cap,A] rew[cap] term[cap](u8) next_obs[cap,O]; `size` = filled rows.
   Row b of the batch takes ring row idx_b = mulhi32(Philox4x32-10(key=seed, ctr=(b, step_lo, step_hi, 0xB0FFE7)).word0, size). */
int rsb_replay_sample(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, int size, int obs_dim, int act_dim,
                      uint64_t seed, uint64_t step, int batch, float *b_obs, int ld_obs, float *b_act, float *b_rew, float *b_term, float *b_next, int ld_next, int *b_idx, void *stream);
/* inputs of one update from the sampled batch (SACTrainer.train_from_torch's obs/actions/next_obs unpacking, util/rlkit_custom.py:238): Xp[2B,O] = [obs; next_obs],
   act[B,A] -> XQ[2B,O+A] = [(obs, -); (obs, act)], XT[B,O+A] = (next_obs, -); clears the loss accumulators d_sums[nsums] and the log-alpha gradient */
int rsb_sac_prepare(const float *d_xp, const float *d_act, float *d_xq, float *d_xt, float *d_sums, int nsums, float *d_g_log_alpha, int batch, int obs_dim, int act_dim, void *stream);
int rsb_normal(uint64_t seed, uint64_t step, uint32_t stream_id, int n, float *d_out, void *stream);
int rsb_bias_relu(float *d_x, const float *d_bias, int rows, int cols, int relu, int nmat, long mat_stride, int bias_stride, void *stream);
int rsb_relu_bwd(float *d_dy, const float *d_y, long n, void *stream);
int rsb_colsum(const float *d_dy, int r0, int r1, int cols, float *d_db, int nmat, long mat_stride, int db_stride, void *stream);
int rsb_head_fwd(const float *d_out, const float *d_eps, int rows, int act_dim, float *d_a, float *d_logpi, float *dst0, int ld0, int r0lo, int r0hi, float *dst1, int ld1, int r1lo, int r1hi, void *stream);
int rsb_head_bwd(const float *d_out, const float *d_eps, const float *d_a, int rows, int batch, int act_dim, const float *d_alpha, const float *d_ga, int ld_ga, float *d_dout, void *stream);
int rsb_sac_losses(const float *d_q, const float *d_qt, const float *d_logpi, const float *d_rew, const float *d_term, const float *d_alpha, float reward_scale, float discount,
                   float target_entropy, int batch, float *d_dq, float *d_y, float *d_sums, float *d_galpha, void *stream);
/* Adam's bias corrections advance on the device (d_bc [4] = {1 - b1^t, sqrt(1 - b2^t), b1^t, b2^t}; {0,0,1,1} at start): rsb_adam_tick advances them by one
   step (and, if given, increments the device-resident update counter); rsb_adam_polyak with tick != 0 does the tick itself first, with tick == 0 it
   expects the caller to have ticked earlier in the same update (off the update's critical path). */
int rsb_adam_tick(double *d_bc, float b1, float b2, int64_t *d_step_counter, void *stream);
int rsb_adam_polyak(float *d_p, const float *d_g, float *d_m, float *d_v, long n, double lr_pi, double lr_q, float b1, float b2, float eps, double *d_bc,
                    float *d_tgt, long tgt_begin, long tgt_end, float tau, int do_soft, float *d_alpha, long log_alpha_idx, int tick, void *stream);

/* ---- fused narrow layers (csrc/rsb_sac_fused.cu): the policy's last layer, the Q networks' last layer and their input gradients are not
   GEMM-shaped (2A <= 32 columns; one column); each is fused with the elementwise kernel next to it, exact fp32 on the CUDA cores.
   d_h2 [rows, 256] hidden activations, d_w2 [256, 2A], d_b2 [2A]; the rest as rsb_head_fwd / rsb_head_bwd / rsb_sac_losses. */
int rsb_policy_head_fwd(const float *d_h2, const float *d_w2, const float *d_b2, const float *d_eps, int rows, int act_dim, float *d_out, float *d_a, float *d_logpi,
                        float *dst0, int ld0, int r0lo, int r0hi, float *dst1, int ld1, int r1lo, int r1hi, void *stream);
/* d_h2q [2, 2B, 256] (rows [0,B): (obs, a_new); [B,2B): (obs, act)), d_wq2 [2, 256], d_bq2 [2]; d_h2t [2, B, 256], d_wt2 [2, 256], d_bt2 [2] (target nets);
   writes q [2, 2B], qt [2, B], dq [2, 2B], dH2q [2, 2B, 256] = (dq w2^T) masked by H2q > 0, y [B], the loss sums and d alpha_loss / d log_alpha */
int rsb_q_losses(const float *d_h2q, const float *d_wq2, const float *d_bq2, const float *d_h2t, const float *d_wt2, const float *d_bt2, const float *d_logpi,
                 const float *d_rew, const float *d_term, const float *d_alpha, float reward_scale, float discount, float target_entropy, int batch,
                 float *d_q, float *d_qt, float *d_dq, float *d_dh2q, float *d_y, float *d_sums, float *d_galpha, void *stream);
/* rows [0, batch) only: d_dout [batch, 2A], d_dh2 [batch, 256] = (dOUT W2^T) masked by H2 > 0 */
int rsb_policy_head_bwd(const float *d_out, const float *d_eps, const float *d_a, const float *d_h2, const float *d_w2, int batch, int act_dim, const float *d_alpha,
                        const float *d_ga, int ld_ga, float *d_dout, float *d_dh2, void *stream);
/* rsb_replay_sample / rsb_normal keyed by DEVICE-resident counters d_ctr = {filled ring rows, update counter} (int64 x 2): no argument of an update
   changes on the host side, so sampling sits inside the captured graph of the update */
int rsb_replay_sample_dev(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, const int64_t *d_ctr, int obs_dim, int act_dim,
                          uint64_t seed, int batch, float *b_obs, int ld_obs, float *b_act, float *b_rew, float *b_term, float *b_next, int ld_next, int *b_idx, void *stream);
int rsb_normal_dev(uint64_t seed, const int64_t *d_ctr, uint32_t stream_id, int n, float *d_out, void *stream);
int rsb_counter_add(int64_t *d_counter, int64_t delta, void *stream);
/* the head of one update in ONE launch (= rsb_replay_sample_dev + rsb_normal_dev + rsb_sac_prepare): draws the batch from the ring (sample != 0; else the
   batch already sits in d_xp / b_act / b_rew / b_term), lays it out as the update's inputs Xp [2B,O], XQ [2B,O+A], XT [B,O+A], clears the loss accumulators
   and the log-alpha gradient, and draws the policy noise eps [2B, A] (noise != 0) under key (seed_noise, update counter, noise_stream).  nsums >= 8.
   In data-parallel runs it is also where the rank waits until its peers have finished reading its gradient bucket of the previous update. */
int rsb_sac_begin(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, const int64_t *d_ctr, int obs_dim, int act_dim,
                  uint64_t seed_ring, int batch, int sample, float *d_xp, int ld_xp, float *b_act, float *b_rew, float *b_term, int *b_idx, float *d_xq, float *d_xt,
                  int ld_xq /* row pitches of Xp and of XQ / XT in floats (>= O, >= O + A; multiples of 4 make them TMA-stageable GEMM operands) */, float *d_sums, int nsums, float *d_g_log_alpha, int noise, uint64_t seed_noise, uint32_t noise_stream, float *d_eps,
                  uint32_t *const *d_dp_flags, const uint32_t *d_dp_local, int dp_rank, int dp_world /* data parallel (below); world <= 1: unused */, void *stream);

/* ---- collector (csrc/rsb_collect.cu): what rlkit's MdpPathCollector.collect_new_paths does per control step and per epoch
   (util/rlkit_custom.py:202,215,223 -> rollout -> TanhGaussianPolicy.get_action; statistics: util/rlkit_custom.py:244-301,315-377) */

/* TanhGaussianPolicy.get_action for n environments in one launch (MakeDeterministic when deterministic != 0):
   obs row i = d_obs + row_i * obs_ld, action row i = d_act + row_i * act_ld with row_i = i (cap == 0) or (slot0 + i) mod cap (the replay
   ring's observations / actions arrays: rsb_step_ring then reads the action where this call put it).  Weights [in, out] row-major,
   hidden must be 256.  a = tanh(mean + exp(clamp(log_std, -20, 2)) eps); eps = Box-Muller of Philox4x32-10(key = seed,
   counter = (env_id_base + i [64 bit], 2, step * 8 + d / 4)): four action dims per Philox block, independent of the sharding. */
int rsb_policy_act(const float *d_W0, const float *d_b0, const float *d_W1, const float *d_b1, const float *d_W2, const float *d_b2,
                   int obs_dim, int act_dim, int hidden, const float *d_obs, long obs_ld, float *d_act, long act_ld, int64_t slot0, int64_t cap, int n,
                   int deterministic, uint64_t seed, uint64_t env_id_base, uint64_t step, void *stream);

/* eval_util.get_generic_path_information / get_custom_generic_path_information reduced on the device from the ring segment a collection
   round wrote (transition (env i, step t) at row (slot0 + t n + i) mod cap).  d_out holds rsb_path_stats_words(n) doubles; the first 16 are
   {sum, sum of squares, max, min} of: rewards (all T n), path returns (n), returns over the first expl_len steps (n), action entries (T n A). */
int rsb_path_stats_words(int n);
int rsb_path_stats(const float *d_rewards, const float *d_actions, int64_t slot0, int64_t cap, int n, int T, int act_dim, int expl_len, double *d_out, void *stream);

/* ---- data parallel (csrc/rsb_dp.cu): the one collective of the path -- the mean of the flat gradient bucket over the ranks, once per update
    -- fused into the optimizer kernel over NVLink peer memory.  d_peer_grads / d_peer_flags: DEVICE arrays [world] of device pointers to
   every rank's gradient bucket / flag words (>= 32 uint32, zero at start) in symmetric memory (torch.distributed._symmetric_memory: buffer_ptrs_dev);
   d_local: 4 private uint32 of this rank, zero at start.  All ranks must call these the same number of times.  Bounded waits: rsb_dp_timeouts()
   returns (and clears) the number of waits that gave up -- 0 on a healthy run; it synchronises the device. */
int rsb_dp_wait_peers_done(const float *const *d_peer_grads, uint32_t *const *d_peer_flags, uint32_t *d_local, int rank, int world, void *stream);
int rsb_adam_polyak_allreduce(const float *const *d_peer_grads, uint32_t *const *d_peer_flags, uint32_t *d_local, int rank, int world,
                              float *d_p, float *d_m, float *d_v, long n, double lr_pi, double lr_q, float b1, float b2, float eps, double *d_bc,
                              float *d_tgt, long tgt_begin, long tgt_end, float tau, int do_soft, float *d_alpha, long log_alpha_idx, void *stream);
int rsb_dp_timeouts(void);
/* diagnostic: %globaltimer [ns] of CTA 0 of the last rsb_adam_polyak_allreduce launch at 7 points (entry, dependency satisfied, READY signalled, peers ready,
   go flag set, reduction + Adam done, tail); synchronises the device */
int rsb_dp_debug_clocks(unsigned long long *host_out8);
#ifdef __cplusplus
}
#endif
#endif
