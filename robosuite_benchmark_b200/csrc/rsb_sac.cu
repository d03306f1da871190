/*
 * rsb_sac.cu -- replay sampling/gather and the non-GEMM parts of the rlkit SAC update as sm_100a kernels (C-ABI: include/rsb_sac.h).
 *
 * Stands behind rlkit's EnvReplayBuffer.random_batch and SACTrainer.train_from_torch (reference call sites
 * util/rlkit_custom.py:235-238).  The dense products of the five 256x256 MLPs run in rsb_tc_gemm.cu (tcgen05 TF32,
 * FP32 accumulate in tensor memory, bias / ReLU / ReLU-backward in its epilogue); everything between them is fused here: Philox index draw + row gather,
 * tanh-Gaussian sampling / log-prob and its backward, TD targets and loss gradients, bias+ReLU, and ONE Adam + Polyak
 * kernel over the flat parameter buffer.  All calls are stream-ordered and capturable in a CUDA graph (no syncs, no allocs).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/rsb_sac.h"
#include "rsb_pdl.h"

static thread_local std::string g_sac_err;
void rsb_sac_set_error(const char *msg) { g_sac_err = msg; }   /* shared with rsb_tc_gemm.cu */
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_sac_err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)

__device__ __forceinline__ void philox4(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0], h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* ---------------------------------------------------------------- replay: Philox index draw (with replacement) + row gather
   index of batch row b at update `step`: Philox(key = seed; counter = (b, step_lo, step_hi, 0xB0FFE7)) -> word0; idx = mulhi(word0, size).
   One warp per row; obs/next_obs/actions rows are copied with coalesced 32-lane loads. */
__global__ void k_replay_sample(const float *__restrict__ obs, const float *__restrict__ act, const float *__restrict__ rew,
                                const unsigned char *__restrict__ term, const float *__restrict__ next_obs,
                                int size, int O, int A, uint64_t seed, uint64_t step, int B,
                                float *__restrict__ b_obs, float *__restrict__ b_act, float *__restrict__ b_rew, float *__restrict__ b_term,
                                float *__restrict__ b_next, int *__restrict__ b_idx, int ld_obs, int ld_next) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  uint32_t c[4] = {(uint32_t)row, (uint32_t)step, (uint32_t)(step >> 32), 0xB0FFE7u};
  philox4(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  int idx = (int)__umulhi(c[0], (uint32_t)size);
  const float *so = obs + (size_t)idx * O, *sn = next_obs + (size_t)idx * O, *sa = act + (size_t)idx * A;
  for (int k = lane; k < O; k += 32) { b_obs[(size_t)row * ld_obs + k] = so[k]; b_next[(size_t)row * ld_next + k] = sn[k]; }
  for (int k = lane; k < A; k += 32) b_act[(size_t)row * A + k] = sa[k];
  if (lane == 0) { b_rew[row] = rew[idx]; b_term[row] = (float)term[idx]; if (b_idx) b_idx[row] = idx; }
}

/* ---------------------------------------------------------------- N(0,1) noise, Philox keyed (seed, step, stream)  */
__global__ void k_normal(uint64_t seed, uint64_t step, uint32_t stream, int n, float *__restrict__ out) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  int i = blockIdx.x * blockDim.x + threadIdx.x;               /* one thread -> 4 values */
  if (4 * i >= n) return;
  uint32_t c[4] = {(uint32_t)i, (uint32_t)step, (uint32_t)(step >> 32), stream};
  philox4(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float z[4];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    float u1 = ((float)(c[2 * p] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(c[2 * p + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1)), sn, cs; sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * p] = rad * cs; z[2 * p + 1] = rad * sn;
  }
  for (int k = 0; k < 4; k++) if (4 * i + k < n) out[4 * i + k] = z[k];
}

/* ---------------------------------------------------------------- bias + ReLU (in place), row-major [rows, cols] */
__global__ void k_bias_relu(float *__restrict__ x, const float *__restrict__ bias, int rows, int cols, int relu, int nmat, long mat_stride, int bias_stride) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x; long per = (long)rows * cols;
  if (i >= per * nmat) return;
  int mt = (int)(i / per); long r = i - (long)mt * per; int c = (int)(r % cols);
  float v = x[(long)mt * mat_stride + r] + bias[mt * bias_stride + c];
  x[(long)mt * mat_stride + r] = relu ? fmaxf(v, 0.0f) : v;
}
/* dZ = dY * (Y > 0)  (in place on dY) */
__global__ void k_relu_bwd(float *__restrict__ dy, const float *__restrict__ y, long n) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dy[i] = y[i] > 0.0f ? dy[i] : 0.0f;
}
/* column sums of dY [rows, cols] rows in [r0, r1) -> db[cols].  Block = 32 columns x 32 row slices: thread (x, y) adds rows r0+y, r0+y+32, ...
   (independent coalesced loads, 4 in flight), the 32 slice sums of a column are added in slice order through shared memory (deterministic).
   A thread per column walking all rows is a chain of (r1-r0) dependent-latency loads: 10-16 us at 128 rows, on the tail of every update. */
__global__ void __launch_bounds__(1024) k_colsum(const float *__restrict__ dy, int r0, int r1, int cols, float *__restrict__ db, int nmat, long mat_stride, int db_stride) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  __shared__ float s[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, mt = blockIdx.y;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  if (c < cols) {
    const float *p = dy + (long)mt * mat_stride + c;
    int r = r0 + threadIdx.y;
    for (; r + 96 < r1; r += 128) {
      a0 += p[(long)r * cols]; a1 += p[(long)(r + 32) * cols]; a2 += p[(long)(r + 64) * cols]; a3 += p[(long)(r + 96) * cols];
    }
    for (; r < r1; r += 32) a0 += p[(long)r * cols];
  }
  s[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int y = 0; y < 32; y++) t += s[y][threadIdx.x];
    db[mt * db_stride + c] = t;
  }
}

/* ---------------------------------------------------------------- tanh-Gaussian head (rlkit TanhGaussianPolicy / TanhNormal)
   out [R, 2A] = (mean | raw log_std); eps [R, A].  a = tanh(mean + exp(clamp(log_std,-20,2)) eps),
   logpi = sum_d [ -0.5 eps^2 - log_std - 0.5 log(2 pi) - log(1 - a^2 + 1e-6) ].
   `a` is written to up to two destinations with leading dims (the action columns of the Q-network input buffers). */
#define LOG_SIG_MAX 2.0f
#define LOG_SIG_MIN -20.0f
__global__ void k_head_fwd(const float *__restrict__ out, const float *__restrict__ eps, int R, int A,
                           float *__restrict__ a_store, float *__restrict__ logpi,
                           float *__restrict__ dst0, int ld0, int row0_lo, int row0_hi, float *__restrict__ dst1, int ld1, int row1_lo, int row1_hi) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= R) return;
  float lp = 0;
  for (int d = 0; d < A; d++) {
    float mu = out[(size_t)r * 2 * A + d], ls = fminf(fmaxf(out[(size_t)r * 2 * A + A + d], LOG_SIG_MIN), LOG_SIG_MAX), e = eps[(size_t)r * A + d];
    float a = tanhf(mu + expf(ls) * e);
    lp += -0.5f * e * e - ls - 0.9189385332046727f - logf(1.0f - a * a + 1e-6f);
    a_store[(size_t)r * A + d] = a;
    if (dst0 && r >= row0_lo && r < row0_hi) dst0[(size_t)(r - row0_lo) * ld0 + d] = a;
    if (dst1 && r >= row1_lo && r < row1_hi) dst1[(size_t)(r - row1_lo) * ld1 + d] = a;
  }
  logpi[r] = lp;
}
/* backward for rows [0, B): upstream g_lp (scalar, same for every row: alpha / B) and g_a [B, A] with leading dim ld_ga
   (the action columns of d(Q input)).  d_out [R, 2A] is fully written (rows >= B get zero: next_obs rows carry no gradient). */
__global__ void k_head_bwd(const float *__restrict__ out, const float *__restrict__ eps, const float *__restrict__ a_store, int R, int B, int A,
                           const float *__restrict__ alpha, float inv_B, const float *__restrict__ g_a, int ld_ga, float *__restrict__ d_out) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= R * A) return;
  int r = i / A, d = i - r * A; float dmu = 0, dls = 0;
  if (r < B) {
    float raw = out[(size_t)r * 2 * A + A + d], ls = fminf(fmaxf(raw, LOG_SIG_MIN), LOG_SIG_MAX), e = eps[i], a = a_store[i];
    float g_lp = alpha[0] * inv_B, one = 1.0f - a * a;
    float dz = g_a[(size_t)r * ld_ga + d] * one + g_lp * 2.0f * a * one / (one + 1e-6f);
    dmu = dz; dls = dz * expf(ls) * e - g_lp;
    if (raw < LOG_SIG_MIN || raw > LOG_SIG_MAX) dls = 0;
  }
  d_out[(size_t)r * 2 * A + d] = dmu; d_out[(size_t)r * 2 * A + A + d] = dls;
}

/* ---------------------------------------------------------------- losses: TD target, dQ, alpha gradient, loss sums
   q [2, 2B]: rows [0,B) = Q_i(obs, a_new), rows [B,2B) = Q_i(obs, act); qt [2, B] = target nets on (next_obs, a').
   dq [2, 2B]: rows [0,B): d policy_loss / dq_i = -1/B on the arg-min network (0 on the other); rows [B,2B): 2 (q_i - y) / B.
   sums[0..5] += qf1_loss, qf2_loss, policy_loss, alpha_loss, mean(logpi), mean(y)   (pre-zeroed by the caller; atomics)
   galpha[0] = d alpha_loss / d log_alpha = -mean(logpi + target_entropy). */
__global__ void k_losses(const float *__restrict__ q, const float *__restrict__ qt, const float *__restrict__ logpi /*[2B]*/,
                         const float *__restrict__ rew, const float *__restrict__ term, const float *__restrict__ alpha_logalpha /* [alpha, log_alpha] */,
                         float reward_scale, float discount, float target_entropy, int B, float *__restrict__ dq, float *__restrict__ ytarget,
                         float *__restrict__ sums, float *__restrict__ galpha) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float l1 = 0, l2 = 0, lp = 0, la = 0, mlp = 0, my = 0, ga = 0;
  if (b < B) {
    float alpha = alpha_logalpha[0], log_alpha = alpha_logalpha[1], invB = 1.0f / (float)B;
    float q1n = q[b], q2n = q[2 * B + b], lpi = logpi[b];
    bool first = q1n <= q2n;                                        /* torch.min routes the gradient to the smaller entry */
    dq[b] = first ? -invB : 0.0f; dq[2 * B + b] = first ? 0.0f : -invB;
    lp = (alpha * lpi - fminf(q1n, q2n)) * invB;
    la = -(log_alpha * (lpi + target_entropy)) * invB; ga = -(lpi + target_entropy) * invB; mlp = lpi * invB;
    float tq = fminf(qt[b], qt[B + b]) - alpha * logpi[B + b];
    float y = reward_scale * rew[b] + (1.0f - term[b]) * discount * tq; ytarget[b] = y; my = y * invB;
    float e1 = q[B + b] - y, e2 = q[3 * B + b] - y;
    dq[B + b] = 2.0f * e1 * invB; dq[3 * B + b] = 2.0f * e2 * invB; l1 = e1 * e1 * invB; l2 = e2 * e2 * invB;
  }
  /* warp reduce, then one atomic per warp */
  for (int o = 16; o > 0; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o); l2 += __shfl_xor_sync(0xffffffffu, l2, o); lp += __shfl_xor_sync(0xffffffffu, lp, o);
    la += __shfl_xor_sync(0xffffffffu, la, o); mlp += __shfl_xor_sync(0xffffffffu, mlp, o); my += __shfl_xor_sync(0xffffffffu, my, o); ga += __shfl_xor_sync(0xffffffffu, ga, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + 0, l1); atomicAdd(sums + 1, l2); atomicAdd(sums + 2, lp); atomicAdd(sums + 3, la); atomicAdd(sums + 4, mlp); atomicAdd(sums + 5, my);
    atomicAdd(galpha, ga);
  }
}

/* ---------------------------------------------------------------- Adam (torch.optim.Adam semantics) + Polyak, one launch over the flat buffer
   params/grads/m/v are flat fp32; the Q segment [tgt_begin, tgt_end) uses lr_q and has a target copy `tgt`, the rest uses lr_pi; the soft update uses the UPDATED parameter, as rlkit does (step, then soft_update). */
__global__ void k_adam_polyak(float *__restrict__ p, const float *__restrict__ gr, float *__restrict__ m, float *__restrict__ v, long n,
                              double lr_pi, double lr_q, float b1, float b2, float eps, const double *__restrict__ bc /* [1-b1^t, sqrt(1-b2^t), b1^t, b2^t] */,
                              float *__restrict__ tgt, long tgt_begin, long tgt_end, float tau, int do_soft,
                              float *__restrict__ alpha_out /* [alpha, log_alpha] refreshed from p[log_alpha_idx] */, long log_alpha_idx) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const bool isq = i >= tgt_begin && i < tgt_end;
  /* torch.optim.Adam keeps step_size = lr / bias_correction1 and sqrt(bias_correction2) as Python doubles, cast to fp32 at use */
  const float step_size = (float)((isq ? lr_q : lr_pi) / bc[0]), bc2s = (float)bc[1];
  float g = gr[i], mi = b1 * m[i] + (1.0f - b1) * g, vi = b2 * v[i] + (1.0f - b2) * g * g; m[i] = mi; v[i] = vi;
  float denom = sqrtf(vi) / bc2s + eps, pi = p[i] - step_size * (mi / denom); p[i] = pi;
  if (do_soft && isq) { long k = i - tgt_begin; tgt[k] = (1.0f - tau) * tgt[k] + tau * pi; }
  if (i == log_alpha_idx) { alpha_out[1] = pi; alpha_out[0] = (float)exp((double)pi); }      /* correctly rounded, like torch's CPU exp */
}
/* bias corrections advance on the device so the whole update is graph-replayable */
__global__ void k_adam_tick(double *__restrict__ bc, double b1, double b2, long long *__restrict__ step_counter) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  double p1 = bc[2] * b1, p2 = bc[3] * b2; bc[2] = p1; bc[3] = p2; bc[0] = 1.0 - p1; bc[1] = sqrt(1.0 - p2);
  if (step_counter) step_counter[0] += 1;    /* the update counter that keys the NEXT update's replay indices and policy noise (rsb_replay_sample_dev) */
}

/* Inputs of the update from the sampled batch, in one launch: XQ[2B, O+A] = [(obs, .) ; (obs, act)] (the first B action slots are filled by
   the policy head later), XT[B, O+A] = (next_obs, .), and the loss accumulators cleared.  Xp[2B, O] = [obs ; next_obs] as sampled. */
__global__ void k_sac_prepare(const float *__restrict__ Xp, const float *__restrict__ act, float *__restrict__ XQ, float *__restrict__ XT,
                              float *__restrict__ sums, int nsums, float *__restrict__ g_log_alpha, int B, int O, int A) {
  pdl_wait(); pdl_trigger();                /* rsb_pdl.h: everything above is independent of earlier kernels */
  const int QI = O + A, i = blockIdx.x * blockDim.x + threadIdx.x, n = 2 * B * QI;
  if (i < nsums) sums[i] = 0.0f;
  if (i == 0) g_log_alpha[0] = 0.0f;
  if (i >= n) return;
  const int r = i / QI, c = i - r * QI;                          /* row of XQ */
  const int b = r < B ? r : r - B;
  if (c < O) { const float v = Xp[(size_t)b * O + c]; XQ[i] = v; if (r < B) XT[(size_t)b * QI + c] = Xp[(size_t)(B + b) * O + c]; }
  else if (r >= B) XQ[i] = act[(size_t)b * A + (c - O)];
}

extern "C" {
const char *rsb_sac_last_error(void) { return g_sac_err.c_str(); }
int rsb_sac_prepare(const float *d_xp, const float *d_act, float *d_xq, float *d_xt, float *d_sums, int nsums, float *d_g_log_alpha, int batch, int obs_dim, int act_dim, void *stream) {
  int n = 2 * batch * (obs_dim + act_dim); if (n < nsums) n = nsums;
  CKS(rsb_launch_pdl(k_sac_prepare, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, 1, d_xp, d_act, d_xq, d_xt, d_sums, nsums, d_g_log_alpha, batch, obs_dim, act_dim)); return 0;
}

int rsb_replay_sample(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, int size, int obs_dim, int act_dim,
                      uint64_t seed, uint64_t step, int batch, float *b_obs, int ld_obs, float *b_act, float *b_rew, float *b_term, float *b_next, int ld_next, int *b_idx, void *stream) {
  if (size <= 0 || batch <= 0) { g_sac_err = "replay_sample: empty buffer or batch"; return 2; }
  int threads = 128, blocks = (batch * 32 + threads - 1) / threads;
  CKS(rsb_launch_pdl(k_replay_sample, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, 1, d_obs, d_act, d_rew, d_term, d_next, size, obs_dim, act_dim, seed, step, batch, b_obs, b_act, b_rew, b_term, b_next, b_idx, ld_obs, ld_next));
  return 0;
}
int rsb_normal(uint64_t seed, uint64_t step, uint32_t stream_id, int n, float *d_out, void *stream) {
  int t = (n + 3) / 4; CKS(rsb_launch_pdl(k_normal, dim3((t + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, seed, step, stream_id, n, d_out)); return 0;
}
int rsb_bias_relu(float *d_x, const float *d_bias, int rows, int cols, int relu, int nmat, long mat_stride, int bias_stride, void *stream) {
  long n = (long)rows * cols * nmat; CKS(rsb_launch_pdl(k_bias_relu, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1, d_x, d_bias, rows, cols, relu, nmat, mat_stride, bias_stride)); return 0;
}
int rsb_relu_bwd(float *d_dy, const float *d_y, long n, void *stream) { CKS(rsb_launch_pdl(k_relu_bwd, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1, d_dy, d_y, n)); return 0; }
int rsb_colsum(const float *d_dy, int r0, int r1, int cols, float *d_db, int nmat, long mat_stride, int db_stride, void *stream) {
  dim3 grid((cols + 31) / 32, nmat); CKS(rsb_launch_pdl(k_colsum, grid, dim3(32, 32), 0, (cudaStream_t)stream, 1, d_dy, r0, r1, cols, d_db, nmat, mat_stride, db_stride)); return 0;
}
int rsb_head_fwd(const float *d_out, const float *d_eps, int rows, int act_dim, float *d_a, float *d_logpi, float *dst0, int ld0, int r0lo, int r0hi, float *dst1, int ld1, int r1lo, int r1hi, void *stream) {
  CKS(rsb_launch_pdl(k_head_fwd, dim3((rows + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, d_out, d_eps, rows, act_dim, d_a, d_logpi, dst0, ld0, r0lo, r0hi, dst1, ld1, r1lo, r1hi)); return 0;
}
int rsb_head_bwd(const float *d_out, const float *d_eps, const float *d_a, int rows, int batch, int act_dim, const float *d_alpha, const float *d_ga, int ld_ga, float *d_dout, void *stream) {
  int n = rows * act_dim; CKS(rsb_launch_pdl(k_head_bwd, dim3((n + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, d_out, d_eps, d_a, rows, batch, act_dim, d_alpha, 1.0f / (float)batch, d_ga, ld_ga, d_dout)); return 0;
}
int rsb_sac_losses(const float *d_q, const float *d_qt, const float *d_logpi, const float *d_rew, const float *d_term, const float *d_alpha, float reward_scale, float discount,
                   float target_entropy, int batch, float *d_dq, float *d_y, float *d_sums, float *d_galpha, void *stream) {
  CKS(rsb_launch_pdl(k_losses, dim3((batch + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, d_q, d_qt, d_logpi, d_rew, d_term, d_alpha, reward_scale, discount, target_entropy, batch, d_dq, d_y, d_sums, d_galpha)); return 0;
}
int rsb_adam_tick(double *d_bc, float b1, float b2, int64_t *d_step_counter, void *stream) {
  CKS(rsb_launch_pdl(k_adam_tick, dim3(1), dim3(1), 0, (cudaStream_t)stream, 1, d_bc, (double)b1, (double)b2, (long long *)d_step_counter)); return 0;
}
int rsb_adam_polyak(float *d_p, const float *d_g, float *d_m, float *d_v, long n, double lr_pi, double lr_q, float b1, float b2, float eps, double *d_bc,
                    float *d_tgt, long tgt_begin, long tgt_end, float tau, int do_soft, float *d_alpha, long log_alpha_idx, int tick, void *stream) {
  if (tick) CKS(rsb_launch_pdl(k_adam_tick, dim3(1), dim3(1), 0, (cudaStream_t)stream, 1, d_bc, (double)b1, (double)b2, (long long *)nullptr));
  CKS(rsb_launch_pdl(k_adam_polyak, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1, d_p, d_g, d_m, d_v, n, lr_pi, lr_q, b1, b2, eps, (const double *)d_bc, d_tgt, tgt_begin, tgt_end, tau, do_soft, d_alpha, log_alpha_idx));
  return 0;
}
}
