/*
 * rsb_sac_fused.cu -- the SAC update's narrow layers fused with the kernels around them (C-ABI: include/rsb_sac.h, "fused narrow layers").
 *
 * The update of rlkit's SACTrainer.train_from_torch (reference call site util/rlkit_custom.py:238) at the reference's batch of 128 is a chain
 * of dependent launches; what bounds it is the NUMBER of links.  Three of the tensor-core products are not GEMM-shaped at all
 * -- the policy's last layer (256 -> 2A <= 32 columns), the Q networks' last layer (256 -> 1: a dot product per row) and their input
 * gradients (an outer product; a 2A-term sum) -- and each sat between two small elementwise kernels.  Here each of the three groups is ONE
 * kernel on the CUDA cores, exact fp32:
 *
 *   k_policy_head_fwd : OUT = H2 W2 + b2;  a = tanh(mean + exp(clamp(log_std)) eps);  log pi          (was: product, k_head_fwd)
 *   k_q_losses        : q = H2q w2 + b, q_target = H2t w2' + b';  TD target, losses, dq;  dH2q = (dq w2^T) . [H2q > 0]
 *                                                                                                      (was: 2 products, k_losses, product)
 *   k_policy_head_bwd : dOUT from (d log pi, dQ/da);  dH2p = (dOUT W2^T) . [H2p > 0]                    (was: k_head_bwd, product)
 *
 * which takes five links out of the chain.  The weight-gradient products of these layers stay on the tensor cores (side stream).
 * Also here: replay sampling and policy noise keyed by DEVICE-resident counters, so that the whole update -- sampling included -- is one
 * CUDA-graph replay with no host-side arguments that change from update to update.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/rsb_sac.h"
#include "rsb_pdl.h"

void rsb_sac_set_error(const char *msg);
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rsb_sac_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); return 1; } } while (0)

namespace {
#define HID 256
#define LOG_SIG_MAX 2.0f
#define LOG_SIG_MIN -20.0f

__device__ __forceinline__ void philox4(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0], h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* ---------------------------------------------------------------- policy: last layer + tanh-Gaussian head
   CTA = 8 rows x 32 output columns (2A <= 32): the 8 hidden rows and W2 sit in shared memory, thread (r, c) owns out[r][c]. */
#define PH_ROWS 8
__global__ void __launch_bounds__(256) k_policy_head_fwd(const float *__restrict__ H2, const float *__restrict__ W2, const float *__restrict__ b2,
                                                         const float *__restrict__ eps, int R, int A, float *__restrict__ out, float *__restrict__ a_store,
                                                         float *__restrict__ logpi, float *__restrict__ dst0, int ld0, int row0_lo, int row0_hi,
                                                         float *__restrict__ dst1, int ld1, int row1_lo, int row1_hi) {
  __shared__ float Hs[PH_ROWS][HID + 1], Ws[HID][33], outs[PH_ROWS][33], lps[PH_ROWS][17];
  const int tid = threadIdx.x, r = tid >> 5, c = tid & 31, A2 = 2 * A, row0 = blockIdx.x * PH_ROWS;
  /* the weights were written by the PREVIOUS update's optimizer kernel, which lies at least two kernels up the stream: complete before this
     kernel could start (rsb_pdl.h: overlap is one kernel deep) -- staged while the preceding layer is still running */
  for (int i = tid; i < HID * A2; i += 256) { const int k = i / A2, cc = i - k * A2; Ws[k][cc] = W2[i]; }
  const float bias = c < A2 ? b2[c] : 0.0f;
  pdl_wait(); pdl_trigger();
#pragma unroll
  for (int rr = 0; rr < PH_ROWS; rr++) Hs[rr][tid] = (row0 + rr < R) ? H2[(size_t)(row0 + rr) * HID + tid] : 0.0f;
  __syncthreads();
  const int row = row0 + r;
  float o = 0.0f;
  if (c < A2) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
    for (int k = 0; k < HID; k += 4) { s0 = fmaf(Hs[r][k], Ws[k][c], s0); s1 = fmaf(Hs[r][k + 1], Ws[k + 1][c], s1); s2 = fmaf(Hs[r][k + 2], Ws[k + 2][c], s2); s3 = fmaf(Hs[r][k + 3], Ws[k + 3][c], s3); }
    o = (s0 + s1) + (s2 + s3) + bias;
    if (row < R) out[(size_t)row * A2 + c] = o;
  }
  outs[r][c] = o;
  __syncthreads();
  if (row < R && c < A) {
    const float mu = outs[r][c], ls = fminf(fmaxf(outs[r][A + c], LOG_SIG_MIN), LOG_SIG_MAX), e = eps[(size_t)row * A + c];
    const float a = tanhf(mu + expf(ls) * e);
    lps[r][c] = -0.5f * e * e - ls - 0.9189385332046727f - logf(1.0f - a * a + 1e-6f);
    a_store[(size_t)row * A + c] = a;
    if (dst0 && row >= row0_lo && row < row0_hi) dst0[(size_t)(row - row0_lo) * ld0 + c] = a;
    if (dst1 && row >= row1_lo && row < row1_hi) dst1[(size_t)(row - row1_lo) * ld1 + c] = a;
  }
  __syncthreads();
  if (row < R && c == 0) { float lp = 0.0f; for (int d = 0; d < A; d++) lp += lps[r][d]; logpi[row] = lp; }     /* dimension order: as k_head_fwd */
}

/* ---------------------------------------------------------------- twin Q: last layer, TD target, losses, dq, dH2q
   One warp per batch row b.  A lane holds 8 hidden units (k = 4 lane .. +3 and 128 + 4 lane .. +3) of the six hidden rows that row b needs:
   H2q[n][b] (Q_n(obs, a_new): policy loss), H2q[n][B + b] (Q_n(obs, act): Bellman error), H2t[n][b] (target nets on (next_obs, a')).
   Semantics of the losses: k_losses (rsb_sac.cu). */
#define QL_WARPS 8
__global__ void __launch_bounds__(32 * QL_WARPS) k_q_losses(const float *__restrict__ H2q, const float *__restrict__ Wq, const float *__restrict__ bq,
                                                            const float *__restrict__ H2t, const float *__restrict__ Wt, const float *__restrict__ bt,
                                                            const float *__restrict__ logpi, const float *__restrict__ rew, const float *__restrict__ term,
                                                            const float *__restrict__ alpha_logalpha, float reward_scale, float discount, float target_entropy, int B,
                                                            float *__restrict__ q, float *__restrict__ qt, float *__restrict__ dq, float *__restrict__ dH2q,
                                                            float *__restrict__ ytarget, float *__restrict__ sums, float *__restrict__ galpha) {
  __shared__ float red[QL_WARPS][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.x * QL_WARPS + w;
  pdl_wait(); pdl_trigger();
  float l1 = 0, l2 = 0, lp = 0, la = 0, mlp = 0, my = 0, ga = 0;
  if (b < B) {
    const int k0 = 4 * lane, k1 = 128 + 4 * lane;
    float4 wq[2][2], wt[2][2], h[2][2][2], ht[2][2];
    float dot[6];
    /* weights by scalar loads: views into the flat parameter buffer are only 4-byte aligned in general; activations are separate 16-byte aligned arrays */
    auto ld4 = [](const float *p) { return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)); };
    auto dot8 = [](const float4 &x0, const float4 &x1, const float4 &w0, const float4 &w1) {
      return x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w; };
#pragma unroll
    for (int n = 0; n < 2; n++) {
      wq[n][0] = ld4(Wq + n * HID + k0); wq[n][1] = ld4(Wq + n * HID + k1);
      wt[n][0] = ld4(Wt + n * HID + k0); wt[n][1] = ld4(Wt + n * HID + k1);
#pragma unroll
      for (int s = 0; s < 2; s++) {                                  /* s = 0: row b (a_new), s = 1: row B + b (act) */
        const float *hp = H2q + ((size_t)n * 2 * B + (size_t)s * B + b) * HID;
        h[n][s][0] = *reinterpret_cast<const float4 *>(hp + k0); h[n][s][1] = *reinterpret_cast<const float4 *>(hp + k1);
        dot[2 * n + s] = dot8(h[n][s][0], h[n][s][1], wq[n][0], wq[n][1]);
      }
      const float *tp = H2t + ((size_t)n * B + b) * HID;
      ht[n][0] = *reinterpret_cast<const float4 *>(tp + k0); ht[n][1] = *reinterpret_cast<const float4 *>(tp + k1);
      dot[4 + n] = dot8(ht[n][0], ht[n][1], wt[n][0], wt[n][1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 6; i++) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
    }
    const float q1n = dot[0] + bq[0], q1b = dot[1] + bq[0], q2n = dot[2] + bq[1], q2b = dot[3] + bq[1], qt1 = dot[4] + bt[0], qt2 = dot[5] + bt[1];
    const float alpha = alpha_logalpha[0], log_alpha = alpha_logalpha[1], invB = 1.0f / (float)B, lpi = logpi[b];
    const bool first = q1n <= q2n;                                   /* torch.min routes the gradient to the smaller entry */
    const float d1n = first ? -invB : 0.0f, d2n = first ? 0.0f : -invB;
    const float tq = fminf(qt1, qt2) - alpha * logpi[B + b];
    const float y = reward_scale * rew[b] + (1.0f - term[b]) * discount * tq;
    const float e1 = q1b - y, e2 = q2b - y, d1b = 2.0f * e1 * invB, d2b = 2.0f * e2 * invB;
    if (lane == 0) {
      q[b] = q1n; q[B + b] = q1b; q[2 * B + b] = q2n; q[3 * B + b] = q2b; qt[b] = qt1; qt[B + b] = qt2;
      dq[b] = d1n; dq[B + b] = d1b; dq[2 * B + b] = d2n; dq[3 * B + b] = d2b; ytarget[b] = y;
      lp = (alpha * lpi - fminf(q1n, q2n)) * invB; la = -(log_alpha * (lpi + target_entropy)) * invB; ga = -(lpi + target_entropy) * invB; mlp = lpi * invB;
      my = y * invB; l1 = e1 * e1 * invB; l2 = e2 * e2 * invB;
    }
    /* dH2q[n][row][k] = dq[n][row] * w2[n][k] where the hidden unit was active */
    const float dd[2][2] = {{d1n, d1b}, {d2n, d2b}};
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int s = 0; s < 2; s++) {
        float *gp = dH2q + ((size_t)n * 2 * B + (size_t)s * B + b) * HID; const float d = dd[n][s];
        const float4 x0 = h[n][s][0], x1 = h[n][s][1], w0 = wq[n][0], w1 = wq[n][1];
        *reinterpret_cast<float4 *>(gp + k0) = make_float4(x0.x > 0 ? d * w0.x : 0.0f, x0.y > 0 ? d * w0.y : 0.0f, x0.z > 0 ? d * w0.z : 0.0f, x0.w > 0 ? d * w0.w : 0.0f);
        *reinterpret_cast<float4 *>(gp + k1) = make_float4(x1.x > 0 ? d * w1.x : 0.0f, x1.y > 0 ? d * w1.y : 0.0f, x1.z > 0 ? d * w1.z : 0.0f, x1.w > 0 ? d * w1.w : 0.0f);
      }
  }
  if (lane == 0) { red[w][0] = l1; red[w][1] = l2; red[w][2] = lp; red[w][3] = la; red[w][4] = mlp; red[w][5] = my; red[w][6] = ga; }
  __syncthreads();
  if (threadIdx.x < 7) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < QL_WARPS; i++) t += red[i][threadIdx.x];
    atomicAdd(threadIdx.x < 6 ? sums + threadIdx.x : galpha, t);
  }
}

/* ---------------------------------------------------------------- policy: head backward + last layer's input gradient
   CTA = one batch row r < B.  Threads d < A: (d mean, d log_std) as k_head_bwd; then thread j: dH2[r][j] = sum_c dOUT[r][c] W2[j][c] where unit j was active. */
__global__ void __launch_bounds__(HID) k_policy_head_bwd(const float *__restrict__ out, const float *__restrict__ eps, const float *__restrict__ a_store,
                                                         const float *__restrict__ H2, const float *__restrict__ W2, int B, int A, const float *__restrict__ alpha,
                                                         const float *__restrict__ g_a, int ld_ga, float *__restrict__ d_out, float *__restrict__ dH2) {
  __shared__ float ds[32];
  const int r = blockIdx.x, j = threadIdx.x, A2 = 2 * A;
  pdl_wait(); pdl_trigger();
  if (j < A) {
    const float raw = out[(size_t)r * A2 + A + j], ls = fminf(fmaxf(raw, LOG_SIG_MIN), LOG_SIG_MAX), e = eps[(size_t)r * A + j], a = a_store[(size_t)r * A + j];
    const float g_lp = alpha[0] / (float)B, one = 1.0f - a * a;
    const float dz = g_a[(size_t)r * ld_ga + j] * one + g_lp * 2.0f * a * one / (one + 1e-6f);
    float dls = dz * expf(ls) * e - g_lp;
    if (raw < LOG_SIG_MIN || raw > LOG_SIG_MAX) dls = 0;
    ds[j] = dz; ds[A + j] = dls;
    d_out[(size_t)r * A2 + j] = dz; d_out[(size_t)r * A2 + A + j] = dls;
  }
  __syncthreads();
  float s = 0.0f;
  const float *wrow = W2 + (size_t)j * A2;
  for (int c = 0; c < A2; c++) s = fmaf(ds[c], wrow[c], s);
  dH2[(size_t)r * HID + j] = H2[(size_t)r * HID + j] > 0.0f ? s : 0.0f;
}

/* ---------------------------------------------------------------- device-resident counters: ctr[0] = filled ring rows, ctr[1] = update counter */
__global__ void k_replay_sample_dev(const float *__restrict__ obs, const float *__restrict__ act, const float *__restrict__ rew,
                                    const unsigned char *__restrict__ term, const float *__restrict__ next_obs, const long long *__restrict__ ctr,
                                    int O, int A, uint64_t seed, int B, float *__restrict__ b_obs, float *__restrict__ b_act, float *__restrict__ b_rew,
                                    float *__restrict__ b_term, float *__restrict__ b_next, int *__restrict__ b_idx, int ld_obs, int ld_next) {
  pdl_wait(); pdl_trigger();
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  const uint32_t size = (uint32_t)ctr[0]; const uint64_t step = (uint64_t)ctr[1];
  uint32_t c[4] = {(uint32_t)row, (uint32_t)step, (uint32_t)(step >> 32), 0xB0FFE7u};
  philox4(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  int idx = (int)__umulhi(c[0], size);
  const float *so = obs + (size_t)idx * O, *sn = next_obs + (size_t)idx * O, *sa = act + (size_t)idx * A;
  for (int k = lane; k < O; k += 32) { b_obs[(size_t)row * ld_obs + k] = so[k]; b_next[(size_t)row * ld_next + k] = sn[k]; }
  for (int k = lane; k < A; k += 32) b_act[(size_t)row * A + k] = sa[k];
  if (lane == 0) { b_rew[row] = rew[idx]; b_term[row] = (float)term[idx]; if (b_idx) b_idx[row] = idx; }
}
__global__ void k_normal_dev(uint64_t seed, const long long *__restrict__ ctr, uint32_t stream, int n, float *__restrict__ out) {
  pdl_wait(); pdl_trigger();
  int i = blockIdx.x * blockDim.x + threadIdx.x;               /* one thread -> 4 values; same arithmetic as k_normal */
  if (4 * i >= n) return;
  const uint64_t step = (uint64_t)ctr[1];
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
__global__ void k_counter_add(long long *__restrict__ p, long long delta) { pdl_wait(); pdl_trigger(); p[0] += delta; }

/* ---------------------------------------------------------------- the head of an update in ONE kernel
   (was: k_replay_sample_dev, k_normal_dev, k_sac_prepare -- three links of the chain).  Warp b handles batch row b:
     sample != 0: ring row idx_b = mulhi(Philox(seed_ring; b, update counter), filled rows) -> obs, next_obs, act, rew, term;  sample == 0: the row
                  already sits in Xp / act / rew / term (explicit batch);
     either way the row is laid out for the update: Xp[b] = obs, Xp[B + b] = next_obs, XQ[b][:O] = XQ[B + b][:O] = obs, XQ[B + b][O:] = act, XT[b][:O] = next_obs;
   thread gid < ceil(2 B A / 4) also draws four N(0,1) of the policy noise (noise != 0), thread gid < nsums clears a loss accumulator. */
__global__ void k_sac_begin(const float *__restrict__ obs, const float *__restrict__ act_ring, const float *__restrict__ rew_ring, const unsigned char *__restrict__ term_ring,
                            const float *__restrict__ next_obs, const long long *__restrict__ ctr, int O, int A, uint64_t seed_ring, int B, int sample,
                            float *__restrict__ Xp, int ldp, float *__restrict__ act, float *__restrict__ rew, float *__restrict__ term, int *__restrict__ b_idx,
                            float *__restrict__ XQ, float *__restrict__ XT, int ldq, float *__restrict__ sums, int nsums, float *__restrict__ g_log_alpha,
                            int noise, uint64_t seed_noise, uint32_t noise_stream, float *__restrict__ eps,
                            uint32_t *const *__restrict__ dp_flags, const uint32_t *__restrict__ dp_local, int dp_rank, int dp_world) {
  pdl_wait(); pdl_trigger();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, b = gid >> 5, lane = gid & 31;
  const uint64_t step = (uint64_t)ctr[1];
  if (gid < nsums) sums[gid] = 0.0f;
  /* data parallel (csrc/rsb_dp.cu): the gradient bucket -- g_log_alpha is its last word -- may be overwritten only after every peer has finished reading it
     for the previous update; the store below is the first write of the new update, every other one comes in later kernels.  Flag word [16 + r] of THIS
     rank's flags holds the last epoch rank r has finished reading; thread r of CTA 0 polls it (bounded), all in parallel. */
  if (blockIdx.x == 0 && dp_world > 1) {
    if ((int)threadIdx.x < dp_world) {
      const uint32_t epoch = dp_local[0]; const uint32_t *f = dp_flags[dp_rank] + 16 + threadIdx.x; bool ok = false;
      for (int it = 0; it < (1 << 24) && !ok; it++) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory"); ok = (int32_t)(v - epoch) >= 0; }
      if (!ok) sums[7] = __int_as_float(0x7fc00000);               /* poisoned statistics word: a missing peer is visible in the log (and rsb_dp_timeouts of the optimizer kernel fires) */
    }
    __syncthreads();
  }
  if (gid == 0) g_log_alpha[0] = 0.0f;
  if (noise && 4 * gid < 2 * B * A) {                                   /* same arithmetic as k_normal */
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)step, (uint32_t)(step >> 32), noise_stream};
    philox4(c, (uint32_t)seed_noise, (uint32_t)(seed_noise >> 32));
    float z[4];
#pragma unroll
    for (int p = 0; p < 2; p++) {
      float u1 = ((float)(c[2 * p] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(c[2 * p + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      float rad = sqrtf(-2.0f * logf(u1)), sn, cs; sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * p] = rad * cs; z[2 * p + 1] = rad * sn;
    }
    for (int k = 0; k < 4; k++) if (4 * gid + k < 2 * B * A) eps[4 * gid + k] = z[k];
  }
  if (b >= B) return;
  const float *so, *sn_, *sa;
  if (sample) {
    uint32_t c[4] = {(uint32_t)b, (uint32_t)step, (uint32_t)(step >> 32), 0xB0FFE7u};
    philox4(c, (uint32_t)seed_ring, (uint32_t)(seed_ring >> 32));
    const int idx = (int)__umulhi(c[0], (uint32_t)ctr[0]);
    so = obs + (size_t)idx * O; sn_ = next_obs + (size_t)idx * O; sa = act_ring + (size_t)idx * A;
    if (lane == 0) { rew[b] = rew_ring[idx]; term[b] = (float)term_ring[idx]; if (b_idx) b_idx[b] = idx; }
  } else { so = Xp + (size_t)b * ldp; sn_ = Xp + (size_t)(B + b) * ldp; sa = act + (size_t)b * A; }
  for (int k = lane; k < O; k += 32) {                                  /* ldp / ldq: row pitches of Xp and of XQ / XT (padded to 16 bytes: TMA-stageable) */
    const float v = so[k], w = sn_[k];
    if (sample) { Xp[(size_t)b * ldp + k] = v; Xp[(size_t)(B + b) * ldp + k] = w; }
    XQ[(size_t)b * ldq + k] = v; XQ[(size_t)(B + b) * ldq + k] = v; XT[(size_t)b * ldq + k] = w;
  }
  for (int k = lane; k < A; k += 32) { const float v = sa[k]; if (sample) act[(size_t)b * A + k] = v; XQ[(size_t)(B + b) * ldq + O + k] = v; }
}
}  // namespace

extern "C" {

int rsb_policy_head_fwd(const float *d_h2, const float *d_w2, const float *d_b2, const float *d_eps, int rows, int act_dim, float *d_out, float *d_a, float *d_logpi,
                        float *dst0, int ld0, int r0lo, int r0hi, float *dst1, int ld1, int r1lo, int r1hi, void *stream) {
  if (act_dim < 1 || act_dim > 16 || rows < 1) { rsb_sac_set_error("policy_head_fwd: act_dim in [1, 16] required"); return 2; }
  CKF(rsb_launch_pdl(k_policy_head_fwd, dim3((rows + PH_ROWS - 1) / PH_ROWS), dim3(256), 0, (cudaStream_t)stream, 1, d_h2, d_w2, d_b2, d_eps, rows, act_dim, d_out, d_a, d_logpi,
                     dst0, ld0, r0lo, r0hi, dst1, ld1, r1lo, r1hi));
  return 0;
}
int rsb_q_losses(const float *d_h2q, const float *d_wq2, const float *d_bq2, const float *d_h2t, const float *d_wt2, const float *d_bt2, const float *d_logpi,
                 const float *d_rew, const float *d_term, const float *d_alpha, float reward_scale, float discount, float target_entropy, int batch,
                 float *d_q, float *d_qt, float *d_dq, float *d_dh2q, float *d_y, float *d_sums, float *d_galpha, void *stream) {
  if (batch < 1) { rsb_sac_set_error("q_losses: empty batch"); return 2; }
  CKF(rsb_launch_pdl(k_q_losses, dim3((batch + QL_WARPS - 1) / QL_WARPS), dim3(32 * QL_WARPS), 0, (cudaStream_t)stream, 1, d_h2q, d_wq2, d_bq2, d_h2t, d_wt2, d_bt2, d_logpi, d_rew, d_term,
                     d_alpha, reward_scale, discount, target_entropy, batch, d_q, d_qt, d_dq, d_dh2q, d_y, d_sums, d_galpha));
  return 0;
}
int rsb_policy_head_bwd(const float *d_out, const float *d_eps, const float *d_a, const float *d_h2, const float *d_w2, int batch, int act_dim, const float *d_alpha,
                        const float *d_ga, int ld_ga, float *d_dout, float *d_dh2, void *stream) {
  if (act_dim < 1 || act_dim > 16 || batch < 1) { rsb_sac_set_error("policy_head_bwd: act_dim in [1, 16] required"); return 2; }
  CKF(rsb_launch_pdl(k_policy_head_bwd, dim3(batch), dim3(HID), 0, (cudaStream_t)stream, 1, d_out, d_eps, d_a, d_h2, d_w2, batch, act_dim, d_alpha, d_ga, ld_ga, d_dout, d_dh2));
  return 0;
}
int rsb_replay_sample_dev(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, const int64_t *d_ctr, int obs_dim, int act_dim,
                          uint64_t seed, int batch, float *b_obs, int ld_obs, float *b_act, float *b_rew, float *b_term, float *b_next, int ld_next, int *b_idx, void *stream) {
  if (batch <= 0 || !d_ctr) { rsb_sac_set_error("replay_sample_dev: bad arguments"); return 2; }
  int threads = 128, blocks = (batch * 32 + threads - 1) / threads;
  CKF(rsb_launch_pdl(k_replay_sample_dev, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, 1, d_obs, d_act, d_rew, d_term, d_next, (const long long *)d_ctr, obs_dim, act_dim, seed, batch,
                     b_obs, b_act, b_rew, b_term, b_next, b_idx, ld_obs, ld_next));
  return 0;
}
int rsb_normal_dev(uint64_t seed, const int64_t *d_ctr, uint32_t stream_id, int n, float *d_out, void *stream) {
  int t = (n + 3) / 4; CKF(rsb_launch_pdl(k_normal_dev, dim3((t + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, seed, (const long long *)d_ctr, stream_id, n, d_out)); return 0;
}
int rsb_sac_begin(const float *d_obs, const float *d_act, const float *d_rew, const uint8_t *d_term, const float *d_next, const int64_t *d_ctr, int obs_dim, int act_dim,
                  uint64_t seed_ring, int batch, int sample, float *d_xp, int ld_xp, float *b_act, float *b_rew, float *b_term, int *b_idx, float *d_xq, float *d_xt,
                  int ld_xq, float *d_sums, int nsums, float *d_g_log_alpha, int noise, uint64_t seed_noise, uint32_t noise_stream, float *d_eps,
                  uint32_t *const *d_dp_flags, const uint32_t *d_dp_local, int dp_rank, int dp_world, void *stream) {
  if (batch <= 0 || !d_ctr || nsums > 32 * batch || nsums < 8 || ld_xp < obs_dim || ld_xq < obs_dim + act_dim) { rsb_sac_set_error("sac_begin: bad arguments"); return 2; }
  if (dp_world > 1 && (!d_dp_flags || !d_dp_local || dp_rank < 0 || dp_rank >= dp_world)) { rsb_sac_set_error("sac_begin: bad data-parallel arguments"); return 2; }
  if (sample && (!d_obs || !d_act || !d_rew || !d_term || !d_next)) { rsb_sac_set_error("sac_begin: sampling needs the replay ring"); return 2; }
  const int threads = 128, blocks = (batch * 32 + threads - 1) / threads;       /* 32 B threads >= ceil(2 B A / 4) for A <= 64 */
  CKF(rsb_launch_pdl(k_sac_begin, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, 1, d_obs, d_act, d_rew, d_term, d_next, (const long long *)d_ctr, obs_dim, act_dim, seed_ring,
                     batch, sample, d_xp, ld_xp, b_act, b_rew, b_term, b_idx, d_xq, d_xt, ld_xq, d_sums, nsums, d_g_log_alpha, noise, seed_noise, noise_stream, d_eps,
                     d_dp_flags, d_dp_local, dp_rank, dp_world));
  return 0;
}
int rsb_counter_add(int64_t *d_counter, int64_t delta, void *stream) {
  CKF(rsb_launch_pdl(k_counter_add, dim3(1), dim3(1), 0, (cudaStream_t)stream, 1, (long long *)d_counter, (long long)delta)); return 0;
}

}  /* extern "C" */
