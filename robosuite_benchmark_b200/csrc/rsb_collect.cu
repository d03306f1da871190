/*
 * rsb_collect.cu -- the collector side of the hot path as sm_100a kernels (C-ABI: include/rsb_sac.h, "collector" section).
 *
 * Stands behind rlkit's MdpPathCollector.collect_new_paths -> rollout -> TanhGaussianPolicy.get_action (the per-step policy forward,
 * reference call sites util/rlkit_custom.py:202,215,223) and behind the per-epoch path statistics of
 * eval_util.get_generic_path_information / get_custom_generic_path_information (util/rlkit_custom.py:244-301,315-377; row a5).
 *
 *   k_policy_act : actions for N environments in ONE launch: obs rows (read in place from the replay ring) -> 256 -> 256 -> (mean | log_std)
 *                  -> a = tanh(mean + exp(clamp(log_std)) eps)  (eps: Philox keyed by (seed, GLOBAL env id, step), so a run does not depend
 *                  on how envs are sharded over GPUs) or a = tanh(mean) (MakeDeterministic), written straight into the ring's action rows
 *                  that k_step reads.  fp32 on the CUDA cores: the forward is 0.65 GFLOP for 4096 envs against 3.4 ms of k_step, and exact
 *                  fp32 keeps the ring contents bit-comparable with the copy-based path; the tensor-core GEMMs are the SAC update's.
 *   k_path_stats : Rewards / Returns / ExplReturns / Actions mean, std, max, min of a collection round, reduced on the device from the ring
 *                  segment the round wrote (fp64 accumulators, fixed order: deterministic).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/rsb_sac.h"

void rsb_sac_set_error(const char *msg);
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rsb_sac_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); return 1; } } while (0)

namespace {

__device__ __forceinline__ void philox4(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0], h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
/* same Box-Muller as the env kernels (csrc/rsb_dev.h box_muller): the two transcendental arguments in double, so that the host oracle
   (numpy double) reproduces the normals to fp32 round-off */
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float *z0, float *z1) {
  double u1 = ((double)a + 0.5) * (1.0 / 4294967296.0), u2 = ((double)b + 0.5) * (1.0 / 4294967296.0);
  double rad = sqrt(-2.0 * log(u1)), ang = 2.0 * 3.14159265358979323846 * u2;
  *z0 = (float)(rad * cos(ang)); *z1 = (float)(rad * sin(ang));
}
__device__ __forceinline__ long ring_row(long slot0, long cap, long e) { if (cap <= 0) return e; long r = slot0 + e; return r >= cap ? r - cap : r; }

#define PA_ROWS 32            /* environments per CTA */
#define PA_LD 36              /* row length (floats) of the transposed activation tiles: 32 + 4 keeps float4 rows 16-byte aligned and the
                                 per-thread float4 stores of layer outputs conflict-free (thread j writes row j: bank offset 4j mod 32) */
#define PA_HID 256
#define PA_THREADS 256

struct PolicyArgs {
  const float *W0, *b0, *W1, *b1, *W2, *b2;       /* weights [in, out] row-major (the ParamStore layout), hidden = 256 */
  const float *obs; float *act;                   /* row-addressed arrays (ring or plain) */
  long obs_ld, act_ld, slot0, cap;
  int O, A, n, deterministic;
  uint64_t seed, env_id_base, step;
};

/* one hidden layer for the CTA's 32 rows: thread j owns output column j, the 32 row accumulators live in registers; the input tile is read as
   warp-uniform float4 (one shared-memory wavefront feeds 4 FMAs of every lane), the weight column element by coalesced global loads (L2-resident:
   the whole policy is 80 k floats) */
__device__ __forceinline__ void dense_relu(const float *__restrict__ W, const float *__restrict__ bias, int K, const float *inT, float *outT, int j) {
  float acc[PA_ROWS];
#pragma unroll
  for (int r = 0; r < PA_ROWS; r++) acc[r] = 0.0f;
  /* the weight column is streamed in groups of 8 elements, the next group in flight while the current one is consumed: with one CTA of 8 warps
     per SM a single load ahead (first version: 110 us per launch, 17 % issue slots) leaves the L2 latency exposed on every k */
  constexpr int G = 8;
  float w[G], wn[G];
#pragma unroll
  for (int u = 0; u < G; u++) w[u] = (u < K) ? W[(size_t)u * PA_HID + j] : 0.0f;
  for (int k0 = 0; k0 < K; k0 += G) {
#pragma unroll
    for (int u = 0; u < G; u++) wn[u] = (k0 + G + u < K) ? W[(size_t)(k0 + G + u) * PA_HID + j] : 0.0f;
#pragma unroll
    for (int u = 0; u < G; u++) {
      if (k0 + u < K) {
        const float4 *x4 = reinterpret_cast<const float4 *>(inT + (k0 + u) * PA_LD);
#pragma unroll
        for (int q = 0; q < PA_ROWS / 4; q++) { const float4 x = x4[q]; acc[4 * q] = fmaf(w[u], x.x, acc[4 * q]); acc[4 * q + 1] = fmaf(w[u], x.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(w[u], x.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w[u], x.w, acc[4 * q + 3]); }
      }
    }
#pragma unroll
    for (int u = 0; u < G; u++) w[u] = wn[u];
  }
  const float b = bias[j]; float4 *o4 = reinterpret_cast<float4 *>(outT + j * PA_LD);
#pragma unroll
  for (int q = 0; q < PA_ROWS / 4; q++) o4[q] = make_float4(fmaxf(acc[4 * q] + b, 0.0f), fmaxf(acc[4 * q + 1] + b, 0.0f), fmaxf(acc[4 * q + 2] + b, 0.0f), fmaxf(acc[4 * q + 3] + b, 0.0f));
}

__global__ void __launch_bounds__(PA_THREADS) k_policy_act(const PolicyArgs p) {
  extern __shared__ __align__(16) float sm[];
  float *xT = sm, *h1T = xT + (size_t)p.O * PA_LD, *h2T = h1T + PA_HID * PA_LD, *outT = h2T + PA_HID * PA_LD;
  const int tid = threadIdx.x; const long e0 = (long)blockIdx.x * PA_ROWS;
  /* observation tile, transposed: xT[k][r] = obs[row(e0 + r)][k]; rows beyond n are zero (their actions are not stored) */
  for (int i = tid; i < PA_ROWS * p.O; i += PA_THREADS) { const int r = i / p.O, k = i - r * p.O; const long e = e0 + r;
    xT[k * PA_LD + r] = e < p.n ? p.obs[(size_t)ring_row(p.slot0, p.cap, e) * p.obs_ld + k] : 0.0f; }
  __syncthreads();
  dense_relu(p.W0, p.b0, p.O, xT, h1T, tid);
  __syncthreads();
  dense_relu(p.W1, p.b1, PA_HID, h1T, h2T, tid);
  __syncthreads();
  /* head: out[r][c], c < 2A -- thread (r = tid % 32, c = tid / 32 + 8 m): the weight loads are warp-uniform, the h2 loads conflict-free */
  { const int r = tid & 31, c0 = tid >> 5, A2 = 2 * p.A; float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int k = 0; k < PA_HID; k++) { const float h = h2T[k * PA_LD + r]; const float *w = p.W2 + (size_t)k * A2;
#pragma unroll
      for (int m = 0; m < 4; m++) { const int c = c0 + 8 * m; if (c < A2) acc[m] = fmaf(h, w[c], acc[m]); } }
#pragma unroll
    for (int m = 0; m < 4; m++) { const int c = c0 + 8 * m; if (c < A2) outT[c * PA_LD + r] = acc[m] + p.b2[c]; } }
  __syncthreads();
  /* tanh-Gaussian head (rlkit TanhGaussianPolicy.get_action): thread (r, blk) handles action dims 4 blk .. 4 blk + 3 of row r; its four
     normals come from ONE Philox block keyed (seed; global env id, stream 2, step * 8 + blk) */
  { const int r = tid & 31, blk = tid >> 5; const long e = e0 + r;
    if (e < p.n && 4 * blk < p.A) {
      float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (!p.deterministic) { const uint64_t id = p.env_id_base + (uint64_t)e; uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), 2u, (uint32_t)(p.step * 8u + (uint64_t)blk)};
        philox4(c, (uint32_t)p.seed, (uint32_t)(p.seed >> 32)); box_muller(c[0], c[1], &z[0], &z[1]); box_muller(c[2], c[3], &z[2], &z[3]); }
      float *dst = p.act + (size_t)ring_row(p.slot0, p.cap, e) * p.act_ld;
#pragma unroll
      for (int k = 0; k < 4; k++) { const int d = 4 * blk + k; if (d < p.A) {
        const float mean = outT[d * PA_LD + r]; float a;
        if (p.deterministic) a = tanhf(mean);
        else { const float ls = fminf(fmaxf(outT[(p.A + d) * PA_LD + r], -20.0f), 2.0f); a = tanhf(mean + expf(ls) * z[k]); }
        dst[d] = a; } }
    } }
}

/* ---------------------------------------------------------------- path statistics of one collection round
   The round wrote transition (env i, step t) to ring row (slot0 + t n + i) mod cap.  out (double):
     [0..3]  rewards  : sum, sum of squares, max, min            over all T n transitions
     [4..7]  returns  : sum, sum of squares, max, min            over the n paths (return = sum of a path's rewards)
     [8..11] returns accumulated over the first `expl_len` steps (util/rlkit_custom.py:338-340 "ExplReturns")
       [12..15] actions : sum, sum of squares, max, min            over all T n A action entries
   CTA c walks envs 256 c .. 256 c + 255 (one thread per env: coalesced over envs at every t) and leaves its 16 partial values in
   out[16 (1 + c) ..]; k_path_stats_final adds the partials in CTA order.  Fixed reduction order => deterministic. */
#define PS_THREADS 256
__device__ __forceinline__ void red4(double &s, double &q, double &mx, double &mn, double *sh, int tid) {
  sh[tid] = s; sh[PS_THREADS + tid] = q; sh[2 * PS_THREADS + tid] = mx; sh[3 * PS_THREADS + tid] = mn; __syncthreads();
  for (int o = PS_THREADS / 2; o > 0; o >>= 1) { if (tid < o) { sh[tid] += sh[tid + o]; sh[PS_THREADS + tid] += sh[PS_THREADS + tid + o];
      sh[2 * PS_THREADS + tid] = fmax(sh[2 * PS_THREADS + tid], sh[2 * PS_THREADS + tid + o]); sh[3 * PS_THREADS + tid] = fmin(sh[3 * PS_THREADS + tid], sh[3 * PS_THREADS + tid + o]); } __syncthreads(); }
  s = sh[0]; q = sh[PS_THREADS]; mx = sh[2 * PS_THREADS]; mn = sh[3 * PS_THREADS]; __syncthreads();
}
__global__ void __launch_bounds__(PS_THREADS) k_path_stats(const float *__restrict__ rew, const float *__restrict__ act, long slot0, long cap, int n, int T, int A, int expl_len, double *__restrict__ out) {
  __shared__ double shd[4 * PS_THREADS];
  const int tid = threadIdx.x, i = blockIdx.x * PS_THREADS + tid; const double inf = 1e300;
  double rs = 0, rq = 0, rmx = -inf, rmn = inf, Rs = 0, Rq = 0, Rmx = -inf, Rmn = inf, Es = 0, Eq = 0, Emx = -inf, Emn = inf, as = 0, aq = 0, amx = -inf, amn = inf;
  if (i < n) {
    double ret = 0, eret = 0;
    for (int t = 0; t < T; t++) {
      long row = (slot0 + (long)t * n + i) % cap;
      const double r = (double)rew[row]; ret += r; if (t < expl_len) eret += r;
      rs += r; rq += r * r; rmx = fmax(rmx, r); rmn = fmin(rmn, r);
      const float *a = act + (size_t)row * A;
      for (int d = 0; d < A; d++) { const double v = (double)a[d]; as += v; aq += v * v; amx = fmax(amx, v); amn = fmin(amn, v); }
    }
    Rs = ret; Rq = ret * ret; Rmx = Rmn = ret; Es = eret; Eq = eret * eret; Emx = Emn = eret;
  }
  red4(rs, rq, rmx, rmn, shd, tid); red4(Rs, Rq, Rmx, Rmn, shd, tid); red4(Es, Eq, Emx, Emn, shd, tid); red4(as, aq, amx, amn, shd, tid);
  if (tid == 0) { double *o = out + 16 * (1 + blockIdx.x); o[0] = rs; o[1] = rq; o[2] = rmx; o[3] = rmn; o[4] = Rs; o[5] = Rq; o[6] = Rmx; o[7] = Rmn;
    o[8] = Es; o[9] = Eq; o[10] = Emx; o[11] = Emn; o[12] = as; o[13] = aq; o[14] = amx; o[15] = amn; }
}
__global__ void k_path_stats_final(double *__restrict__ out, int nblk) {
  const int k = threadIdx.x; if (k >= 16) return;
  const int kind = k & 3; double v = out[16 + k];
  for (int c = 1; c < nblk; c++) { const double w = out[16 * (1 + c) + k]; v = kind < 2 ? v + w : (kind == 2 ? fmax(v, w) : fmin(v, w)); }
  out[k] = v;
}

}  // namespace

extern "C" {

int rsb_policy_act(const float *d_W0, const float *d_b0, const float *d_W1, const float *d_b1, const float *d_W2, const float *d_b2,
                   int obs_dim, int act_dim, int hidden, const float *d_obs, long obs_ld, float *d_act, long act_ld, int64_t slot0, int64_t cap, int n,
                   int deterministic, uint64_t seed, uint64_t env_id_base, uint64_t step, void *stream) {
  if (hidden != PA_HID) { rsb_sac_set_error("policy_act: the kernel is built for the benchmark's 256-wide hidden layers"); return 2; }
  if (obs_dim < 1 || obs_dim > 512 || act_dim < 1 || act_dim > 16 || n < 1) { rsb_sac_set_error("policy_act: obs_dim in [1,512], act_dim in [1,16], n >= 1 required"); return 2; }
  if (cap > 0 && (slot0 < 0 || slot0 >= cap || n > cap)) { rsb_sac_set_error("policy_act: ring slot out of range"); return 2; }
  const size_t smem = ((size_t)obs_dim + 2 * PA_HID + 32) * PA_LD * sizeof(float);
  static size_t s_limit[64]; int dev = 0; CKC(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && s_limit[dev] < smem) { CKC(cudaFuncSetAttribute(k_policy_act, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); s_limit[dev] = smem; }
  PolicyArgs p{d_W0, d_b0, d_W1, d_b1, d_W2, d_b2, d_obs, d_act, obs_ld, act_ld, (long)slot0, (long)cap, obs_dim, act_dim, n, deterministic, seed, env_id_base, step};
  k_policy_act<<<(n + PA_ROWS - 1) / PA_ROWS, PA_THREADS, smem, (cudaStream_t)stream>>>(p);
  CKC(cudaGetLastError()); return 0;
}

int rsb_path_stats_words(int n) { return 16 * (1 + (n + PS_THREADS - 1) / PS_THREADS); }
int rsb_path_stats(const float *d_rewards, const float *d_actions, int64_t slot0, int64_t cap, int n, int T, int act_dim, int expl_len, double *d_out, void *stream) {
  if (n < 1 || T < 1 || cap < 1 || slot0 < 0 || slot0 >= cap) { rsb_sac_set_error("path_stats: bad arguments"); return 2; }
  const int nblk = (n + PS_THREADS - 1) / PS_THREADS;
  k_path_stats<<<nblk, PS_THREADS, 0, (cudaStream_t)stream>>>(d_rewards, d_actions, (long)slot0, (long)cap, n, T, act_dim, expl_len, d_out);
  k_path_stats_final<<<1, 32, 0, (cudaStream_t)stream>>>(d_out, nblk);
  CKC(cudaGetLastError()); return 0;
}

}  /* extern "C" */
