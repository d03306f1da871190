/*
 * rsb_dp.cu -- data-parallel SAC: the gradient all-reduce FUSED into the optimizer kernel, over NVLink peer memory (C-ABI: include/rsb_sac.h,
 * "data parallel").
 *
 * The reference has no parallelism; the batched system adds exactly one collective: the mean of the flat gradient bucket over
 * the ranks, once per update.  With NCCL that is graph(gradients) -> ncclAllReduce -> graph(Adam): a 0.95 MB message costs
 * ~36 us per update at 8 ranks (profiles/r1), a third of the update itself.  Here the ranks' gradient buckets live in SYMMETRIC memory
 * (torch.distributed._symmetric_memory: every rank holds device pointers to every rank's bucket and flag words), and ONE kernel does
 *
 *     cross-rank barrier "gradients of update e are complete"     (flag words written with st.release.sys through NVLink, polled with ld.acquire.sys)
 *     g[i] = (1 / W) * sum_{r = 0..W-1} bucket_r[i]               (peer loads, 16 bytes per lane, fixed rank order: every rank gets the SAME bits)
 *     Adam + Polyak on g[i]                                       (as k_adam_polyak, rsb_sac.cu)
 *     signal "rank done reading update e"                         (the next update's head kernel waits for it before it touches the bucket)
 *
 * so the whole data-parallel update is one CUDA-graph replay, with no host synchronisation point between gradients and optimizer.
 * Every spin is bounded; a wait that gives up raises a device counter (rsb_dp_timeouts) that the training loop checks once per epoch.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/rsb_sac.h"
#include "rsb_pdl.h"

void rsb_sac_set_error(const char *msg);
#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rsb_sac_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); return 1; } } while (0)

#define DP_MAX_WORLD 16
#define DP_READY 0            /* flag word [DP_READY + r]: rank r's bucket is complete for epoch <value> */
#define DP_DONE 16            /* flag word [DP_DONE + r]:  rank r has finished reading my bucket for epoch <value> */
#define DP_SPIN (1 << 24)

__device__ unsigned int g_dp_timeouts;
__device__ unsigned long long g_dp_clk[8];      /* diagnostic (rsb_dp_debug_clocks): %globaltimer [ns] of CTA 0 of the last launch at entry, dependency satisfied, READY signalled,
                                                   peers ready, go flag set, own reduction + Adam done, kernel tail */
#define DPCLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_dp_clk[i] = t_; } } while (0)

namespace {
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory"); return v; }
/* flag stores without a release fence: what the flag announces was written by EARLIER KERNELS of this stream (the gradient bucket) or has been consumed by this
   thread's own dependent arithmetic (the peers' buckets) -- both are globally performed when the store issues; a sys-scope release costs ~2 us here (measured) */
__device__ __forceinline__ void st_relaxed_sys(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_peer4(const float *p) { float4 v; asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float ld_peer1(const float *p) { float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory"); return v; }
/* (epochs only grow and never wrap in practice: 2^32 updates) */
__device__ __forceinline__ bool wait_ge_sys(const uint32_t *p, uint32_t v) {
  for (int it = 0; it < DP_SPIN; it++) { if ((int32_t)(ld_acquire_sys(p) - v) >= 0) return true; __nanosleep(20); }
  atomicAdd(&g_dp_timeouts, 1u); return false;
}

struct DpArgs {
  const float *const *peer_grads;       /* device array [world]: every rank's gradient bucket (symmetric memory) */
  uint32_t *const *peer_flags;          /* device array [world]: every rank's flag words (symmetric memory, >= 32 words, zero at start) */
  uint32_t *local;                      /* this rank's private words: [0] epoch of the last completed update, [1] go flag, [2] CTA completion count */
  int rank, world;
};

__device__ __forceinline__ void adam_one(float g, long i, float *p, float *m, float *v, double lr_pi, double lr_q, float b1, float b2, float eps, const double *bc,
                                         float *tgt, long tgt_begin, long tgt_end, float tau, int do_soft, float *alpha_out, long log_alpha_idx, float *g_local) {
  const bool isq = i >= tgt_begin && i < tgt_end;
  const float step_size = (float)((isq ? lr_q : lr_pi) / bc[0]), bc2s = (float)bc[1];
  const float mi = b1 * m[i] + (1.0f - b1) * g, vi = b2 * v[i] + (1.0f - b2) * g * g; m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2s + eps, pi = p[i] - step_size * (mi / denom); p[i] = pi;
  if (do_soft && isq) { const long k = i - tgt_begin; tgt[k] = (1.0f - tau) * tgt[k] + tau * pi; }
  if (i == log_alpha_idx) { alpha_out[1] = pi; alpha_out[0] = (float)exp((double)pi); }
  (void)g_local;
}

/* persistent grid: gridDim.x CTAs, all resident (<= one CTA per SM is launched) */
__global__ void __launch_bounds__(512) k_adam_polyak_allreduce(DpArgs a, float *__restrict__ p, float *__restrict__ m, float *__restrict__ v, long n,
                                                               double lr_pi, double lr_q, float b1, float b2, float eps, const double *__restrict__ bc,
                                                               float *__restrict__ tgt, long tgt_begin, long tgt_end, float tau, int do_soft,
                                                               float *__restrict__ alpha_out, long log_alpha_idx) {
  DPCLK(0);
  pdl_wait(); pdl_trigger();
  DPCLK(1);
  const int tid = threadIdx.x;
  const uint32_t epoch = a.local[0] + 1u;
  /* ---- barrier: every rank's bucket is complete.  CTA 0 talks to the peers, the other CTAs wait for its go flag. */
  if (blockIdx.x == 0) {
    if (tid < a.world) st_relaxed_sys(a.peer_flags[tid] + DP_READY + a.rank, epoch);
    DPCLK(2);
    if (tid < a.world) wait_ge_sys(a.peer_flags[a.rank] + DP_READY + tid, epoch);
    __syncthreads();
    DPCLK(3);
    if (tid == 0) st_release_gpu(a.local + 1, epoch);
    DPCLK(4);
  } else {
    if (tid == 0) { for (int it = 0; it < DP_SPIN; it++) { if ((int32_t)(ld_acquire_gpu(a.local + 1) - epoch) >= 0) break; __nanosleep(20); } }
    __syncthreads();
  }
  /* ---- mean over the ranks in fixed order + Adam.  float4 body, scalar tail; the buckets are 16-byte aligned (symmetric allocations). */
  const float inv = 1.0f / (float)a.world;
  const long n4 = n >> 2, stride = (long)gridDim.x * blockDim.x;
  for (long q = (long)blockIdx.x * blockDim.x + tid; q < n4; q += stride) {
    /* all ranks' loads are issued before the first one is used: one NVLink round trip per element, not `world` of them in sequence */
    float4 x[DP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; r++) if (r < a.world) x[r] = ld_peer4(a.peer_grads[r] + 4 * q);
    float4 s = x[0];
#pragma unroll
    for (int r = 1; r < DP_MAX_WORLD; r++) if (r < a.world) { s.x += x[r].x; s.y += x[r].y; s.z += x[r].z; s.w += x[r].w; }      /* fixed rank order: same bits on every rank */
    const float g4[4] = {s.x * inv, s.y * inv, s.z * inv, s.w * inv};
#pragma unroll
    for (int k = 0; k < 4; k++) adam_one(g4[k], 4 * q + k, p, m, v, lr_pi, lr_q, b1, b2, eps, bc, tgt, tgt_begin, tgt_end, tau, do_soft, alpha_out, log_alpha_idx, nullptr);
  }
  if (blockIdx.x == 0 && tid < (int)(n & 3)) {
    const long i = (n4 << 2) + tid; float s = ld_peer1(a.peer_grads[0] + i);
    for (int r = 1; r < a.world; r++) s += ld_peer1(a.peer_grads[r] + i);
    adam_one(s * inv, i, p, m, v, lr_pi, lr_q, b1, b2, eps, bc, tgt, tgt_begin, tgt_end, tau, do_soft, alpha_out, log_alpha_idx, nullptr);
  }
  /* ---- the last CTA to finish tells every peer that this rank no longer reads their buckets, and closes the epoch */
  __syncthreads();
  DPCLK(5);
  __shared__ unsigned int last;
  if (tid == 0) { __threadfence(); last = (atomicAdd(a.local + 2, 1u) == gridDim.x - 1) ? 1u : 0u; }
  __syncthreads();
  if (last) {
    if (tid < a.world) st_relaxed_sys(a.peer_flags[tid] + DP_DONE + a.rank, epoch);      /* (every CTA passed a __threadfence() after its last peer load) */
    if (tid == 0) { a.local[2] = 0u; a.local[0] = epoch; }
  }
  DPCLK(6);
}

/* the next update may overwrite this rank's bucket only after every peer has finished reading it (epoch = last completed update) */
__global__ void k_dp_wait_peers_done(DpArgs a) {
  pdl_wait(); pdl_trigger();
  const uint32_t epoch = a.local[0];
  if ((int)threadIdx.x < a.world) wait_ge_sys(a.peer_flags[a.rank] + DP_DONE + threadIdx.x, epoch);
}
}  // namespace

extern "C" {

int rsb_dp_debug_clocks(unsigned long long *host_out8) {
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  return cudaMemcpyFromSymbol(host_out8, g_dp_clk, 8 * sizeof(unsigned long long)) != cudaSuccess;
}

int rsb_dp_timeouts(void) {
  unsigned int h = 0, z = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&h, g_dp_timeouts, sizeof(h)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_dp_timeouts, &z, sizeof(z));
  return (int)h;
}

int rsb_dp_wait_peers_done(const float *const *d_peer_grads, uint32_t *const *d_peer_flags, uint32_t *d_local, int rank, int world, void *stream) {
  if (world < 2 || world > DP_MAX_WORLD || rank < 0 || rank >= world) { rsb_sac_set_error("dp: world in [2, 16] and 0 <= rank < world required"); return 2; }
  DpArgs a{d_peer_grads, d_peer_flags, d_local, rank, world};
  CKD(rsb_launch_pdl(k_dp_wait_peers_done, dim3(1), dim3(32), 0, (cudaStream_t)stream, 1, a));
  return 0;
}

int rsb_adam_polyak_allreduce(const float *const *d_peer_grads, uint32_t *const *d_peer_flags, uint32_t *d_local, int rank, int world,
                              float *d_p, float *d_m, float *d_v, long n, double lr_pi, double lr_q, float b1, float b2, float eps, double *d_bc,
                              float *d_tgt, long tgt_begin, long tgt_end, float tau, int do_soft, float *d_alpha, long log_alpha_idx, void *stream) {
  if (world < 2 || world > DP_MAX_WORLD || rank < 0 || rank >= world) { rsb_sac_set_error("dp: world in [2, 16] and 0 <= rank < world required"); return 2; }
  static int sms[64]; int dev = 0; CKD(cudaGetDevice(&dev));
  if (dev < 64 && sms[dev] == 0) { cudaDeviceProp prop; CKD(cudaGetDeviceProperties(&prop, dev)); sms[dev] = prop.multiProcessorCount; }
  int grid = dev < 64 ? sms[dev] : 128;
  const long need = ((n >> 2) + 511) / 512; if (need < grid) grid = (int)(need > 0 ? need : 1);
  DpArgs a{d_peer_grads, d_peer_flags, d_local, rank, world};
  CKD(rsb_launch_pdl(k_adam_polyak_allreduce, dim3(grid), dim3(512), 0, (cudaStream_t)stream, 1, a, d_p, d_m, d_v, n, lr_pi, lr_q, b1, b2, eps, (const double *)d_bc,
                     d_tgt, tgt_begin, tgt_end, tau, do_soft, d_alpha, log_alpha_idx));
  return 0;
}

}  /* extern "C" */
