/*
 * rsb_tc_gemm.cu -- the dense products of the SAC update on the sm_100a tensor cores (C-ABI: include/rsb_gemm.h).
 *
 * Stands behind the Linear layers of rlkit's FlattenMlp / TanhGaussianPolicy in SACTrainer.train_from_torch (reference call site
 * util/rlkit_custom.py:238): forward X W + b (+ReLU), activation gradients dY W^T (masked by the ReLU), weight gradients X^T dY.
 *
 *   C[b][m,n] = epilogue( sum_k A[b][m,k] * B[b][k,n] )       fp32 in HBM, TF32 x TF32 -> FP32 on tcgen05, accumulator in tensor memory
 *
 * One CTA of 256 threads owns a 128 x n_tile tile of C (n_tile in {16,32,64,128}, chosen per launch so small-batch updates still spread
 * over many SMs).  The contraction runs in chunks of 64:
 *   - every thread stages its share of the A chunk (128 x 64) and the B chunk (n_tile x 64) with `cp.async` straight into the tensor
 *     core's canonical K-major no-swizzle layout: core matrices of 8 rows x 16 bytes (4 fp32), 128 bytes each, the 16 core matrices of
 *     one 8-row group back to back (so: byte offset of element (r, k) = (r/8)*2048 + (k/4)*128 + (r%8)*16 + (k%4)*4).  Because the
 *     loader addresses elements by (row stride, column stride), transposed operands cost nothing extra: X^T and W^T are read in place.
 *     A lane takes row (lane%8) and 16-byte column group (lane/8) of a 8 x 16 element patch, so a quarter-warp writes one whole core
 *     matrix (128 contiguous bytes, conflict-free) and reads full 32-byte sectors; operands whose contraction index is contiguous and
 *     16-byte aligned move as 16-byte copies, everything else as 4-byte copies; out-of-range rows/columns are zero-filled (src-size 0).
 *   - four stages (three at n_tile = 128): the copies of the first chunks leave before tensor memory is even allocated, so a K = 256 layer
 *     has its whole operand set in flight at once; per chunk one elected thread issues `tcgen05.mma.cta_group::1.kind::tf32` (M = 128,
 *     N = n_tile, K = 8 each: two core matrices along K, descriptor start address advanced by 256 bytes) and `tcgen05.commit`s to the
 *     stage's mbarrier, which is what frees the stage for the chunk S further on.
 *   - epilogue: the 8 warps read the accumulator with `tcgen05.ld.32x32b.x16` (warp w owns tensor-memory lanes 32*(w%4).., the two warps
 *     of a lane quarter split the columns), apply bias / ReLU / ReLU-mask / accumulate and store rows of C.
 * Shared-memory matrix descriptor (64 bit): [0,14) start address >> 4, [16,30) leading byte offset >> 4 = distance between the two core
 * matrices along K (128), [32,46) stride byte offset >> 4 = distance between 8-row groups (2048), [46,48) = 1 (sm_100), [61,64) = 0 (no swizzle).
 * Instruction descriptor (32 bit): [4,6) D format 1 = F32, [7,10) A format 2 = TF32, [10,13) B format 2 = TF32, bits 15/16 = 0 (both K-major),
 * [17,23) N >> 3, [24,29) M >> 4.
 * Every mbarrier wait is bounded: a wait that gives up raises a device-side counter (rsb_gemm_timeouts(), checked by the training loop once
 * per epoch and by the tests) AND the CTA writes NaN into its tile of C, so a protocol error reaches the losses and the parameters as NaN --
 * never a hung GPU, never silently stale activations or gradients.
 */
#include <cuda.h>                 /* CUtensorMap + enums only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no link against libcuda) */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "../../include/rsb_gemm.h"
#include "../../include/rsb_sac.h"
#include "rsb_pdl.h"

void rsb_sac_set_error(const char *msg);   /* rsb_sac.cu: the string rsb_sac_last_error() returns */

#define BM 128
#define KC 64
#define BN_MAX 128
#define NTHREADS 256
#define A_BYTES (BM * KC * 4)
#define B_BYTES (BN_MAX * KC * 4)
#define CTRL_BYTES 1024                         /* control block in front of the stages: 4 mbarriers, tensor-memory slot, bias row */
#define SMEM_MAX 232448                         /* 227 KB: the most one CTA may ask for */
#define TMEM_COLS 128

struct GemmArgs {
  const float *a, *b, *bias, *mask;
  float *c;
  long long a_rs, a_cs, a_bs, b_ks, b_ns, b_bs, c_rs, c_bs, bias_bs, mask_rs, mask_bs, a_kbs, b_kbs;
  int m, n, k, flags, n_tile, lbo16, sbo16, splits, cps, recv_off, stages, k_block;   /* splits: CTAs of one cluster sharing a C tile along K; cps: chunks per split */
  int a_tma, b_tma;         /* operand staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) instead of the strided cp.async gather: 1 = K-major (contraction index contiguous
                               in memory), 2 = MN-major (row / column index contiguous: transposed activations, weights stored [in, out]); 0 = gather.  16-byte strides needed. */
  int z_is_block;           /* third TMA coordinate: 1 = the k-block index (stack_k), 0 = the batch index */
  int a_zbcast, b_zbcast;   /* the operand is shared by all batch entries (batch stride 0: both Q networks read the same input): its map has one slice, z = 0 */
};
#define FBAR_OFF 640        /* four "stage full" mbarriers (TMA completion) inside the control block, behind the bias row */

__device__ unsigned int g_gemm_timeouts;
__device__ int g_mn_swap;                     /* diagnostic (rsb_gemm_debug_mn_swap): exchange the two byte-offset fields of the MN-major descriptor */
__device__ long long g_gemm_clk[12];         /* phase clocks of CTA (0,0,0)'s thread 0 of the last launch (diagnostic, rsb_gemm_debug_clocks) */
#define CLK(i) do { if (tid == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) g_gemm_clk[i] = clock64(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

/* one box of a 3-D tensor map (k, row, z) -> shared memory; completion is reported in bytes on the stage's "full" barrier */
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int it = 0; it < (1 << 20); it++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  atomicAdd(&g_gemm_timeouts, 1u);
  return false;
}

/* stage `rows` x KC elements of an operand: element (r, k) of the tile lives at base + r*rs + k*cs; rows >= rvalid and columns >= kvalid are zero.
   A thread keeps its (row in the core matrix, column) and walks down the 8-row groups with constant pointer increments. */
__device__ __forceinline__ void load_tile(uint32_t dst, const float *base, long long rs, long long cs, int rows, int rvalid, int kvalid, bool vec16) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rr = lane & 7, kq = lane >> 3;
  if (vec16) {                                        /* lane -> row rr of the group, 16-byte column k4 = 4*(warp%4) + kq; warps 0-3 / 4-7 take even / odd groups */
    const int k4 = (warp & 3) * 4 + kq, kb = min(16, max(0, (kvalid - k4 * 4) * 4));
    int r = (warp >> 2) * 8 + rr;
    const float *src = base + (long long)r * rs + k4 * 4;
    uint32_t d = dst + (warp >> 2) * 2048 + k4 * 128 + rr * 16;
    for (int j = 0; j < (rows >> 4); j++, r += 16, src += 16 * rs, d += 4096) {
      const int bytes = (r < rvalid) ? kb : 0;
      cp_async16(d, bytes ? src : base, bytes);
    }
  } else {                                            /* lane -> row rr, element kq of the 16-byte columns k4 = warp and warp + 8 */
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k4 = warp + 8 * h, k = k4 * 4 + kq;
      const bool okk = k < kvalid;
      int r = rr;
      const float *src = base + (long long)rr * rs + (long long)k * cs;
      uint32_t d = dst + k4 * 128 + rr * 16 + kq * 4;
      for (int rg = 0; rg < (rows >> 3); rg++, r += 8, src += 8 * rs, d += 2048) {
        const bool ok = okk && (r < rvalid);
        cp_async4(d, ok ? src : base, ok ? 4 : 0);
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tf32(GemmArgs g, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = blockIdx.x % g.splits;                       /* = %cluster_ctarank: the cluster is (splits, 1, 1) */
  const int n0 = (blockIdx.x / g.splits) * g.n_tile, m0 = blockIdx.y * BM, bz = blockIdx.z;
  const uint32_t bar0 = smem_u32(smem), tslot = bar0 + 32, rbar = bar0 + 40, sbase = bar0 + CTRL_BYTES;
  volatile uint32_t *tslot_p = (volatile uint32_t *)(smem + 32);
  float *sbias = (float *)(smem + 64);
  const int S = g.stages;                   /* stages: 3 x 64 KB or 4 x (32 KB + n_tile x 256 B) */
  const uint32_t stage_bytes = A_BYTES + (uint32_t)g.n_tile * (KC * 4);

  const int kbeg = rank * g.cps * KC, klen = min(g.k, kbeg + g.cps * KC) - kbeg;        /* this CTA's slice of the contraction */
  /* the contraction index may run over `k / k_block` separately strided blocks (two networks' activations side by side): chunk at absolute
     k sits at block k / k_block, offset k % k_block; k_block is a multiple of the chunk, so a chunk never straddles blocks */
  const float *A0 = g.a + (long long)bz * g.a_bs + (long long)m0 * g.a_rs;
  const float *B0 = g.b + (long long)bz * g.b_bs + (long long)n0 * g.b_ns;
  auto a_at = [&](int kabs) { const int blk = kabs / g.k_block; return A0 + (long long)blk * g.a_kbs + (long long)(kabs - blk * g.k_block) * g.a_cs; };
  auto b_at = [&](int kabs) { const int blk = kabs / g.k_block; return B0 + (long long)blk * g.b_kbs + (long long)(kabs - blk * g.k_block) * g.b_ks; };
  const float *A = a_at(kbeg), *B = b_at(kbeg);
  const bool a16 = g.a_cs == 1 && (g.a_rs & 3) == 0 && (g.a_kbs & 3) == 0 && ((uintptr_t)A0 & 15) == 0;
  const bool b16 = g.b_ks == 1 && (g.b_ns & 3) == 0 && (g.b_kbs & 3) == 0 && ((uintptr_t)B0 & 15) == 0;
  const int mvalid = g.m - m0, nvalid = g.n - n0, nchunks = (klen + KC - 1) / KC;

  CLK(0);
  /* prologue that touches no data of an earlier kernel: tensor memory, barriers.  Launched with programmatic stream serialization
     (rsb_pdl.h) this part runs while the PRECEDING kernel of the chain is still working. */
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int c = 0; c < 4; c++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0 + 8 * c) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(rbar) : "memory");
#pragma unroll
    for (int c = 0; c < 4; c++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0 + FBAR_OFF + 8 * c) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    /* split-K: this CTA will receive splits x 128 x (n_tile/splits) partial sums = 512 x n_tile bytes on `rbar` (one phase) */
    if (g.splits > 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(rbar), "r"(512 * g.n_tile) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tslot_p;
  /* split-K: tell the cluster that this CTA runs and its receive barrier is armed; waited for just before the first remote store */
  if (g.splits > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  pdl_wait();                                                   /* the producer of A / B / mask / C has completed */
  pdl_trigger();
  CLK(1);
  /* TMA staging of a chunk (one elected thread): the chunk's 64 k-values are two 128-byte swizzle atoms [rows x 32 floats]; rows / k beyond the tensor are
     zero-filled by the TMA unit.  Coordinates (k inside the block, row, z) with z = batch index or k-block index. */
  const uint32_t tma_bytes = (g.a_tma ? (uint32_t)A_BYTES : 0u) + (g.b_tma ? (uint32_t)g.n_tile * (KC * 4) : 0u);
  auto tma_chunk = [&](int stage, int krel) {                   /* krel: offset of the chunk inside this CTA's K slice */
    const uint32_t fb = bar0 + FBAR_OFF + 8 * stage, a_dst = sbase + stage * stage_bytes, b_dst = a_dst + A_BYTES;
    const int kabs = kbeg + krel, blk = kabs / g.k_block, kin = kabs - blk * g.k_block, z = g.z_is_block ? blk : bz;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(tma_bytes) : "memory");
    /* K-major: two atoms [rows x 32 k] (box 32 k x rows).  MN-major: rows/32 atoms [64 k x 32 rows] (box 32 rows x 64 k), 8 KB apart */
    const int za = g.a_zbcast ? 0 : z, zb = g.b_zbcast ? 0 : z;
    if (g.a_tma == 1) { tma_load_3d(a_dst, &tmap_a, kin, m0, za, fb); tma_load_3d(a_dst + BM * 128, &tmap_a, kin + 32, m0, za, fb); }
    else if (g.a_tma == 2) { for (int j = 0; j < BM / 32; j++) tma_load_3d(a_dst + j * (KC * 128), &tmap_a, m0 + 32 * j, kin, za, fb); }
    if (g.b_tma == 1) { tma_load_3d(b_dst, &tmap_b, kin, n0, zb, fb); tma_load_3d(b_dst + g.n_tile * 128, &tmap_b, kin + 32, n0, zb, fb); }
    else if (g.b_tma == 2) { for (int j = 0; j < (g.n_tile >> 5); j++) tma_load_3d(b_dst + j * (KC * 128), &tmap_b, n0 + 32 * j, kin, zb, fb); }
  };
  const bool any_tma = (g.a_tma | g.b_tma) != 0;
  /* the first S chunks */
  for (int c = 0; c < S; c++) {
    if (c < nchunks) {
      if (any_tma && tid == 64) tma_chunk(c, c * KC);
      if (!g.a_tma) load_tile(sbase + c * stage_bytes, a_at(kbeg + c * KC), g.a_rs, g.a_cs, BM, mvalid, klen - c * KC, a16);
      if (!g.b_tma) load_tile(sbase + c * stage_bytes + A_BYTES, b_at(kbeg + c * KC), g.b_ns, g.b_ks, g.n_tile, nvalid, klen - c * KC, b16);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  if (tid < g.n_tile) sbias[tid] = (g.bias && tid < nvalid) ? __ldg(g.bias + (long long)bz * g.bias_bs + n0 + tid) : 0.0f;     /* read in the epilogue, after the chunk loop's CTA barriers */
  CLK(2);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((g.a_tma == 2 ? 1u : 0u) << 15) | ((g.b_tma == 2 ? 1u : 0u) << 16) |      /* bits 15 / 16: A / B is MN-major */
                         ((uint32_t)(g.n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint64_t desc_hi = ((uint64_t)g.lbo16 << 16) | ((uint64_t)g.sbo16 << 32) | (1ull << 46);
  /* TMA-staged operand: K-major, 128-byte swizzle -- layout type 2 in bits [61,64), stride byte offset = 8 rows x 128 B = 1024, leading byte offset unused (1);
     the k-th product of a chunk reads atom k / 4 (rows x 128 B apart) at byte offset 32 (k % 4) inside the swizzled 128-byte rows */
  const uint64_t desc_hi_sw = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
  /* TMA-staged MN-major operand.  For 32-bit operands the tensor core accepts exactly one MN-major layout: 128-byte swizzle with 32-BYTE atoms (descriptor
     layout type 1; TMA swizzle CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a 128-byte row holds 32 consecutive rows / columns of the operand for ONE k, the four
     32-byte pieces of a row are XOR-ed with (k mod 4); the pattern repeats every 4 k.  Leading byte offset = distance between 32-wide atoms along the row
     index (64 k x 128 B = 8192), stride byte offset = distance between 4-k groups (512); the k-th product of a chunk (8 k) starts 1024 k bytes in. */
  const uint64_t desc_hi_mn = g_mn_swap ? (((uint64_t)(512 >> 4) << 16) | ((uint64_t)((KC * 128) >> 4) << 32) | (1ull << 46) | (1ull << 61))
                                        : (((uint64_t)((KC * 128) >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61));

  int st = 0, ph = 0;                                           /* stage and mbarrier phase of chunk i: i % S, (i / S) & 1 */
  for (int i = 0; i < nchunks; i++) {
    if (S == 4) asm volatile("cp.async.wait_group 3;\n" ::: "memory");      /* exactly S-1 groups are younger than chunk i's */
    else if (S == 3) asm volatile("cp.async.wait_group 2;\n" ::: "memory");
    else if (S == 2) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (i == 0) CLK(3);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int kmma = (min(KC, klen - i * KC) + 7) >> 3;
      const uint32_t a_addr = sbase + st * stage_bytes, b_addr = a_addr + A_BYTES;
      if (any_tma) mbar_wait(bar0 + FBAR_OFF + 8 * st, ph);     /* the TMA boxes of this chunk have landed (same stage / phase sequence as the "free" barriers) */
      for (int kk = 0; kk < kmma; kk++) {
        const uint64_t da = g.a_tma == 1 ? (desc_hi_sw | (uint64_t)(((a_addr + (kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4) & 0x3FFF))
                          : g.a_tma == 2 ? (desc_hi_mn | (uint64_t)(((a_addr + kk * 1024) >> 4) & 0x3FFF))
                                         : (desc_hi | (uint64_t)(((a_addr + kk * 256) >> 4) & 0x3FFF));
        const uint64_t db = g.b_tma == 1 ? (desc_hi_sw | (uint64_t)(((b_addr + (kk >> 2) * (g.n_tile * 128) + (kk & 3) * 32) >> 4) & 0x3FFF))
                          : g.b_tma == 2 ? (desc_hi_mn | (uint64_t)(((b_addr + kk * 1024) >> 4) & 0x3FFF))
                                         : (desc_hi | (uint64_t)(((b_addr + kk * 256) >> 4) & 0x3FFF));
        const uint32_t acc = (i > 0 || kk > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar0 + 8 * st) : "memory");
    }
    if (i + S < nchunks) {                                      /* refill this stage once its products have read it */
      const int k1 = (i + S) * KC;
      mbar_wait(bar0 + 8 * st, ph);
      if (any_tma && tid == 64) tma_chunk(st, k1);
      if (!g.a_tma) load_tile(sbase + st * stage_bytes, a_at(kbeg + k1), g.a_rs, g.a_cs, BM, mvalid, klen - k1, a16);
      if (!g.b_tma) load_tile(sbase + st * stage_bytes + A_BYTES, b_at(kbeg + k1), g.b_ns, g.b_ks, g.n_tile, nvalid, klen - k1, b16);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (i + 1 < nchunks && ++st == S) { st = 0; ph ^= 1; }
  }
  CLK(4);
  const bool done = mbar_wait(bar0 + 8 * st, ph);               /* the last commit covers every product issued before it */
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  CLK(5);
  /* epilogue.  (1) every warp takes its share of the accumulator out of tensor memory (warp w reads lanes 32*(w%4).., one row of C per thread,
     `tcgen05.ld.32x32b.x16`; warps w and w+4 take alternate 16-column slabs) and sends each group of 4 columns to the CTA that owns them:
     CTA j of the cluster owns columns [j, j+1) * n_tile/splits of the tile and keeps one receive panel per sender (rows padded by 4 words).
     Without split-K the panel is the CTA's own and lies over the (now free) stages; with split-K the panels lie behind the stages, because a
     fast CTA may deliver while the owner's products still read its stages, and the stores go through distributed shared memory.
     The remote stores are `st.async` with mbarrier completion: the owner arms one receive barrier with the 512 x n_tile bytes it expects.
     (2) CTA barrier, or wait on the receive barrier (no cluster-wide release fence on the critical path).  (3) the owner sums its panels in sender order (deterministic) and runs bias / ReLU / mask / accumulate;
     a thread handles 4 consecutive columns of a row, so the lanes of a warp write contiguous pieces of C.  Nothing remote is touched after
     the barrier, so the CTAs of a cluster retire independently. */
  const int w = g.n_tile / g.splits, ldp = w + 4, wsh = 31 - __clz(w >> 2);       /* w/4 = 1 << wsh column groups per owner */
  float *recv = (float *)(smem + CTRL_BYTES + (g.splits > 1 ? g.recv_off : 0));
  if (g.splits > 1) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");    /* every CTA of the cluster runs, receive barriers armed */
  {
    const int q = warp & 3, half = warp >> 2, r = q * 32 + lane;
    const uint32_t mine = smem_u32(recv + (rank * BM + r) * ldp);
    for (int c0 = half * 16; c0 < g.n_tile; c0 += 32) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int c = c0 + j, owner = c >> (wsh + 2), cl = c & (w - 1);
        if (g.splits > 1) {                                     /* asynchronous remote store that reports its 16 bytes on the owner's receive barrier */
          uint32_t dst, dbar;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(dst) : "r"(mine + cl * 4), "r"(owner));
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(dbar) : "r"(rbar), "r"(owner));
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                       ::"r"(dst), "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]), "r"(dbar) : "memory");
        } else
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(mine + cl * 4), "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]) : "memory");
      }
    }
  }
  CLK(8);
  bool got = true;
  if (g.splits == 1) __syncthreads();
  else got = mbar_wait(rbar, 0);                                /* every partial sum addressed to this CTA has landed */
  CLK(9);
  {
    const bool relu = g.flags & RSB_GEMM_RELU, accum = g.flags & RSB_GEMM_ACCUMULATE;
    const int cbase = rank * w;                                 /* first tile column this CTA owns */
    float *cbase_p = g.c + (long long)bz * g.c_bs + n0;
    const float *mbase_p = g.mask ? g.mask + (long long)bz * g.mask_bs + n0 : nullptr;
    const bool st16 = ((g.c_rs & 3) == 0) && (((uintptr_t)cbase_p & 15) == 0);
    for (int t = tid; t < (BM << wsh); t += NTHREADS) {
      const int row = t >> wsh, cl = (t - (row << wsh)) * 4, c = cbase + cl, m = m0 + row;
      const float *lp = recv + row * ldp + cl;
      float4 x = *reinterpret_cast<const float4 *>(lp);
      for (int sidx = 1; sidx < g.splits; sidx++) {
        const float4 pv = *reinterpret_cast<const float4 *>(lp + sidx * BM * ldp);
        x.x += pv.x; x.y += pv.y; x.z += pv.z; x.w += pv.w;
      }
      if (m < g.m && c < nvalid) {
        float *crow = cbase_p + (long long)m * g.c_rs;
        const float *mrow = mbase_p ? mbase_p + (long long)m * g.mask_rs : nullptr;
        const float4 bv = *reinterpret_cast<const float4 *>(sbias + c);
        float f[4] = {x.x + bv.x, x.y + bv.y, x.z + bv.z, x.w + bv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int nn = c + j;
          if (nn < nvalid) {
            if (relu) f[j] = fmaxf(f[j], 0.0f);
            if (mrow) f[j] = (mrow[nn] > 0.0f) ? f[j] : 0.0f;
            if (accum) f[j] += crow[nn];
          }
        }
        if (!(done && got)) f[0] = f[1] = f[2] = f[3] = __int_as_float(0x7fc00000);      /* a bounded wait gave up: the tile is poisoned with NaN so the failure
                                                                                       reaches the losses / parameters instead of leaving last update's values in C */
        if (st16 && c + 4 <= nvalid) *reinterpret_cast<float4 *>(crow + c) = make_float4(f[0], f[1], f[2], f[3]);
        else {
#pragma unroll
          for (int j = 0; j < 4; j++) if (c + j < nvalid) crow[c + j] = f[j];
        }
      }
    }
  }
  CLK(10);
  CLK(6);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  CLK(7);
}

static int g_swap_offsets = 0, g_force_splits = 0;

/* ---- TMA descriptors.  An operand is TMA-eligible when its contraction index is contiguous in memory and every stride is a multiple of 16 bytes: then
   a [rows x 64] chunk is two `cp.async.bulk.tensor` boxes issued by ONE thread (the strided cp.async gather costs 16-64 instructions per thread and chunk:
   9.2 k of the 16.5 k cycles of a 128 x 256 x 256 product in round 1).  The map is 3-D (k, row, z) with z = batch or k-block index.  RSB_GEMM_TMA=0 disables. */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int g_tma_mode = -1, g_last_tma = 0;     /* diagnostic: -1 = as RSB_GEMM_TMA says (default on), 0 = force the cp.async gather, 1 = on; staging mode of the last launch: A in bits 0-1, B in bits 2-3 (0 gather, 1 TMA K-major, 2 TMA MN-major) */
extern "C" void rsb_gemm_debug_tma(int mode) { g_tma_mode = mode; }
extern "C" void rsb_gemm_debug_mn_swap(int swap) { cudaMemcpyToSymbol(g_mn_swap, &swap, sizeof(int)); }
extern "C" int rsb_gemm_debug_last_tma(void) { return g_last_tma; }
static EncodeTiledFn encode_tiled() {
  if (g_tma_mode == 0) return nullptr;
  static EncodeTiledFn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true;
    const char *ev = getenv("RSB_GEMM_TMA");
    if (!(ev && atoi(ev) == 0)) {
      void *p = nullptr; cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
  }
  return fn;
}
/* 3-D view of an operand: element (i, o, z) at base + z * zs + o * rs + i (floats), i = the index that is contiguous in memory (k for a K-major operand,
   the row / column index for an MN-major one), o = the other one; box = 32 i x box_rows o */
static bool make_map(CUtensorMap *map, const float *base, long k_ext, long rows, long nz, long rs, long zs, int box_rows, bool atom32 = false) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn || k_ext < 32 || (rs & 3) || (nz > 1 && (zs & 3)) || ((uintptr_t)base & 15) || rs < k_ext) return false;
  if (nz <= 1) { nz = 1; zs = rows * rs; if (zs & 3) zs = (zs + 3) & ~3L; }
  const cuuint64_t dims[3] = {(cuuint64_t)k_ext, (cuuint64_t)rows, (cuuint64_t)nz};
  const cuuint64_t strides[2] = {(cuuint64_t)rs * 4, (cuuint64_t)zs * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" void rsb_gemm_debug_swap_offsets(int swap) { g_swap_offsets = swap; }
extern "C" void rsb_gemm_debug_splits(int splits) { g_force_splits = splits; }

extern "C" int rsb_gemm_debug_clocks(long long *host_out12) {
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  return cudaMemcpyFromSymbol(host_out12, g_gemm_clk, 12 * sizeof(long long)) != cudaSuccess;
}

extern "C" int rsb_gemm_timeouts(void) {
  unsigned int h = 0, z = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&h, g_gemm_timeouts, sizeof(h)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_gemm_timeouts, &z, sizeof(z));
  return (int)h;
}

/* launch geometry of one product (pure host arithmetic; tests/test_abi.py checks its invariants without a GPU).
   plan = {n_tile, splits, chunks per split, stages, byte offset of the receive panels behind the stages, dynamic shared memory, grid.x, CTAs} */
extern "C" int rsb_gemm_plan(int m, int n, int k, int batch, int n_tile, int force_splits, int *plan) {
  if (m < 1 || n < 1 || k < 1 || batch < 1 || !plan) { rsb_sac_set_error("rsb_gemm_plan: bad arguments"); return 1; }
  const int mt = (m + BM - 1) / BM;
  if (n_tile == 0) {                                  /* widest tile that still gives >= 32 CTAs; small outputs: the narrowest tile that covers n */
    n_tile = 128;
    while (n_tile > 16 && ((n + n_tile - 1) / n_tile) * mt * batch < 32) n_tile >>= 1;
    while (n_tile > 16 && (n_tile >> 1) >= n) n_tile >>= 1;
  }
  if (n_tile != 16 && n_tile != 32 && n_tile != 64 && n_tile != 128) { rsb_sac_set_error("rsb_gemm_tf32: n_tile must be 16, 32, 64 or 128"); return 1; }
  /* split the contraction over a cluster of 2 or 4 CTAs while the launch still fits one wave: at the update's sizes the time of a product is
     the time one SM needs to pull its operands in, so more SMs with less each is what helps */
  const int ctas = ((n + n_tile - 1) / n_tile) * mt * batch, nchunks_all = (k + KC - 1) / KC;
  int splits = force_splits;
  static int max_ctas = 0;                                       /* developer knob RSB_GEMM_MAX_CTAS: CTAs one product may occupy */
  if (max_ctas == 0) { const char *ev = getenv("RSB_GEMM_MAX_CTAS"); max_ctas = ev ? atoi(ev) : 148; if (max_ctas < 1) max_ctas = 148; }
  if (splits == 0) { splits = 4; while (splits > 1 && (splits > nchunks_all || ctas * splits > max_ctas)) splits >>= 1; }
  if (splits != 1 && splits != 2 && splits != 4) { rsb_sac_set_error("rsb_gemm_tf32: splits must be 1, 2 or 4"); return 1; }
  int cps = (nchunks_all + splits - 1) / splits;
  while (splits > 1 && (splits - 1) * cps >= nchunks_all) { splits >>= 1; cps = (nchunks_all + splits - 1) / splits; }   /* no empty slice */
  /* shared memory: the stages this launch can fill (4 x (32 KB + n_tile x 256 B); 3 at n_tile = 128, 2 if that tile is also split) and, with
     split-K, the receive panels behind them; without split-K the single panel lies over the stages */
  int stages = (n_tile == 128) ? (splits > 1 ? 2 : 3) : 4;
  if (cps < stages) stages = cps;
  const size_t stage_bytes = A_BYTES + (size_t)n_tile * KC * 4, panels = (size_t)splits * BM * (n_tile / splits + 4) * 4;
  while (stages > 1 && CTRL_BYTES + (size_t)stages * stage_bytes + (splits > 1 ? panels : 0) > SMEM_MAX) stages--;
  size_t smem_bytes = CTRL_BYTES + (size_t)stages * stage_bytes + (splits > 1 ? panels : 0);
  if (smem_bytes < CTRL_BYTES + panels) smem_bytes = CTRL_BYTES + panels;
  if (smem_bytes > SMEM_MAX) { rsb_sac_set_error("rsb_gemm_plan: shared memory"); return 1; }
  plan[0] = n_tile; plan[1] = splits; plan[2] = cps; plan[3] = stages; plan[4] = (int)((size_t)stages * stage_bytes); plan[5] = (int)smem_bytes;
  plan[6] = ((n + n_tile - 1) / n_tile) * splits; plan[7] = ctas * splits;
  return 0;
}

extern "C" int rsb_gemm_tf32(const float *d_a, long a_rs, long a_cs, long a_bs, const float *d_b, long b_ks, long b_ns, long b_bs, float *d_c, long c_rs,
                             long c_bs, int m, int n, int k, int batch, const float *d_bias, long bias_bs, const float *d_mask, long mask_rs, long mask_bs,
                             int flags, int n_tile, int k_block, long a_kbs, long b_kbs, void *stream) {
  if (m < 1 || n < 1 || k < 1 || batch < 1 || !d_a || !d_b || !d_c) { rsb_sac_set_error("rsb_gemm_tf32: bad arguments"); return 1; }
  if (k_block <= 0) { k_block = ((k + KC - 1) / KC) * KC; a_kbs = b_kbs = 0; }
  if (k_block % KC != 0) { rsb_sac_set_error("rsb_gemm_tf32: k_block must be a multiple of 64"); return 1; }
  int plan[8];
  if (rsb_gemm_plan(m, n, k, batch, n_tile, g_force_splits, plan) != 0) return 1;
  n_tile = plan[0];
  const int splits = plan[1];
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(k_gemm_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    attr_set[dev] = (e == cudaSuccess);
  }
  if (e != cudaSuccess) { rsb_sac_set_error(cudaGetErrorString(e)); return 1; }
  GemmArgs g;
  g.a = d_a; g.b = d_b; g.bias = d_bias; g.mask = d_mask; g.c = d_c;
  g.a_rs = a_rs; g.a_cs = a_cs; g.a_bs = a_bs; g.b_ks = b_ks; g.b_ns = b_ns; g.b_bs = b_bs; g.c_rs = c_rs; g.c_bs = c_bs;
  g.bias_bs = bias_bs; g.mask_rs = mask_rs; g.mask_bs = mask_bs;
  g.m = m; g.n = n; g.k = k; g.flags = flags; g.n_tile = n_tile; g.k_block = k_block; g.a_kbs = a_kbs; g.b_kbs = b_kbs;
  g.lbo16 = g_swap_offsets ? (2048 >> 4) : (128 >> 4);
  g.sbo16 = g_swap_offsets ? (128 >> 4) : (2048 >> 4);
  g.splits = splits; g.cps = plan[2]; g.stages = plan[3]; g.recv_off = plan[4];
  /* TMA staging where the operand allows it (K-contiguous, 16-byte strides); the third coordinate is the k-block (stack_k products) or the batch */
  const int nblk = (k + k_block - 1) / k_block;
  alignas(64) CUtensorMap ta, tb; memset(&ta, 0, sizeof ta); memset(&tb, 0, sizeof tb);
  g.z_is_block = nblk > 1; g.a_tma = g.b_tma = 0;
  g.a_zbcast = (nblk == 1 && batch > 1 && a_bs == 0); g.b_zbcast = (nblk == 1 && batch > 1 && b_bs == 0);
  if (!(nblk > 1 && batch > 1)) {
    const long k_ext = nblk > 1 ? k_block : k, nza = g.a_zbcast ? 1 : (nblk > 1 ? nblk : batch), nzb = g.b_zbcast ? 1 : (nblk > 1 ? nblk : batch);
    if (a_cs == 1) g.a_tma = make_map(&ta, d_a, k_ext, m, nza, a_rs, nblk > 1 ? a_kbs : a_bs, BM) ? 1 : 0;
    if (!g.a_tma && a_rs == 1 && nblk == 1) g.a_tma = make_map(&ta, d_a, m, k, nza, a_cs, a_bs, KC, true) ? 2 : 0;                         /* A^T in memory */
    if (b_ks == 1) g.b_tma = make_map(&tb, d_b, k_ext, n, nzb, b_ns, nblk > 1 ? b_kbs : b_bs, n_tile) ? 1 : 0;
    if (!g.b_tma && b_ns == 1 && nblk == 1 && n_tile >= 32) g.b_tma = make_map(&tb, d_b, n, k, nzb, b_ks, b_bs, KC, true) ? 2 : 0;         /* B stored [k, n] */
  }
  g_last_tma = g.a_tma | (g.b_tma << 2);
  const size_t smem_bytes = (size_t)plan[5];
  dim3 grid(plan[6], (m + BM - 1) / BM, batch);
  e = rsb_launch_pdl(k_gemm_tf32, grid, dim3(NTHREADS, 1, 1), smem_bytes, (cudaStream_t)stream, splits, g, ta, tb);
  if (e != cudaSuccess) { rsb_sac_set_error(cudaGetErrorString(e)); return 1; }
  return 0;
}
