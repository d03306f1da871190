/*
 * rsb_gemm.h -- C-ABI of the tensor-core GEMM used by the SAC update (part of librsb_cuda.so; kernel in csrc/rsb_tc_gemm.cu).
 *
 * Reference interface replaced: the `torch.nn.Linear` forward/backward products inside rlkit's FlattenMlp / TanhGaussianPolicy as
 * driven by SACTrainer.train_from_torch (un-vendored rlkit @ b7f97b2, reached from util/rlkit_custom.py:238).
 *
 *   C[b][m,n] = epilogue( sum_k A[b][m,k] * B[b][k,n] ),   b < batch
 *
 * A, B are fp32 in HBM with arbitrary ELEMENT strides (so X, X^T, W, W^T need no copies); the products run on the sm_100a tensor
 * cores as TF32 x TF32 -> FP32 (`tcgen05.mma.kind::tf32`, accumulator in tensor memory).  C is row-major with row stride c_rs.
 * epilogue: v += bias[b][n] (if bias); v = max(v, 0) (flag RELU); v = mask[b][m,n] > 0 ? v : 0 (if mask: the ReLU backward of a layer whose
 * output is `mask`); v += C[b][m,n] (flag ACCUMULATE).  All pointers are DEVICE pointers; `stream` is a cudaStream_t as void*; the call
 * is stream-ordered, non-blocking and CUDA-graph capturable.  Returns 0, or non-zero with rsb_sac_last_error() set.
 */
#ifndef RSB_GEMM_H
#define RSB_GEMM_H
#ifdef __cplusplus
extern "C" {
#endif
#define RSB_GEMM_RELU 1
#define RSB_GEMM_ACCUMULATE 2
/* n_tile: 0 = choose (16/32/64/128 output columns per CTA; one CTA owns a 128 x n_tile tile of C).
   k_block > 0 (a multiple of 64): the contraction index runs over k / k_block blocks that lie a_kbs / b_kbs elements apart (A[m, k] at
   a + (k / k_block) * a_kbs + m * a_rs + (k % k_block) * a_cs, likewise B) -- e.g. the twin Q networks' activation gradients side by side, so that
   dX = dH_1 W_1^T + dH_2 W_2^T is ONE product; k_block = 0: plain strides. */
int rsb_gemm_tf32(const float *d_a, long a_rs, long a_cs, long a_bs, const float *d_b, long b_ks, long b_ns, long b_bs, float *d_c, long c_rs, long c_bs,
                  int m, int n, int k, int batch, const float *d_bias, long bias_bs, const float *d_mask, long mask_rs, long mask_bs, int flags, int n_tile,
                  int k_block, long a_kbs, long b_kbs, void *stream);
/* launch geometry rsb_gemm_tf32 would use (host arithmetic only, no device needed): plan[8] = {n_tile, splits (CTAs of a cluster sharing a tile along K),
   64-wide chunks per split, pipeline stages, byte offset of the split-K receive panels, dynamic shared memory bytes, grid.x, total CTAs} */
int rsb_gemm_plan(int m, int n, int k, int batch, int n_tile, int force_splits, int *plan);
/* device-side watchdog: number of mbarrier waits that gave up since the last call (0 on a healthy run); synchronises the device */
int rsb_gemm_timeouts(void);
/* diagnostic: exchange the two byte-offset fields of the shared-memory matrix descriptors (0 = as documented in csrc/rsb_tc_gemm.cu) */
void rsb_gemm_debug_swap_offsets(int swap);
/* diagnostic: operand staging.  An operand with one unit stride and every other stride a multiple of 16 bytes is staged by TMA (`cp.async.bulk.tensor`,
   128-byte swizzle, one issuing thread): K-major when the contraction index is the contiguous one (activations, W^T), MN-major when the row / column index
   is (X^T, weights stored [in, out]; needs a tile of >= 32 columns for B); anything else by the strided `cp.async` gather.  mode 0 forces the gather for
   every operand, 1 / -1 = TMA where eligible (RSB_GEMM_TMA=0 in the environment also disables it); rsb_gemm_debug_last_tma: staging of the last launch,
   A in bits 0-1, B in bits 2-3 (0 gather, 1 TMA K-major, 2 TMA MN-major).  All paths feed the same products in the same order: results are bit-identical
   (tests/test_gpu_tc_gemm.py). */
void rsb_gemm_debug_tma(int mode);
void rsb_gemm_debug_mn_swap(int swap);    /* diagnostic: exchange the leading / stride byte offsets of the MN-major swizzled descriptor (0 = as documented in the kernel) */
int rsb_gemm_debug_last_tma(void);
/* diagnostic: force the number of CTAs (1, 2, 4) that share one tile of C along the contraction; 0 = choose */
void rsb_gemm_debug_splits(int splits);
/* diagnostic: SM clock of CTA (0,0,0) of the last launch at 12 points ([0] entry, [1] tensor memory + barriers ready and the preceding kernel complete, [2] first copies issued, [3] first chunk landed,
   [4] products issued, [5] accumulator complete, [8] accumulator parked in shared memory, [9] CTA/cluster barrier passed, [10] C written, [6] cluster
   released, [7] exit; [11] unused); synchronises the device */
int rsb_gemm_debug_clocks(long long *host_out12);
#ifdef __cplusplus
}
#endif
#endif
