/*
 * rsb_pdl.h -- programmatic dependent launch (sm_90+) for the SAC update's kernel chain.
 *
 * The update at the reference's batch size is a chain of ~19 dependent launches of 2-6 us each: what bounds it is the
 * launch-to-launch latency, not arithmetic.  Every kernel of the chain is launched with cudaLaunchAttributeProgrammaticStreamSerialization and
 * has the same shape:
 *
 *     prologue that touches NO data of an earlier kernel (index arithmetic, tensor-memory allocation, mbarrier init)
 *     pdl_wait();       griddepcontrol.wait: returns once the preceding kernel of the stream has COMPLETED and its writes are visible
 *     pdl_trigger();    griddepcontrol.launch_dependents: the next kernel of the stream may now be scheduled; it runs its own prologue and
 *                       parks in its pdl_wait() while this kernel does its work
 *     ... loads, math, stores ...
 *
 * Triggering only AFTER the wait keeps the overlap one kernel deep: when kernel n starts, kernel n-2 and everything before it has completed,
 * so ordering stays transitive along the stream exactly as with ordinary launches; nothing is read before the wait.  Inside a captured CUDA
 * graph the attribute becomes a programmatic edge between the two kernel nodes.  RSB_PDL=0 launches everything the ordinary way.
 */
#ifndef RSB_PDL_H
#define RSB_PDL_H
#include <cuda_runtime.h>
#include <stdlib.h>

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

inline bool rsb_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("RSB_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on != 0;
}

/* launch `kernel` on `stream` with the programmatic-serialization attribute (and, optionally, a cluster shape) */
template <typename... KArgs, typename... Args>
inline cudaError_t rsb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2]; int na = 0;
  if (cluster_x > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cluster_x; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
  if (rsb_pdl_enabled()) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
