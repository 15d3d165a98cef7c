// vine_launch.cuh — programmatic dependent launch (PDL) for the chains of short kernels of the PPO iteration.
//
// A PPO minibatch of the reference network is ~21 dependent launches of 4-50 us each; in a captured graph every
// kernel -> kernel edge costs the drain of the first grid plus the launch latency of the second.  Launching the second
// kernel with cudaLaunchAttributeProgrammaticStreamSerialization lets its grid be scheduled while the first one still runs;
// every kernel of the chain calls grid_dependency_sync() as its FIRST statement:
//   griddepcontrol.wait               blocks until the grids this launch depends on have completed and flushed their memory
//                                     (so the memory semantics are exactly those of ordinary stream order), then
//   griddepcontrol.launch_dependents  allows the NEXT kernel of the stream to be scheduled early in its turn.
// Both instructions are no-ops in a launch without the attribute; a kernel launched normally after one of these still waits
// for its full completion.  Off by default (plain stream order); vine_set_programmatic_launch(1) turns the attribute on for the
// launches that follow.  Measured: it pays for chains of few-microsecond kernels (the rollout at 4096 envs: -5 %), not for the
// update's 15-50 us kernels, whose early-scheduled successors only take SM slots.
#pragma once
#include <cuda_runtime.h>

namespace vine_launch {

inline int& programmatic_launch_enabled() {
  static int enabled = 0;
  return enabled;
}

__device__ __forceinline__ void grid_dependency_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename Kernel, typename... Args>
inline cudaError_t launch(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = programmatic_launch_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace vine_launch
