// Kernel launch with the programmatic-stream-serialization attribute (programmatic dependent launch): consecutive
// kernels of a chain overlap the launch latency and prologue of kernel i+1 with the tail of kernel i. The kernels call
// pdl_launch_dependents() / pdl_wait() (common.cuh). g_opt_pdl = 0 (gh_set_option("pdl", 0)) launches them plainly.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace gh {

static int g_opt_pdl = 1;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace gh
