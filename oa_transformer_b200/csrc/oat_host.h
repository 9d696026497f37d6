// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/oat.h"

namespace oat {

// thread-local last-error string, surfaced through oat_last_error()
char* last_error_buffer();
int set_error(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(OAT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return OAT_OK;
}

inline cudaStream_t as_stream(oat_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();

// OAT_PDL=1 turns programmatic dependent launch on (default: plain stream order; measured no gain, see common.cu).
bool pdl_enabled();

// Launch `kern` so that it may overlap its prologue with the tail of the previous kernel in the stream (the kernel must
// call pdl_wait() before its first global access - oat_ptx.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace oat

#define OAT_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return oat::set_error(OAT_ERR_ARG, __VA_ARGS__);     \
  } while (0)
