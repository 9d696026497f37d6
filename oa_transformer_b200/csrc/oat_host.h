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

}  // namespace oat

#define OAT_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return oat::set_error(OAT_ERR_ARG, __VA_ARGS__);     \
  } while (0)
