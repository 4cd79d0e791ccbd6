// Host-side helpers shared between the translation units of libsimba_b200.so (not part of the ABI).
#pragma once
#include "../../include/simba_b200.h"

namespace simba {
// records the message behind simba_last_error() and returns `code`
int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const simba_model_config_t* model_config(const simba_model_t* m);
}   // namespace simba

#define SIMBA_CUDA_TRY(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return simba::set_error(SIMBA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)
