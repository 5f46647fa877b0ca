// tezip_b200 -- shared helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tezip_b200.h"

namespace tz {

void set_error(const char *fmt, ...);
void count_launch(long long n = 1);

#define TZ_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      tz::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TZ_ECUDA;                                                                        \
    }                                                                                         \
  } while (0)

#define TZ_CHECK_LAUNCH()                                                                     \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      tz::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TZ_ECUDA;                                                                        \
    }                                                                                         \
    tz::count_launch();                                                                       \
  } while (0)

#define TZ_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      tz::set_error(__VA_ARGS__);                                                             \
      return TZ_EINVAL;                                                                       \
    }                                                                                         \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// number of SMs of the current device (cached per device)
int sm_count();

}  // namespace tz
