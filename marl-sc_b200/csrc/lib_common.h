// lib_common.h - error reporting and launch accounting shared by the translation units of the library.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "../../include/marlsc_b200.h"

namespace marlsc {
constexpr int kMaxDevices = 64;   // per-device caches of kernel attributes
extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;
int set_error(int code, const std::string& msg);
}  // namespace marlsc

#define MARLSC_CUDA(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return marlsc::set_error(MARLSC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
