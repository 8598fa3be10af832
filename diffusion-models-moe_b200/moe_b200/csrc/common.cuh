// Shared host-side plumbing for libmoe_b200.so: error reporting, launch accounting, small helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../../include/moe_b200.h"

namespace moe {

char* last_error_buf();                    // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);  // formats into last_error_buf, returns code
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MOE_OK;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

}  // namespace moe

#define MOE_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return moe::fail(code, __VA_ARGS__); \
  } while (0)
