// Shared host-side plumbing for libmoe_b200.so: error reporting, launch accounting, small helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>
#include <utility>

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

// Programmatic dependent launch: every kernel of this library is launched with the
// programmaticStreamSerialization attribute, runs its prologue while its predecessor in the stream drains,
// and executes pdl_wait() before it reads or writes global memory the predecessor may touch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MOE_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// <<<grid, block, smem, stream>>> with the PDL attribute
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
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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
