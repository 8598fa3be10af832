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

constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

// SM count of the CURRENT device (a process may drive several GPUs: every cache below is per device)
inline int sm_count() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;  // B200
    n[dev] = v;
  }
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember (kernel, device) pairs
inline int ensure_dynamic_smem(const void* kfn, size_t bytes) {
  struct Entry { const void* fn; int dev; size_t bytes; };
  static Entry configured[256];
  static std::atomic<int> n_configured{0};
  const int dev = current_device();
  const int n = n_configured.load(std::memory_order_acquire);
  for (int i = 0; i < n; ++i)
    if (configured[i].fn == kfn && configured[i].dev == dev && configured[i].bytes >= bytes) return MOE_OK;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
  const int slot = n_configured.load(std::memory_order_relaxed);
  if (slot < 256) {      // (a race between two host threads costs one redundant cudaFuncSetAttribute, nothing else)
    configured[slot] = Entry{kfn, dev, bytes};
    n_configured.store(slot + 1, std::memory_order_release);
  }
  return MOE_OK;
}

}  // namespace moe

#define MOE_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return moe::fail(code, __VA_ARGS__); \
  } while (0)
