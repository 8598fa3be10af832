// libmoe_b200.so -- library-wide plumbing: error channel, launch counter, TMA tensor-map encoding.
#include "common.cuh"
#include "tcgen05.cuh"

#include <cudaTypedefs.h>
#include <string.h>

namespace moe {

std::atomic<long long> g_launches{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols, bool swizzle128) {
  return make_tmap_bf16_2d_sw(out, base, rows, cols, box_rows, box_cols, swizzle128 ? 128 : 0);
}

int make_tmap_bf16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (cols * 2) % 16 != 0)
    return fail(MOE_ERR_INVALID_ARGUMENT, "TMA needs 16-byte aligned base and row pitch (cols=%llu)",
                (unsigned long long)cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                        : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MOE_OK;
}

int make_tmap_bf16_kblocks(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                           uint32_t box_kblocks) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || cols % 64 != 0)
    return fail(MOE_ERR_INVALID_ARGUMENT, "k-block tensor map needs a 16-byte aligned base and cols %% 64 == 0 (cols=%llu)",
                (unsigned long long)cols);
  cuuint64_t gdim[3] = {64, rows, cols / 64};
  cuuint64_t gstr[2] = {cols * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return MOE_OK;
}

}  // namespace moe

extern "C" {

int moe_abi_version(void) { return 3; }
const char* moe_last_error(void) { return moe::last_error_buf(); }
long long moe_launch_count(void) { return moe::g_launches.load(); }
void moe_reset_launch_count(void) { moe::g_launches.store(0); }

}  // extern "C"
