// The sampler step around the UNet call: classifier-free-guidance combine + DDIM (eta = 0)
// update in ONE streaming pass over the latents instead of the ~8 elementwise ATen kernels the stock pipeline issues
// per step [upstream StableDiffusionPipeline.__call__ / DDIMScheduler.step]:
//     eps    = eps_uncond + guidance * (eps_cond - eps_uncond)
//     x0     = (x - sqrt(1 - a_t) * eps) / sqrt(a_t)
//     x_prev = sqrt(a_prev) * x0 + sqrt(1 - a_prev) * eps
// HBM-bound (3 reads + 1 write per element), fp32 math, 16-byte accesses; latents are tiny (4 x 64 x 64 per sample), so
// the point is one launch per step rather than bandwidth.
#include "common.cuh"

namespace moe {

struct StepCoef {
  float guidance, sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, sqrt_1m_a_prev;
};

__device__ __forceinline__ float ddim_one(float eu, float ec, float x, const StepCoef& c) {
  const float eps = fmaf(c.guidance, ec - eu, eu);
  const float x0 = (x - c.sqrt_1m_a_t * eps) / c.sqrt_a_t;
  return fmaf(c.sqrt_a_prev, x0, c.sqrt_1m_a_prev * eps);
}

__global__ void __launch_bounds__(256) cfg_ddim_f32_kernel(const float* __restrict__ eu, const float* __restrict__ ec,
                                                           const float* __restrict__ x, float* __restrict__ out, long long n,
                                                           const StepCoef c, int vec_ok) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = vec_ok ? (n >> 2) : 0;
  for (long long i = tid; i < nvec; i += nthreads) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(eu) + i), b = __ldg(reinterpret_cast<const float4*>(ec) + i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<float4*>(out)[i] =
        make_float4(ddim_one(a.x, b.x, v.x, c), ddim_one(a.y, b.y, v.y, c), ddim_one(a.z, b.z, v.z, c), ddim_one(a.w, b.w, v.w, c));
  }
  for (long long i = (nvec << 2) + tid; i < n; i += nthreads) out[i] = ddim_one(eu[i], ec[i], x[i], c);
}

__global__ void __launch_bounds__(256) cfg_ddim_bf16_kernel(const __nv_bfloat16* __restrict__ eu, const __nv_bfloat16* __restrict__ ec,
                                                            const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            long long n, const StepCoef c, int vec_ok) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = vec_ok ? (n >> 3) : 0;      // 8 bf16 per 16 bytes
  for (long long i = tid; i < nvec; i += nthreads) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(eu) + i), b = __ldg(reinterpret_cast<const uint4*>(ec) + i);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __bfloat1622float2(pa[j]), fb = __bfloat1622float2(pb[j]), fv = __bfloat1622float2(pv[j]);
      po[j] = __floats2bfloat162_rn(ddim_one(fa.x, fb.x, fv.x, c), ddim_one(fa.y, fb.y, fv.y, c));
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
  for (long long i = (nvec << 3) + tid; i < n; i += nthreads)
    out[i] = __float2bfloat16(ddim_one(__bfloat162float(eu[i]), __bfloat162float(ec[i]), __bfloat162float(x[i]), c));
}

}  // namespace moe

extern "C" {

int moe_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_prev, long long n, int is_bf16,
                      float guidance, float alpha_t, float alpha_prev, void* stream) {
  using namespace moe;
  MOE_REQUIRE(n >= 0 && alpha_t > 0.f && alpha_t <= 1.f && alpha_prev >= 0.f && alpha_prev <= 1.f, MOE_ERR_INVALID_ARGUMENT,
              "moe_cfg_ddim_step: n=%lld alpha_t=%f alpha_prev=%f", n, alpha_t, alpha_prev);
  if (n == 0) return MOE_OK;
  MOE_REQUIRE(eps_uncond && eps_cond && x && x_prev, MOE_ERR_INVALID_ARGUMENT, "moe_cfg_ddim_step: NULL pointer");
  StepCoef c;
  c.guidance = guidance;
  c.sqrt_a_t = sqrtf(alpha_t);
  c.sqrt_1m_a_t = sqrtf(1.f - alpha_t);
  c.sqrt_a_prev = sqrtf(alpha_prev);
  c.sqrt_1m_a_prev = sqrtf(1.f - alpha_prev);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(eps_uncond) | reinterpret_cast<uintptr_t>(eps_cond) | reinterpret_cast<uintptr_t>(x) |
                       reinterpret_cast<uintptr_t>(x_prev)) & 15) == 0;
  const long long per_cta = 256LL * (is_bf16 ? 8 : 4);
  long long ctas = (n + per_cta - 1) / per_cta;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (ctas > cap) ctas = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t le;
  if (is_bf16)
    le = launch_pdl(cfg_ddim_bf16_kernel, dim3(static_cast<unsigned>(ctas)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(eps_uncond),
                    static_cast<const __nv_bfloat16*>(eps_cond), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(x_prev),
                    n, c, vec_ok);
  else
    le = launch_pdl(cfg_ddim_f32_kernel, dim3(static_cast<unsigned>(ctas)), dim3(256), 0, st, static_cast<const float*>(eps_uncond),
                    static_cast<const float*>(eps_cond), static_cast<const float*>(x), static_cast<float*>(x_prev), n, c, vec_ok);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_cfg_ddim_step launch: %s", cudaGetErrorString(le));
  return check_launch("moe_cfg_ddim_step");
}

}  // extern "C"
