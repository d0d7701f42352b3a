// One fused AdamW step over a flat fp32 parameter range (HBM-bound: 28 B read + 14 B written per parameter) with the
// transformers==4.48.0 AdamW rule, plus the bf16 shadow refresh the GEMMs read.  Also the flat fp32->bf16 cast.
#include "common.cuh"

namespace vb {

template <bool G16>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const void* __restrict__ gv, float* __restrict__ m, float* __restrict__ v, bf16* __restrict__ shadow, long long n,
             float step_size, float lr_wd, float beta1, float beta2, float ob1, float ob2, float eps, float grad_scale,
             const float* __restrict__ sched_dev) {
  pdl_enter();
  if (sched_dev) {  // {step_size, lr*weight_decay} read from device memory: a captured CUDA graph can follow an lr schedule
    step_size = sched_dev[0];
    lr_wd = sched_dev[1];
  }
  const long long n4 = n >> 2;
  const void* gvp = gv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv;
    if constexpr (G16) {  // gradients arrive as bf16 (data-parallel all-reduce ran on a bf16 copy)
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(gvp) + i);
      const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
      gv = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gv = __ldg(reinterpret_cast<const float4*>(gvp) + i);
    }
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    float* gp = reinterpret_cast<float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = gp[e] * grad_scale;
      mp[e] = beta1 * mp[e] + ob1 * gg;
      vp[e] = beta2 * vp[e] + ob2 * gg * gg;
      const float denom = sqrtf(vp[e]) + eps;
      pp[e] = pp[e] - step_size * (mp[e] / denom);
      pp[e] = pp[e] - lr_wd * pp[e];
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float gg = (G16 ? __bfloat162float(reinterpret_cast<const bf16*>(gvp)[i]) : reinterpret_cast<const float*>(gvp)[i]) * grad_scale;
    const float mm = beta1 * m[i] + ob1 * gg;
    const float vv = beta2 * v[i] + ob2 * gg * gg;
    float pp = p[i] - step_size * (mm / (sqrtf(vv) + eps));
    pp -= lr_wd * pp;
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (shadow) shadow[i] = __float2bfloat16(pp);
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_enter();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + i);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}

static unsigned flat_grid(long long n4) {
  long long g = (n4 + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace vb

using namespace vb;

extern "C" int vault_adamw_step(float* p, const void* g, int32_t grad_is_bf16, float* m, float* v, void* shadow_bf16, int64_t n, double lr,
                                double beta1, double beta2, double eps, double weight_decay, int32_t correct_bias, int32_t step, float grad_scale,
                                const float* sched_dev, void* stream) {
  VB_REQUIRE(p && g && m && v && n >= 0, "adamw_step: bad arguments");
  VB_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)g & (grad_is_bf16 ? 7 : 15)) == 0,
             "adamw_step: buffers must be 16-byte aligned (bf16 gradients: 8)");
  VB_REQUIRE(shadow_bf16 == nullptr || ((uintptr_t)shadow_bf16 & 7) == 0, "adamw_step: shadow must be 8-byte aligned");
  if (n == 0) return VAULT_OK;
  double step_size = lr;  // hyper-parameters arrive as doubles: 1-beta is formed in double like torch's Python scalars, then rounded once
  if (correct_bias) {
    VB_REQUIRE(step >= 1, "adamw_step: step must be >= 1 with correct_bias");
    step_size = lr * sqrt(1.0 - pow(beta2, step)) / (1.0 - pow(beta1, step));
  }
  if (grad_is_bf16) {
    launch(adamw_kernel<true>, dim3(flat_grid(n >> 2)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, reinterpret_cast<bf16*>(shadow_bf16), n, (float)step_size,
                                                                    weight_decay > 0.0 ? (float)(lr * weight_decay) : 0.f, (float)beta1,
                                                                    (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, grad_scale, sched_dev);
    } else {
    launch(adamw_kernel<false>, dim3(flat_grid(n >> 2)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, reinterpret_cast<bf16*>(shadow_bf16), n, (float)step_size,
                                                                    weight_decay > 0.0 ? (float)(lr * weight_decay) : 0.f, (float)beta1,
                                                                    (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, grad_scale, sched_dev);
    }
  return check_launch("adamw_kernel");
}

extern "C" int vault_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  VB_REQUIRE(src && dst_bf16 && n >= 0, "cast_f32_bf16: bad arguments");
  VB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 7) == 0, "cast_f32_bf16: misaligned buffers");
  if (n == 0) return VAULT_OK;
  launch(cast_kernel, dim3(flat_grid(n >> 2)), dim3(256), 0, (cudaStream_t)stream, src, reinterpret_cast<bf16*>(dst_bf16), n);
  return check_launch("cast_kernel");
}
