// Data-parallel optimizer step as ONE kernel over NVSwitch multicast: gradient reduce-scatter (in the switch) + AdamW + weight all-gather.
//
// Every rank owns 1/world of each finished gradient range.  For its slice it
//   1. pulls the SUM of all ranks' fp32 gradients with multimem.ld_reduce.add.f32 (the switch fetches the 16 bytes from every replica's HBM
//      and adds them: one NVLink load returns the reduced value, no rank ever sees another rank's partial gradients),
//   2. applies the transformers==4.48.0 AdamW rule (same arithmetic as adamw.cu) with ITS OWN m / v -- the optimizer state of a slice is
//      only ever touched by its owner, so the 30 B/parameter of AdamW HBM traffic become 30/world + 10,
//   3. pushes the new bf16 shadow (what the next forward's GEMMs read) to EVERY replica (itself included) with multimem.st, and the new
//      fp32 masters too -- except for the 64-parameter blocks flagged in `local_only_bits`: the dense projection matrices (86 % of the
//      model) are only ever read through their shadow, so their masters stay SHARDED (a rank's copy is current for the slices it owns;
//      vault_mc_broadcast_f32 brings all replicas up to date when somebody wants to read the Parameters: evaluation, checkpoint), while
//      biases, LayerNorm affines and embedding tables, which kernels read in fp32, are replicated at once.
// Gradients travel as fp32 (bit-faithful sum) or as a bf16 copy (half the NVLink bytes; the switch accumulates in fp32).  Link bytes per
// parameter and GPU: up 2-4 (its gradients, read once by the switch) + 2/world, down 2-4/world + 2 -- the NVLS all-reduce's, while the
// AdamW traffic on HBM drops by (world-1)/world.
// The flat gradient, master and shadow buffers are symmetric memory (torch.distributed._symmetric_memory: empty + rendezvous ->
// multicast_ptr); replicas cannot drift because each parameter has exactly one writer.  Cross-rank ordering is the caller's: a barrier
// before the launch (every rank's gradients of the range are final) and one at the end of the step (every slice has been written
// everywhere; gradients may be overwritten) -- VaultTrainStep uses the symmetric-memory handle's stream-ordered barrier, which times out
// instead of hanging.  Replaces NCCL all-reduce + cast + a full-range AdamW launch per range
// (ref:vault/tmsc_utils/trainer.py:244-254 optimizer, SURVEY.md section 8e data parallelism).
#include "common.cuh"

namespace vb {
namespace {

__device__ __forceinline__ float4 mm_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ uint4 mm_ld_reduce_bf16x8(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32x4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mm_st_bf16x8(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct McAdamParams {
  float* p_local;        // this rank's copy of the masters
  float* p_mc;           // multicast address of the same range (write side)
  const uint32_t* local_only_bits;  // bit b = 1: the masters of 64-parameter block b (counted from parameter 0 of the flat buffer) stay local
  long long first_param; // flat index of this slice's first parameter
  const void* g_mc;      // multicast address of the gradients (fp32, or the bf16 copy)
  float* m;
  float* v;
  void* shadow_mc;       // multicast address of the bf16 shadow
  long long n8;          // units of 8 parameters
  float step_size, lr_wd, beta1, beta2, ob1, ob2, eps, grad_scale;
  const float* sched_dev;
};

__device__ __forceinline__ void adam4(float4& p, const float4& g, float4& m, float4& v, const McAdamParams& a, float step_size, float lr_wd) {
  float* pp = reinterpret_cast<float*>(&p);
  const float* gp = reinterpret_cast<const float*>(&g);
  float* mp = reinterpret_cast<float*>(&m);
  float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) {  // identical operation order to adamw_kernel: a 1-GPU and an N-GPU run differ only by the gradient sum
    const float gg = gp[e] * a.grad_scale;
    mp[e] = a.beta1 * mp[e] + a.ob1 * gg;
    vp[e] = a.beta2 * vp[e] + a.ob2 * gg * gg;
    const float denom = sqrtf(vp[e]) + a.eps;
    pp[e] = pp[e] - step_size * (mp[e] / denom);
    pp[e] = pp[e] - lr_wd * pp[e];
  }
}

constexpr int kMcThreads = 512;
// Sizing: an NVLink round trip through the switch is microseconds, so the kernel is sized by bytes in flight -- every thread has 128 bytes of
// gradient pulls outstanding (4 fp32 units or 8 bf16 units of 8 parameters), 2 CTAs/SM x 512 threads = 128 KB per SM; ~30 SMs cover the
// bandwidth-latency product of the link and the rest of the GPU stays with the backward pass (grid bounded by the caller: `ctas`).

template <bool G16>
__global__ void __launch_bounds__(kMcThreads, 2) mc_adamw_kernel(const McAdamParams a) {
  constexpr int kMcUnroll = G16 ? 8 : 4;  // 128 bytes of gradient pulls in flight per thread either way
  pdl_enter();
  float step_size = a.step_size, lr_wd = a.lr_wd;
  if (a.sched_dev) {
    step_size = a.sched_dev[0];
    lr_wd = a.sched_dev[1];
  }
  const long long stride = (long long)gridDim.x * kMcThreads;
  for (long long i0 = (long long)blockIdx.x * kMcThreads + threadIdx.x; i0 < a.n8; i0 += stride * kMcUnroll) {
    float4 g[G16 ? 1 : kMcUnroll][2];
    uint4 g16[G16 ? kMcUnroll : 1];  // bf16 payload: the 8 gradients of a unit stay packed (4 registers) until they are used
#pragma unroll
    for (int u = 0; u < kMcUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < a.n8) {
        if constexpr (G16) {
          g16[u] = mm_ld_reduce_bf16x8(reinterpret_cast<const char*>(a.g_mc) + 16 * i);
        } else {
          g[u][0] = mm_ld_reduce_f32x4(reinterpret_cast<const float*>(a.g_mc) + 8 * i);
          g[u][1] = mm_ld_reduce_f32x4(reinterpret_cast<const float*>(a.g_mc) + 8 * i + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kMcUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < a.n8) {
        uint32_t pk[4];
        bool local_only = false;
        if (a.local_only_bits) {
          const long long blk = (a.first_param + 8 * i) >> 6;
          local_only = (__ldg(a.local_only_bits + (blk >> 5)) >> (blk & 31)) & 1u;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // one half (4 parameters) at a time: the moments of a unit never all live in registers together
          float4 p = *reinterpret_cast<const float4*>(a.p_local + 8 * i + 4 * h);
          float4 m = *reinterpret_cast<const float4*>(a.m + 8 * i + 4 * h);
          float4 v = *reinterpret_cast<const float4*>(a.v + 8 * i + 4 * h);
          float4 gh;
          if constexpr (G16) {
            const float2 lo = unpack_bf16x2(h == 0 ? g16[u].x : g16[u].z), hi = unpack_bf16x2(h == 0 ? g16[u].y : g16[u].w);
            gh = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            gh = g[u][h];
          }
          adam4(p, gh, m, v, a, step_size, lr_wd);
          *reinterpret_cast<float4*>(a.m + 8 * i + 4 * h) = m;
          *reinterpret_cast<float4*>(a.v + 8 * i + 4 * h) = v;
          if (local_only) *reinterpret_cast<float4*>(a.p_local + 8 * i + 4 * h) = p;
          else mm_st_f32x4(a.p_mc + 8 * i + 4 * h, p);
          pk[2 * h] = pack_bf16x2(p.x, p.y);
          pk[2 * h + 1] = pack_bf16x2(p.z, p.w);
        }
        mm_st_bf16x8(reinterpret_cast<char*>(a.shadow_mc) + 16 * i, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
    }
  }
  __threadfence_system();  // the multicast stores are performed before this kernel counts as complete for the barrier that follows it
}

// masters of a slice -> every replica (consolidation of sharded masters before the parameters are read)
__global__ void __launch_bounds__(kMcThreads) mc_broadcast_f32_kernel(const float* __restrict__ src, float* dst_mc, long long n4) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * kMcThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kMcThreads)
    mm_st_f32x4(dst_mc + 4 * i, __ldg(reinterpret_cast<const float4*>(src) + i));
  __threadfence_system();
}

}  // namespace
}  // namespace vb

using namespace vb;

extern "C" int vault_mc_adamw_step(float* p_local, float* p_mc, const uint32_t* local_only_bits, int64_t first_param, const void* g_mc,
                                   int32_t grad_is_bf16, float* m, float* v, void* shadow_mc, int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int32_t correct_bias,
                                   int32_t step, float grad_scale, const float* sched_dev, int32_t ctas, void* stream) {
  VB_REQUIRE(p_local && p_mc && g_mc && m && v && shadow_mc && n >= 0, "mc_adamw_step: null pointer");
  VB_REQUIRE(first_param >= 0 && first_param % 8 == 0, "mc_adamw_step: first_param must be a multiple of 8");
  VB_REQUIRE(n % 8 == 0, "mc_adamw_step: a slice must be a multiple of 8 parameters (n = %lld)", (long long)n);
  VB_REQUIRE((((uintptr_t)p_local | (uintptr_t)p_mc | (uintptr_t)g_mc | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow_mc) & 15) == 0,
             "mc_adamw_step: buffers must be 16-byte aligned");
  VB_REQUIRE(ctas > 0, "mc_adamw_step: ctas must be positive");
  if (n == 0) return VAULT_OK;
  double step_size = lr;
  if (correct_bias) {
    VB_REQUIRE(step >= 1, "mc_adamw_step: step must be >= 1 with correct_bias");
    step_size = lr * sqrt(1.0 - pow(beta2, step)) / (1.0 - pow(beta1, step));
  }
  McAdamParams a;
  a.p_local = p_local; a.p_mc = p_mc; a.local_only_bits = local_only_bits; a.first_param = first_param; a.g_mc = g_mc; a.m = m; a.v = v; a.shadow_mc = shadow_mc;
  a.n8 = n / 8;
  a.step_size = (float)step_size;
  a.lr_wd = weight_decay > 0.0 ? (float)(lr * weight_decay) : 0.f;
  a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.ob1 = (float)(1.0 - beta1); a.ob2 = (float)(1.0 - beta2);
  a.eps = (float)eps; a.grad_scale = grad_scale; a.sched_dev = sched_dev;
  const int unroll = grad_is_bf16 ? 8 : 4;
  const long long need = (a.n8 + (long long)kMcThreads * unroll - 1) / ((long long)kMcThreads * unroll);
  const dim3 grid((unsigned)(need < ctas ? need : ctas)), block(kMcThreads);
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_is_bf16) launch(mc_adamw_kernel<true>, grid, block, 0, st, a);
  else launch(mc_adamw_kernel<false>, grid, block, 0, st, a);
  return check_launch("mc_adamw_kernel");
}

extern "C" int vault_mc_broadcast_f32(const float* src_local, float* dst_mc, int64_t n, int32_t ctas, void* stream) {
  VB_REQUIRE(src_local && dst_mc && n >= 0 && n % 4 == 0 && ctas > 0, "mc_broadcast_f32: bad arguments");
  VB_REQUIRE((((uintptr_t)src_local | (uintptr_t)dst_mc) & 15) == 0, "mc_broadcast_f32: buffers must be 16-byte aligned");
  if (n == 0) return VAULT_OK;
  const long long n4 = n / 4, need = (n4 + kMcThreads - 1) / kMcThreads;
  launch(mc_broadcast_f32_kernel, dim3((unsigned)(need < ctas ? need : ctas)), dim3(kMcThreads), 0, (cudaStream_t)stream, src_local, dst_mc, n4);
  return check_launch("mc_broadcast_f32_kernel");
}
