// vault_b200 -- shared device/host helpers for the sm_100a kernels.
// Everything here is hand-written PTX wrappers (mbarrier, TMA, tcgen05/TMEM), Philox, and the C-ABI error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#include "../../include/vault_b200.h"

namespace vb {

// ---- host-side error plumbing (abi.cu) ------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);  // records message for vault_last_error(), returns code
int check_launch(const char* what);        // cudaGetLastError() -> VAULT_ERR_LAUNCH
int device_sm_count();

#define VB_REQUIRE(cond, ...)                                      \
  do {                                                             \
    if (!(cond)) return ::vb::fail(VAULT_ERR_INVALID, __VA_ARGS__); \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with
// griddepcontrol.wait (before its first global-memory access) + griddepcontrol.launch_dependents: the NEXT kernel's CTAs may be
// scheduled and run their prologue (smem carve-up, barrier init, TMEM alloc, descriptor prefetch) while this one drains, and
// block in their own griddepcontrol.wait until this grid has completed and flushed.  ~600 dependent launches per training step
// make the launch/drain gap matter.  VAULT_B200_PDL=0 disables the attribute (the device instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VAULT_B200_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
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
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);  // errors surface through check_launch()
}

// same with a thread-block cluster of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline void launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- small device utilities -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx: 2 ulp, exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU of HF BertIntermediate / ViltIntermediate ("gelu") = x * Phi(x), with Phi from Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7):  1 - Phi(|x|) = 0.5 * P(t) * exp(-x^2/2),  t = 1 / (1 + p |x| / sqrt 2),  P = t (a1 + t (a2 + ... a5 t)).
// The GELU epilogues share the SM's issue slots with a saturated tensor pipe, so they are written for instruction count:
//   gelu(x)  = relu(x) - |x| * h            h = 0.5 P(t) exp(-x^2/2)          (13 instructions)
//   gelu'(x) = (x >= 0 ? 1 - h : h) + x * phi(x),   phi(x) = exp(-x^2/2) / sqrt(2 pi)   (one shared exponential)
__device__ __forceinline__ float gelu_tail_h(float ax, float e) {
  const float t = fast_rcp(fmaf(0.23164189f /* 0.3275911 / sqrt 2 */, ax, 1.0f));
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);  // 0.5 * a5, 0.5 * a4
  poly = fmaf(poly, t, 0.7107068705f);                  // 0.5 * a3
  poly = fmaf(poly, t, -0.142248368f);                  // 0.5 * a2
  poly = fmaf(poly, t, 0.127414796f);                   // 0.5 * a1
  return poly * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float e = fast_exp2(x * x * -0.72134752044448170368f);  // exp(-x^2/2)
  const float ax = fabsf(x);
  return fmaf(-ax, gelu_tail_h(ax, e), fmaxf(x, 0.0f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float e = fast_exp2(x * x * -0.72134752044448170368f);
  const float h = gelu_tail_h(fabsf(x), e);
  const float cdf = x >= 0.0f ? 1.0f - h : h;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// value and derivative together (training forward): one exponential, one tail polynomial
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& d) {
  const float e = fast_exp2(x * x * -0.72134752044448170368f);
  const float ax = fabsf(x);
  const float h = gelu_tail_h(ax, e);
  g = fmaf(-ax, h, fmaxf(x, 0.0f));
  const float cdf = x >= 0.0f ? 1.0f - h : h;
  d = fmaf(x * 0.39894228040143267794f, e, cdf);
}

// Two elements per instruction: Blackwell's packed fp32 pipe (fma/mul/add.f32x2 -> FFMA2 / FMUL2 / FADD2, IEEE rn per lane, so the results
// are bit-identical to the scalar forms above).  The GELU epilogues are bound by issue slots and latency next to the MMA warps; the
// polynomial part of a pair costs 10 instructions instead of 20 (MUFU, |x| and max stay per element).
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 gelu_tail_h2(float2 ax, float2 e) {
  const float2 d = fma2(splat2(0.23164189f), ax, splat2(1.0f));
  const float2 t = make_float2(fast_rcp(d.x), fast_rcp(d.y));
  float2 poly = fma2(splat2(0.5307027145f), t, splat2(-0.7265760135f));
  poly = fma2(poly, t, splat2(0.7107068705f));
  poly = fma2(poly, t, splat2(-0.142248368f));
  poly = fma2(poly, t, splat2(0.127414796f));
  return mul2(mul2(poly, t), e);
}
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 a = mul2(mul2(x, x), splat2(-0.72134752044448170368f));
  const float2 e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 h = gelu_tail_h2(ax, e);
  return fma2(make_float2(-ax.x, -ax.y), h, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}
__device__ __forceinline__ float2 gelu_erf_grad2(float2 x) {
  const float2 a = mul2(mul2(x, x), splat2(-0.72134752044448170368f));
  const float2 e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
  const float2 h = gelu_tail_h2(make_float2(fabsf(x.x), fabsf(x.y)), e);
  const float2 cdf = make_float2(x.x >= 0.0f ? 1.0f - h.x : h.x, x.y >= 0.0f ? 1.0f - h.y : h.y);
  return fma2(mul2(x, splat2(0.39894228040143267794f)), e, cdf);
}

// ---- Philox4x32-10 (counter-based dropout masks: forward and backward regenerate the same bits) ------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x5eedu;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// keep-mask for 4 consecutive elements starting at element index 4*q of dropout site `site`
// keep iff 32-bit draw >= p * 2^32
__device__ __forceinline__ uint4 dropout_bits4(uint64_t seed, uint32_t site, uint64_t q) { return Philox(seed)(q, site); }
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (mbarrier.try_wait may suspend the thread for a system-dependent time)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (surfacing as a launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("vault_b200: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---- TMA -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset (and signals the same-offset mbarrier) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
      "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_dst),
      "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole warp
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one thread: arrive on mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the same-offset mbarrier of every CTA in `mask` (smem slots filled by multicast are shared property)
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// warp-collective: 32 lanes x 32 consecutive fp32 columns; lane l of warp w reads TMEM lane 32*(w%4)+l
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, 128B swizzle (layout_type 2), descriptor version 1 (Blackwell).
//   K-major  tile (rows x 64 bf16, one 128B row per MN index): SBO = 1024 (8 rows), LBO unused.
//   MN-major tile (64 k-rows x 64 MN elements per 8 KB box, boxes stacked along MN): SBO = 1024 (8 k-rows),
//            LBO = bytes between consecutive 64-wide MN boxes.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Same for an MN-major operand of 4-byte elements (tf32): the swizzle works on 32-byte pieces (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
// descriptor layout type 1 "128B, base 32B").  Canonical atom = 4 k-rows x 128 B (32 MN elements): SBO = bytes between consecutive
// 4-row groups (512 when the rows are dense), LBO = bytes between consecutive 32-wide MN atoms.
__device__ __forceinline__ uint64_t umma_desc_sw128_base32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor (kind::f16 / kind::tf32): fp32 accumulate, dense, no negate.
//   fmt: 0 = f16, 1 = bf16, 2 = tf32;  a_mn/b_mn: 1 = operand is MN-major in shared memory
__host__ __device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;           // c_format = F32
  d |= fmt << 7;          // a_format
  d |= fmt << 10;         // b_format
  d |= (a_mn & 1u) << 15; // a_major
  d |= (b_mn & 1u) << 16; // b_major
  d |= (N >> 3) << 17;    // n_dim
  d |= (M >> 4) << 24;    // m_dim
  return d;
}

}  // namespace vb
