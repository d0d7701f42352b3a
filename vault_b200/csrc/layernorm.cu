// LayerNorm forward / backward, HBM-bound: one warp per row, the row lives in registers (128-bit loads), warp-shuffle
// reductions, two-pass variance.  x is the fp32 residual stream; y goes out as bf16 (GEMM operand) and/or fp32 (BERT's
// post-LN residual).  Optional Philox dropout on the output (BERT embeddings, HF:models/bert/modeling_bert.py:110-111).
#include "common.cuh"

namespace vb {

constexpr int kLnWarps = 8;

#ifndef VB_LN_FWD_CTAS
#define VB_LN_FWD_CTAS 4  // resident CTAs per SM asked of the compiler: 4 (64 registers) 11.34 us at 11,808 x 768; 0 = its own choice (56 registers) 11.5; 5 (48 registers, spills) 12.9
#endif
#if VB_LN_FWD_CTAS > 0
#define VB_LN_FWD_BOUNDS __launch_bounds__(kLnWarps * 32, VB_LN_FWD_CTAS)
#else
#define VB_LN_FWD_BOUNDS __launch_bounds__(kLnWarps * 32)
#endif
template <int VPL>  // float4 vectors per lane: cols = 128 * VPL
__global__ void VB_LN_FWD_BOUNDS
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ y_bf16,
              float* __restrict__ y_f32, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, float eps,
              float drop_p, unsigned long long seed, const unsigned long long* seed_dev, unsigned site) {
  pdl_enter();
  constexpr int cols = 128 * VPL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kLnWarps + warp;
  if (row >= rows) return;
  if (drop_p > 0.f && seed_dev) seed += *seed_dev;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = __ldg(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / cols);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / cols) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const uint32_t thr = dropout_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (drop_p > 0.f) {
      const uint4 bits = dropout_bits4(seed, site, (unsigned long long)(row * (cols / 4) + c4));
      o.x = bits.x >= thr ? o.x * keep_scale : 0.f;
      o.y = bits.y >= thr ? o.y * keep_scale : 0.f;
      o.z = bits.z >= thr ? o.z * keep_scale : 0.f;
      o.w = bits.w >= thr ? o.w * keep_scale : 0.f;
    }
    if (y_f32) reinterpret_cast<float4*>(y_f32 + row * cols)[c4] = o;
    if (y_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y_bf16 + row * cols)[c4] = pk;
    }
  }
}

// Backward.  HBM-bound (16-20 B per element) but each row needs two warp-wide reductions between its loads and its stores, so
// a load-then-compute warp leaves the memory system idle most of the time.  Here one producer warp streams whole rows
// (x, dy, residual gradient) into a 16-slot shared-memory ring with 1-D bulk TMA copies (cp.async.bulk + mbarrier
// complete_tx), ~150 KB in flight per SM, and 8 consumer warps take rows from the ring: the memory pipe never waits for
// arithmetic.  Per-column sums (dgamma, dbeta, bias-gradient column sum of the bf16 output) stay in the consumers' registers
// and meet once per CTA: O(SMs * cols) fp32 atomics.
constexpr int kLnBwdWarps = 8;    // consumer warps
constexpr int kLnBwdSlots = 16;   // rows in flight per CTA
#ifndef VB_LN_PRODUCERS
#define VB_LN_PRODUCERS 4  // measured at 11,808 x 768: 1 lane 30.5 us, 2 lanes 29.6, 4 lanes 29.4
#endif
constexpr int kLnBwdProducers = VB_LN_PRODUCERS;  // producer lanes

// The consumers' per-column partial sums (up to three sets: dgamma, dbeta, bias-gradient column sums) meet once per CTA: one trip through
// shared memory for all sets together (two named barriers in total), then one red.global.add.v4 per 4 columns and set.
template <int VPL>
__device__ __forceinline__ void cta_colsum_flush3(float4 (&a0)[VPL], float4 (&a1)[VPL], float4 (&a2)[VPL], float* __restrict__ d0,
                                                  float* __restrict__ d1, float* __restrict__ d2, uint8_t* smem, int warp, int lane) {
  constexpr int kRow = 32 * VPL + 1;
  float4* red = reinterpret_cast<float4*>(smem);  // [3][kLnBwdWarps][kRow]
  asm volatile("bar.sync 1, %0;" ::"n"(kLnBwdWarps * 32));  // every consumer is done with the ring
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (d0) red[(0 * kLnBwdWarps + warp) * kRow + lane + 32 * i] = a0[i];
    if (d1) red[(1 * kLnBwdWarps + warp) * kRow + lane + 32 * i] = a1[i];
    if (d2) red[(2 * kLnBwdWarps + warp) * kRow + lane + 32 * i] = a2[i];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kLnBwdWarps * 32));
  auto reduce_set = [&](int s, float* __restrict__ dst) {
    if (dst == nullptr) return;
    for (int c4 = warp * 32 + lane; c4 < 32 * VPL; c4 += kLnBwdWarps * 32) {
      float4 a = red[(s * kLnBwdWarps) * kRow + c4];
#pragma unroll
      for (int w = 1; w < kLnBwdWarps; ++w) {
        const float4 t = red[(s * kLnBwdWarps + w) * kRow + c4];
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c4), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    }
  };
  reduce_set(0, d0);
  reduce_set(1, d1);
  reduce_set(2, d2);
}

template <int VPL>
__global__ void __launch_bounds__((kLnBwdWarps + 1) * 32, 1)
ln_bwd_kernel(const float* __restrict__ dy_f32, const bf16* __restrict__ dy_bf16, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres, float* __restrict__ dx_f32,
              bf16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum, long long rows,
              float drop_p, unsigned site, float out_p, unsigned out_site, unsigned long long seed, const unsigned long long* seed_dev) {
  constexpr int cols = 128 * VPL;
  constexpr uint32_t kRow32 = cols * 4, kRow16 = cols * 2;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  // slot layout: [x f32][dres f32 (opt)][dy f32 (opt)][dy bf16 (opt)]
  const uint32_t off_res = kRow32;
  const uint32_t off_dy32 = off_res + (dres ? kRow32 : 0u);
  const uint32_t off_dy16 = off_dy32 + (dy_f32 ? kRow32 : 0u);
  const uint32_t slot_bytes = off_dy16 + (dy_bf16 ? kRow16 : 0u);
  const uint32_t sbase = smem_u32(ln_smem);
  const uint32_t bars = sbase + kLnBwdSlots * slot_bytes;  // full[16] | empty[16]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kLnBwdSlots; ++s) {
      mbar_init(bars + 8u * s, 1);
      mbar_init(bars + 8u * (kLnBwdSlots + s), 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();
  const long long my_rows = rows > blockIdx.x ? (rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;  // rows b, b+grid, ...

  if (warp == kLnBwdWarps) {
    // ===================== producer: bulk-copy rows into the ring =====================
    if (lane < kLnBwdProducers) {  // lane l streams the rows k = l, l + P, ...: P issue threads in lock-step (one thread's wait + expect_tx + 3-4
                                   // bulk copies per row is ~450 cycles, close to the ~540 cycles a row may take at full HBM rate)
      for (long long k = lane; k < my_rows; k += kLnBwdProducers) {
        const int slot = (int)(k % kLnBwdSlots);
        const uint32_t ph = (uint32_t)((k / kLnBwdSlots) & 1);
        mbar_wait(bars + 8u * (kLnBwdSlots + slot), ph ^ 1u);
        const long long row = blockIdx.x + k * gridDim.x;
        const uint32_t dst = sbase + slot * slot_bytes, fb = bars + 8u * slot;
        mbar_expect_tx(fb, slot_bytes);
        bulk_load(dst, x + row * cols, kRow32, fb);
        if (dres) bulk_load(dst + off_res, dres + row * cols, kRow32, fb);
        if (dy_f32) bulk_load(dst + off_dy32, dy_f32 + row * cols, kRow32, fb);
        if (dy_bf16) bulk_load(dst + off_dy16, dy_bf16 + row * cols, kRow16, fb);
      }
    }
    return;
  }

  // ===================== consumers =====================
  if ((drop_p > 0.f || out_p > 0.f) && seed_dev) seed += *seed_dev;
  const uint32_t out_thr = dropout_threshold(out_p);
  const float out_scale = out_p > 0.f ? 1.0f / (1.0f - out_p) : 1.0f;
  const uint32_t thr = dropout_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  float4 dg[VPL], db[VPL], dcs[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) dg[i] = db[i] = dcs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long k = warp; k < my_rows; k += kLnBwdWarps) {
    const int slot = (int)(k % kLnBwdSlots);
    const uint32_t ph = (uint32_t)((k / kLnBwdSlots) & 1);
    const long long row = blockIdx.x + k * gridDim.x;
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);  // issued before the wait: off the critical path
    mbar_wait(bars + 8u * slot, ph);
    const uint8_t* sl = ln_smem + slot * slot_bytes;
    const float4* sx = reinterpret_cast<const float4*>(sl);
    float4 d[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c4 = lane + 32 * i;
      const float4 xv = sx[c4];
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy_f32) dv = reinterpret_cast<const float4*>(sl + off_dy32)[c4];
      if (dy_bf16) {
        const uint2 pk = reinterpret_cast<const uint2*>(sl + off_dy16)[c4];
        const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
        dv.x += a.x; dv.y += a.y; dv.z += b.x; dv.w += b.y;
      }
      if (drop_p > 0.f) {
        const uint4 bits = dropout_bits4(seed, site, (unsigned long long)(row * (cols / 4) + c4));
        dv.x = bits.x >= thr ? dv.x * keep_scale : 0.f;
        dv.y = bits.y >= thr ? dv.y * keep_scale : 0.f;
        dv.z = bits.z >= thr ? dv.z * keep_scale : 0.f;
        dv.w = bits.w >= thr ? dv.w * keep_scale : 0.f;
      }
      const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += dv.x * xh.x; dg[i].y += dv.y * xh.y; dg[i].z += dv.z * xh.z; dg[i].w += dv.w * xh.w;
      db[i].x += dv.x; db[i].y += dv.y; db[i].z += dv.z; db[i].w += dv.w;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);  // L1-resident
      d[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh.x + d[i].y * xh.y) + (d[i].z * xh.z + d[i].w * xh.w);
    }
    const float c1 = warp_sum(s1) * (1.0f / cols);
    const float c2 = warp_sum(s2) * (1.0f / cols);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c4 = lane + 32 * i;
      const float4 xv = sx[c4];
      float4 o;
      o.x = rs * (d[i].x - c1 - (xv.x - mu) * rs * c2);
      o.y = rs * (d[i].y - c1 - (xv.y - mu) * rs * c2);
      o.z = rs * (d[i].z - c1 - (xv.z - mu) * rs * c2);
      o.w = rs * (d[i].w - c1 - (xv.w - mu) * rs * c2);
      if (dres) {
        const float4 r = reinterpret_cast<const float4*>(sl + off_res)[c4];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx_f32 + row * cols)[c4] = o;
      if (dx_bf16) {
        if (out_p > 0.f) {
          const uint4 bits = dropout_bits4(seed, out_site, (unsigned long long)(row * (cols / 4) + c4));
          o.x = bits.x >= out_thr ? o.x * out_scale : 0.f;
          o.y = bits.y >= out_thr ? o.y * out_scale : 0.f;
          o.z = bits.z >= out_thr ? o.z * out_scale : 0.f;
          o.w = bits.w >= out_thr ? o.w * out_scale : 0.f;
        }
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * cols)[c4] = pk;
        if (dcolsum) {  // sum exactly what the consuming GEMMs will read (bf16-rounded, dropout applied)
          const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
          dcs[i].x += a.x; dcs[i].y += a.y; dcs[i].z += b.x; dcs[i].w += b.y;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8u * (kLnBwdSlots + slot));  // slot free for the producer
  }
  if (dgamma == nullptr && dbeta == nullptr && dcolsum == nullptr) return;
  // the ring is idle now (every row of this CTA has been consumed by the time all consumers pass the first bar.sync)
  cta_colsum_flush3<VPL>(dg, db, dcs, dgamma, dbeta, dcolsum, ln_smem, warp, lane);
}

template <int VPL>
int ln_fwd_launch(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean, float* rstd, long long rows,
                  float eps, float p, unsigned long long seed, const unsigned long long* seed_dev, unsigned site, cudaStream_t st) {
  const long long grid = (rows + kLnWarps - 1) / kLnWarps;
  launch(ln_fwd_kernel<VPL>, dim3((unsigned)grid), dim3(kLnWarps * 32), 0, st, x, gamma, beta, reinterpret_cast<bf16*>(y_bf16), y_f32, mean, rstd, rows, eps, p,
                                                               seed, seed_dev, site);
  return check_launch("ln_fwd_kernel");
}
template <int VPL>
int ln_bwd_launch(const float* dy_f32, const void* dy_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dcolsum, long long rows, float p,
                  unsigned site, float out_p, unsigned out_site, unsigned long long seed, const unsigned long long* seed_dev, cudaStream_t st) {
  constexpr int cols = 128 * VPL;
  const int slot_bytes = cols * 4 + (dres ? cols * 4 : 0) + (dy_f32 ? cols * 4 : 0) + (dy_bf16 ? cols * 2 : 0);
  int smem = kLnBwdSlots * slot_bytes + 2 * kLnBwdSlots * 8;
  const int red_bytes = 3 * kLnBwdWarps * (32 * VPL + 1) * 16;
  if (smem < red_bytes) smem = red_bytes;
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "ln_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 200 * 1024;
  }
  if (smem > 200 * 1024) return fail(VAULT_ERR_INVALID, "ln_bwd: row too wide for the shared-memory ring (%d bytes)", smem);
  long long grid = rows;
  const long long cap = (long long)device_sm_count();
  if (grid > cap) grid = cap;
  launch(ln_bwd_kernel<VPL>, dim3((unsigned)grid), dim3((kLnBwdWarps + 1) * 32), (size_t)smem, st, dy_f32, reinterpret_cast<const bf16*>(dy_bf16), x, mean,
         rstd, gamma, dres, dx_f32, reinterpret_cast<bf16*>(dx_bf16), dgamma, dbeta, dcolsum, rows, p, site, out_p, out_site, seed, seed_dev);
  return check_launch("ln_bwd_kernel");
}

}  // namespace vb

#define VB_LN_DISPATCH(cols, CALL)                                                                      \
  switch ((cols) / 128) {                                                                               \
    case 1: return CALL(1);                                                                             \
    case 2: return CALL(2);                                                                             \
    case 4: return CALL(4);                                                                             \
    case 6: return CALL(6);                                                                             \
    case 8: return CALL(8);                                                                             \
    default: return vb::fail(VAULT_ERR_INVALID, "layernorm: cols=%d not in {128,256,512,768,1024}", (int)(cols)); \
  }

extern "C" int vault_layernorm_fwd_drop(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean, float* rstd,
                                        int64_t rows, int32_t cols, float eps, float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                                        void* stream) {
  using namespace vb;
  VB_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
  VB_REQUIRE(rows >= 0 && cols > 0 && cols % 128 == 0, "layernorm_fwd: bad shape rows=%lld cols=%d", (long long)rows, cols);
  VB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "layernorm_fwd: dropout_p=%f", dropout_p);
  if (rows == 0) return VAULT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CALL(V) ln_fwd_launch<V>(x, gamma, beta, y_bf16, y_f32, mean, rstd, rows, eps, dropout_p, seed, reinterpret_cast<const unsigned long long*>(seed_dev), site, st)
  VB_LN_DISPATCH(cols, CALL)
#undef CALL
}

extern "C" int vault_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean, float* rstd,
                                   int64_t rows, int32_t cols, float eps, void* stream) {
  return vault_layernorm_fwd_drop(x, gamma, beta, y_bf16, y_f32, mean, rstd, rows, cols, eps, 0.f, 0, nullptr, 0, stream);
}

extern "C" int vault_layernorm_bwd_drop(const float* dy_f32, const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                        const float* gamma, const float* dres_f32, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                        float* dcolsum, int64_t rows, int32_t cols, float in_p, uint32_t in_site, float out_p, uint32_t out_site,
                                        uint64_t seed, const uint64_t* seed_dev, void* stream) {
  using namespace vb;
  VB_REQUIRE((dy_f32 || dy_bf16) && x && mean && rstd && gamma && dx_f32, "layernorm_bwd: null pointer");
  VB_REQUIRE(rows >= 0 && cols > 0 && cols % 128 == 0, "layernorm_bwd: bad shape rows=%lld cols=%d", (long long)rows, cols);
  if (rows == 0) return VAULT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define CALL(V) ln_bwd_launch<V>(dy_f32, dy_bf16, x, mean, rstd, gamma, dres_f32, dx_f32, dx_bf16, dgamma, dbeta, dcolsum, rows, in_p, in_site, out_p, out_site, seed, reinterpret_cast<const unsigned long long*>(seed_dev), st)
  VB_LN_DISPATCH(cols, CALL)
#undef CALL
}

extern "C" int vault_layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                   const float* gamma, const float* dres_f32, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                   int64_t rows, int32_t cols, void* stream) {
  return vault_layernorm_bwd_drop(dy_f32, dy_bf16, x, mean, rstd, gamma, dres_f32, dx_f32, dx_bf16, dgamma, dbeta, nullptr, rows, cols, 0.f, 0, 0.f, 0, 0,
                                  nullptr, stream);
}
