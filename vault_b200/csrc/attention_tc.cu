// Fused masked-softmax attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory), head_dim 64,
// whole key range of one (sample, head) resident in shared memory, no dropout (the ViLT stack; the BERT stack with its
// probability dropout and 40-128 token sequences stays on the mma.sync kernels of attention.cu).
//
//   forward : CTA = (b, h, 128 query rows), 128 threads, thread = query row = TMEM lane.
//             TMA {Q, K, V} -> S = Q K^T (UMMA 128 x S_pad x 64, fp32 in TMEM) -> tcgen05.ld, thread-local row max / exp2 / sum
//             -> P (bf16) written to smem in the 128B-swizzled K-major operand layout -> O = P V (V read un-transposed as the
//             MN-major B operand) -> scale by 1/rowsum, store ctx + log-sum-exp.  TMEM 256 columns, ~97 KB smem: 2 CTAs / SM.
//   backward: CTA = (b, h), 256 threads, key-block major, nothing recomputed twice and no atomics:
//             per 128-key block   S^T = K Q^T, dP^T = V dO^T  (two UMMAs into TMEM, thread = key row)
//                                 P^T = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - delta[q])   -> bf16 -> smem (K-major, K = queries)
//                                 dV = P^T dO, dK = dS^T Q (B operands MN-major), dQ += dS K  (A = dS^T read as MN-major, M = queries)
//             dQ accumulates in TMEM across the key blocks; dK / dV leave after each block.  TMEM 512 columns, 1 CTA / SM.
// The softmax convention (natural-log LSE of the scaled, masked scores) is the one of attention.cu, so either forward feeds
// either backward.
#include "common.cuh"

namespace vb {

int encode_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer);

namespace {

constexpr float kLog2eTc = 1.4426950408889634f;
constexpr float kLn2Tc = 0.6931471805599453f;

struct AttnTcParams {
  const uint8_t* key_mask;  // [B,S]
  bf16* ctx;                // fwd out / bwd in  [B*S, H]
  float* lse;               // [B,heads,S]
  const bf16* dctx;         // [B*S, H]
  float* delta;             // [B,heads,S] (optional output of the backward)
  bf16* dqkv;               // [B*S, 3H]
  int B, S, heads;
  int S_pad;  // key/query range the MMAs cover: fwd multiple of 32, bwd multiple of 64
  int S64;    // rows of the K / V (/ Q / dO) tiles in shared memory (multiple of the 64-row TMA box)
  float scale_log2, scale;
  long long* trace;  // VAULT_B200_ATTN_TRACE=1: clock64 stamps of thread 0 (16 per CTA), read back and printed by the host
};

#define VB_STAMP(i)                                                   \
  do {                                                                \
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[(long long)cta_lin * 16 + (i)] = clock64(); \
  } while (0)

// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 32 consecutive elements (64 bytes) of row `row` starting at element column 32*c of a K-major, 128B-swizzled operand whose
// 64-element column chunks are `chunk_stride` bytes apart (rows 128 bytes apart inside a chunk).  pk = 16 packed bf16 pairs.
__device__ __forceinline__ void store_row32(uint32_t buf, uint32_t chunk_stride, int row, int c, const uint32_t (&pk)[16]) {
  const uint32_t rowaddr = buf + (uint32_t)(c >> 1) * chunk_stride + (uint32_t)row * 128u;
  const uint32_t seg0 = (uint32_t)(c & 1) * 4u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t seg = (seg0 + i) ^ ((uint32_t)row & 7u);
    st_shared_v4(rowaddr + (seg << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// A warp owns 32 accumulator rows (lane = row) of 64 fp32 columns in TMEM and writes them, scaled and rounded to bf16, to 32 rows
// of a row-major global matrix.  Each lane storing its own 128-byte row would cost 32 cache lines per store instruction, so the
// rows are staged in this warp's 4 KB of shared memory (16-byte segments XOR-swizzled by row) and leave 4 full rows per instruction.
//   stage: smem address of the warp's 32 x 128 B staging tile;  gptr0: global address of row 0 / column 0;  n_valid: rows to write
__device__ __forceinline__ void store_rows_coalesced(uint32_t taddr, uint32_t stage, float mul, bf16* gptr0, long long row_stride, int n_valid,
                                                     int lane) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)(c * 32), r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * mul, __uint_as_float(r[2 * i + 1]) * mul);
    store_row32(stage, 0u, lane, c, pk);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = 4 * it + (lane >> 3), seg = lane & 7;
    const uint4 v = ld_shared_v4(stage + (uint32_t)rr * 128u + (uint32_t)((seg ^ (rr & 7)) << 4));
    if (rr < n_valid) *reinterpret_cast<uint4*>(gptr0 + (long long)rr * row_stride + seg * 8) = v;
  }
  __syncwarp();
}

// ============================================================ forward ============================================================
constexpr int kFwdThreads = 128;

__global__ void __launch_bounds__(kFwdThreads) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int S = p.S, S_pad = p.S_pad, S64 = p.S64, H = p.heads * 64;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  VB_STAMP(0);

  const uint32_t sQ = base;  // 128 x 64 bf16; reused as chunk 0 (keys 0..63) of P once S = Q K^T has completed
  const uint32_t sK = sQ + 16384u;
  const uint32_t sV = sK + (uint32_t)S64 * 128u;
  const uint32_t sP1 = sV + (uint32_t)S64 * 128u;  // chunks 1.. of P
  const uint32_t misc_off = 16384u + 2u * (uint32_t)S64 * 128u + (uint32_t)(S64 / 64 - 1) * 16384u;
  const uint32_t misc = base + misc_off;
  const uint32_t bar_load = misc, bar_s = misc + 8, bar_o = misc + 16, tmem_slot = misc + 24;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + misc_off + 24);
  uint32_t* mw_s = reinterpret_cast<uint32_t*>(gen + misc_off + 32);
  const uint32_t ncols = S_pad <= 64 ? 64u : (S_pad <= 128 ? 128u : 256u);

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  VB_STAMP(1);
  pdl_enter();
  VB_STAMP(2);

  const int row_g0 = b * S;  // first row of this sample in the [B*S, 3H] buffer
  if (tid == 0) {
    mbar_expect_tx(bar_load, (uint32_t)(2 + 2 * (S64 / 64)) * 8192u);
    tma_load_2d(sQ, &tmQKV, bar_load, h * 64, row_g0 + qb * 128);
    tma_load_2d(sQ + 8192u, &tmQKV, bar_load, h * 64, row_g0 + qb * 128 + 64);
    for (int j = 0; j < S64 / 64; ++j) {
      tma_load_2d(sK + (uint32_t)j * 8192u, &tmQKV, bar_load, H + h * 64, row_g0 + 64 * j);
      tma_load_2d(sV + (uint32_t)j * 8192u, &tmQKV, bar_load, 2 * H + h * 64, row_g0 + 64 * j);
    }
  }
  // key validity as bit words (keys beyond S are rows of the next sample or TMA zero fill: invalid)
  const int nch = S_pad / 32;
  for (int c = warp; c < nch; c += kFwdThreads / 32) {
    const int key = 32 * c + lane;
    const bool v = key < S && p.key_mask[(long long)b * S + key] != 0;
    const uint32_t w = __ballot_sync(0xffffffffu, v);
    if (lane == 0) mw_s[c] = w;
  }
  __syncthreads();
  if (tid == 0) {
    VB_STAMP(3);
    mbar_wait(bar_load, 0);
    VB_STAMP(4);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(1u, 128u, (uint32_t)S_pad, 0u, 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc_mma_f16(tmem, umma_desc_sw128(sQ + k * 32u, 16u, 1024u), umma_desc_sw128(sK + k * 32u, 16u, 1024u), idesc, k > 0 ? 1u : 0u);
    tc_commit(bar_s);
  }
  __syncwarp();
  uint32_t mw[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mw[c] = c < nch ? mw_s[c] : 0u;

  const int row = qb * 128 + tid;                    // query index inside the sample
  const bool warp_active = qb * 128 + warp * 32 < S;  // warp-uniform: rows beyond S are never stored
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  float l = 0.f, msc = 0.f;
  mbar_wait(bar_s, 0);
  VB_STAMP(5);
  tc_fence_after();
  if (warp_active) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nch) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (mw[c] == 0xffffffffu) {  // all 32 keys valid (the common case): no per-element select
          float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[i])); m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[i + 2])); m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
          }
          m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, ((mw[c] >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
        }
      }
    }
    msc = m == -INFINITY ? 0.f : m * p.scale_log2;
    VB_STAMP(6);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nch) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        uint32_t pk[16];
        if (mw[c] == 0xffffffffu) {
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc));
            l0 += p0; l1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          l += l0 + l1;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ((mw[c] >> (2 * i)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -msc)) : 0.f;
            const float p1 = ((mw[c] >> (2 * i + 1)) & 1u) ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -msc)) : 0.f;
            l += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        }
        // chunk 0 of P overlays Q (dead: the S MMA has completed), the others follow V
        const uint32_t buf = (c >> 1) == 0 ? sQ : sP1 - 16384u;
        store_row32(buf, 16384u, tid, c, pk);
      }
    }
  }
  VB_STAMP(7);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  VB_STAMP(8);
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc(1u, 128u, 64u, 0u, 1u);
    for (int ks = 0; ks < S_pad / 16; ++ks) {
      const uint32_t pa = ((ks >> 2) == 0 ? sQ : sP1 + (uint32_t)((ks >> 2) - 1) * 16384u) + (uint32_t)(ks & 3) * 32u;
      tc_mma_f16(tmem, umma_desc_sw128(pa, 16u, 1024u), umma_desc_sw128(sV + (uint32_t)ks * 2048u, 8192u, 1024u), idesc, ks > 0 ? 1u : 0u);
    }
    tc_commit(bar_o);
  }
  __syncwarp();
  VB_STAMP(9);
  mbar_wait(bar_o, 0);
  VB_STAMP(10);
  tc_fence_after();
  if (warp_active) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    // K is dead since the S MMA completed: its first 16 KB stage the output tile (4 KB per warp)
    store_rows_coalesced(trow, sK + (uint32_t)warp * 4096u, inv, p.ctx + ((long long)row_g0 + qb * 128 + warp * 32) * H + h * 64, (long long)H,
                         S - (qb * 128 + warp * 32), lane);
    if (row < S && p.lse != nullptr) p.lse[((long long)b * p.heads + h) * S + row] = msc * kLn2Tc + __logf(fmaxf(l, 1e-30f));
  }
  VB_STAMP(11);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
  VB_STAMP(12);
}

// ============================================================ backward ============================================================
constexpr int kBwdThreads = 256;

__global__ void __launch_bounds__(kBwdThreads, 1)
    attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int S = p.S, S_pad = p.S_pad, H = p.heads * 64;  // S_pad in {128, 192}
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  VB_STAMP(0);

  const uint32_t tile = (uint32_t)S_pad * 128u;                // one [S_pad x 64] bf16 operand tile
  const uint32_t pbuf = (uint32_t)(S_pad / 64) * 16384u;       // one [128 keys x S_pad queries] bf16 buffer
  const uint32_t sQ = base, sK = sQ + tile, sV = sK + tile, sdO = sV + tile;
  const uint32_t sdS = sdO + tile;   // dS^T of the current key block
  const uint32_t sPt = sdS + pbuf;   // P^T of the current key block
  const uint32_t f_off = 4u * tile + 2u * pbuf;
  float* sLse = reinterpret_cast<float*>(gen + f_off);
  float* sDelta = sLse + S_pad;
  const uint32_t misc_off = f_off + 2u * (uint32_t)S_pad * 4u;
  const uint32_t misc = base + misc_off;
  const uint32_t bar_load = misc, bar_s = misc + 8, bar_o = misc + 16, tmem_slot = misc + 24;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + misc_off + 24);

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  VB_STAMP(1);
  pdl_enter();

  const long long row_g0 = (long long)b * S;
  const long long bh = (long long)b * p.heads + h;
  const uint32_t id_s = umma_idesc(1u, 128u, (uint32_t)S_pad, 0u, 0u);
  const uint32_t col_dP = (uint32_t)S_pad, col_dV = 0u, col_dK = 64u, col_dQ = 384u;
  // S^T = K Q^T and dP^T = V dO^T of key block kb (one thread)
  auto issue_scores = [&](int kb) {
    const uint32_t aK = sK + (uint32_t)kb * 16384u, aV = sV + (uint32_t)kb * 16384u;  // rows past a tile run into the next tile: finite, masked
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc_mma_f16(tmem, umma_desc_sw128(aK + k * 32u, 16u, 1024u), umma_desc_sw128(sQ + k * 32u, 16u, 1024u), id_s, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc_mma_f16(tmem + col_dP, umma_desc_sw128(aV + k * 32u, 16u, 1024u), umma_desc_sw128(sdO + k * 32u, 16u, 1024u), id_s, k > 0 ? 1u : 0u);
    tc_commit(bar_s);
  };
  if (warp == 0) {
    // warp 0: operand loads and the first key block's score MMAs, while warps 1-7 compute delta
    if (lane == 0) {
      mbar_expect_tx(bar_load, 4u * tile);
      for (int j = 0; j < S_pad / 64; ++j) {
        const int r = (int)row_g0 + 64 * j;
        tma_load_2d(sK + (uint32_t)j * 8192u, &tmQKV, bar_load, H + h * 64, r);
        tma_load_2d(sQ + (uint32_t)j * 8192u, &tmQKV, bar_load, h * 64, r);
        tma_load_2d(sV + (uint32_t)j * 8192u, &tmQKV, bar_load, 2 * H + h * 64, r);
        tma_load_2d(sdO + (uint32_t)j * 8192u, &tmDO, bar_load, h * 64, r);
      }
      mbar_wait(bar_load, 0);
      tc_fence_after();
      issue_scores(0);
    }
    __syncwarp();
  } else {
    // delta[q] = sum_d dO[q,d] O[q,d], 8 lanes per query row (one 16-byte segment each), all loads of a thread in flight together;
    // queries beyond S get lse = +inf so that their probabilities vanish
    constexpr int kDeltaThreads = kBwdThreads - 32;
    uint4 xs[7], ys[7];
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int idx = (tid - 32) + it * kDeltaThreads, q = idx >> 3, seg = idx & 7;
      xs[it] = ys[it] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < S_pad * 8 && q < S) {
        xs[it] = *reinterpret_cast<const uint4*>(p.dctx + (row_g0 + q) * H + h * 64 + seg * 8);
        ys[it] = *reinterpret_cast<const uint4*>(p.ctx + (row_g0 + q) * H + h * 64 + seg * 8);
      }
    }
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int idx = (tid - 32) + it * kDeltaThreads, q = idx >> 3, seg = idx & 7;
      float2 u, v;
      float d = 0.f;
      u = unpack_bf16x2(xs[it].x); v = unpack_bf16x2(ys[it].x); d += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(xs[it].y); v = unpack_bf16x2(ys[it].y); d += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(xs[it].z); v = unpack_bf16x2(ys[it].z); d += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(xs[it].w); v = unpack_bf16x2(ys[it].w); d += u.x * v.x + u.y * v.y;
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (seg == 0 && idx < S_pad * 8) {
        sDelta[q] = d;
        sLse[q] = q < S ? p.lse[bh * S + q] * kLog2eTc : INFINITY;
        if (q < S && p.delta != nullptr) p.delta[bh * S + q] = d;
      }
    }
  }
  __syncthreads();
  VB_STAMP(2);

  const int nkb = (S + 127) / 128;            // key blocks of 128
  const int nqb = S_pad > 128 ? 2 : 1;        // 128-row dQ blocks; the second one starts at S_pad - 128 (overlap, no over-read)
  const int nchh = S_pad / 64;                // 32-column chunks per warp half
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);

  for (int kb = 0; kb < nkb; ++kb) {
    const int key0 = kb * 128;
    if (tid == 0 && kb > 0) {
      VB_STAMP(3 + 6 * kb);
      tc_fence_after();
      issue_scores(kb);
    }
    __syncwarp();
    const int krow = quad * 32 + lane;  // row of this thread in the key block = TMEM lane
    const int key = key0 + krow;
    const bool kvalid = key < S && p.key_mask[row_g0 + key] != 0;
    const bool warp_live = key0 + quad * 32 < S;  // warp-uniform
    mbar_wait(bar_s, (uint32_t)(kb & 1));
    VB_STAMP(4 + 6 * kb);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      if (cc < nchh) {
        const int c = half * nchh + cc;  // 32-query chunk
        uint32_t pkP[16], pkS[16];
        if (warp_live) {
          uint32_t s[32], dp[32];
          tmem_ld32(trow + (uint32_t)(c * 32), s);
          tmem_ld32(trow + col_dP + (uint32_t)(c * 32), dp);
          tmem_ld_wait();
          const float4* l4 = reinterpret_cast<const float4*>(sLse + c * 32);
          const float4* d4 = reinterpret_cast<const float4*>(sDelta + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ls = l4[i], dl = d4[i];
            const float p0 = kvalid ? fast_exp2(fmaf(__uint_as_float(s[4 * i]), p.scale_log2, -ls.x)) : 0.f;
            const float p1 = kvalid ? fast_exp2(fmaf(__uint_as_float(s[4 * i + 1]), p.scale_log2, -ls.y)) : 0.f;
            const float p2 = kvalid ? fast_exp2(fmaf(__uint_as_float(s[4 * i + 2]), p.scale_log2, -ls.z)) : 0.f;
            const float p3 = kvalid ? fast_exp2(fmaf(__uint_as_float(s[4 * i + 3]), p.scale_log2, -ls.w)) : 0.f;
            pkP[2 * i] = pack_bf16x2(p0, p1);
            pkP[2 * i + 1] = pack_bf16x2(p2, p3);
            pkS[2 * i] = pack_bf16x2(p0 * (__uint_as_float(dp[4 * i]) - dl.x), p1 * (__uint_as_float(dp[4 * i + 1]) - dl.y));
            pkS[2 * i + 1] = pack_bf16x2(p2 * (__uint_as_float(dp[4 * i + 2]) - dl.z), p3 * (__uint_as_float(dp[4 * i + 3]) - dl.w));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pkP[i] = pkS[i] = 0u;
        }
        store_row32(sPt, 16384u, krow, c, pkP);
        store_row32(sdS, 16384u, krow, c, pkS);
      }
    }
    VB_STAMP(5 + 6 * kb);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    VB_STAMP(6 + 6 * kb);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id_kn = umma_idesc(1u, 128u, 64u, 0u, 1u);  // A K-major (P^T / dS^T rows = keys), B MN-major
      for (int ks = 0; ks < S_pad / 16; ++ks) {
        const uint32_t off = (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u;
        tc_mma_f16(tmem + col_dV, umma_desc_sw128(sPt + off, 16u, 1024u), umma_desc_sw128(sdO + (uint32_t)ks * 2048u, 8192u, 1024u), id_kn,
                   ks > 0 ? 1u : 0u);
      }
      for (int ks = 0; ks < S_pad / 16; ++ks) {
        const uint32_t off = (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u;
        tc_mma_f16(tmem + col_dK, umma_desc_sw128(sdS + off, 16u, 1024u), umma_desc_sw128(sQ + (uint32_t)ks * 2048u, 8192u, 1024u), id_kn,
                   ks > 0 ? 1u : 0u);
      }
      // dQ[q,:] += sum_keys dS[q,key] K[key,:] : A = dS^T buffer read as the MN-major operand (M = queries, 64-query chunks 16 KB apart)
      const uint32_t id_nn = umma_idesc(1u, 128u, 64u, 1u, 1u);
      for (int qb = 0; qb < nqb; ++qb) {
        const uint32_t qoff = (uint32_t)((qb == 0 ? 0 : S_pad - 128) / 64) * 16384u;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          tc_mma_f16(tmem + col_dQ + (uint32_t)(64 * qb), umma_desc_sw128(sdS + qoff + (uint32_t)ks * 2048u, 16384u, 1024u),
                     umma_desc_sw128(sK + (uint32_t)key0 * 128u + (uint32_t)ks * 2048u, 8192u, 1024u), id_nn, (kb > 0 || ks > 0) ? 1u : 0u);
      }
      tc_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, (uint32_t)(kb & 1));
    VB_STAMP(7 + 6 * kb);
    tc_fence_after();
    {
      // half 0 stores dV rows, half 1 stores dK rows (x softmax scale) of this key block; staging in the (now idle) P^T buffer
      const int r0 = key0 + quad * 32;
      store_rows_coalesced(trow + (half == 0 ? col_dV : col_dK), sPt + (uint32_t)warp * 4096u, half == 0 ? 1.f : p.scale,
                           p.dqkv + (row_g0 + r0) * 3LL * H + (half == 0 ? 2 * H : H) + h * 64, 3LL * H, S - r0, lane);
    }
    tc_fence_before();
    __syncthreads();  // TMEM columns 0.. and the P^T / dS^T buffers are rewritten by the next key block
    VB_STAMP(8 + 6 * kb);
  }
  // dQ: warp half = 128-row block; block 1 starts at query S_pad - 128 and only its rows >= 128 are new
  tc_fence_after();
  if (half < nqb) {
    const int r0 = (half == 0 ? 0 : S_pad - 128) + quad * 32;  // first query row of this warp
    if (half == 0 || r0 >= 128)
      store_rows_coalesced(trow + col_dQ + (uint32_t)(64 * half), sPt + (uint32_t)warp * 4096u, p.scale, p.dqkv + (row_g0 + r0) * 3LL * H + h * 64,
                           3LL * H, S - r0, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512u);
  VB_STAMP(15);
}

// VAULT_B200_ATTN_TRACE=1 (debug): per-CTA phase stamps, printed (synchronously) after the launch
long long* trace_buffer(int ctas) {
  static long long* buf = nullptr;
  static int cap = 0;
  const char* e = getenv("VAULT_B200_ATTN_TRACE");
  if (e == nullptr || e[0] != '1') return nullptr;
  if (ctas > cap) {
    if (buf) cudaFree(buf);
    cudaMalloc(&buf, (size_t)ctas * 16 * sizeof(long long));
    cap = ctas;
  }
  cudaMemset(buf, 0, (size_t)ctas * 16 * sizeof(long long));
  return buf;
}
void trace_print(const char* what, const long long* dbuf, int ctas) {
  if (dbuf == nullptr) return;
  long long* h = (long long*)malloc((size_t)ctas * 16 * sizeof(long long));
  cudaDeviceSynchronize();
  cudaMemcpy(h, dbuf, (size_t)ctas * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  const int picks[4] = {0, ctas / 3, ctas / 2, ctas - 1};
  for (int k = 0; k < 4; ++k) {
    const long long* t = h + (size_t)picks[k] * 16;
    printf("%s cta %d:", what, picks[k]);
    for (int i = 1; i < 16; ++i) printf(" %lld", t[i] ? t[i] - t[0] : -1LL);
    printf("\n");
  }
  free(h);
}

int g_attn_impl = 0;  // 0 auto, 1 mma.sync kernels only, 2 tcgen05 kernels wherever the shape allows

template <typename K>
int set_smem_tc(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "attention: cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
  return VAULT_OK;
}

}  // namespace

// shapes served by the tensor-memory kernels (no dropout): forward 65..256 keys, backward 65..192
bool attn_tc_fwd_ok(int S, float dropout_p) { return g_attn_impl != 1 && dropout_p == 0.f && S > 64 && S <= 256; }
bool attn_tc_bwd_ok(int S, float dropout_p) { return g_attn_impl != 1 && dropout_p == 0.f && S > 64 && S <= 192; }

int attn_fwd_tc(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int B, int S, int heads, cudaStream_t st) {
  AttnTcParams p{};
  p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse;
  p.B = B; p.S = S; p.heads = heads;
  p.S_pad = (S + 31) / 32 * 32;
  p.S64 = (S + 63) / 64 * 64;
  p.scale = 0.125f; p.scale_log2 = 0.125f * kLog2eTc;
  const int H = heads * 64;
  CUtensorMap tm;
  int rc = encode_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  const int smem = 1024 + 16384 + 2 * p.S64 * 128 + (p.S64 / 64 - 1) * 16384 + 128;
  if ((rc = set_smem_tc(attn_fwd_tc_kernel, smem))) return rc;
  const int ctas = (S + 127) / 128 * heads * B;
  p.trace = trace_buffer(ctas);
  launch(attn_fwd_tc_kernel, dim3((S + 127) / 128, heads, B), dim3(kFwdThreads), (size_t)smem, st, tm, p);
  trace_print("attn_fwd_tc", p.trace, ctas);
  return check_launch("attn_fwd_tc_kernel");
}

int attn_bwd_tc(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta, void* dqkv, int B, int S,
                int heads, cudaStream_t st) {
  AttnTcParams p{};
  p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(const_cast<void*>(ctx)); p.lse = const_cast<float*>(lse);
  p.dctx = reinterpret_cast<const bf16*>(dctx); p.delta = delta; p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.B = B; p.S = S; p.heads = heads;
  p.S_pad = p.S64 = (S + 63) / 64 * 64;
  p.scale = 0.125f; p.scale_log2 = 0.125f * kLog2eTc;
  const int H = heads * 64;
  CUtensorMap tmQ, tmD;
  int rc = encode_tmap_2d(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H, (uint64_t)B * S, (uint64_t)3 * H, 64, 64);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dctx, (uint64_t)H, (uint64_t)B * S, (uint64_t)H, 64, 64);
  if (rc) return rc;
  const int smem = 1024 + 4 * p.S_pad * 128 + 2 * (p.S_pad / 64) * 16384 + 2 * p.S_pad * 4 + 64;
  if ((rc = set_smem_tc(attn_bwd_tc_kernel, smem))) return rc;
  p.trace = trace_buffer(heads * B);
  launch(attn_bwd_tc_kernel, dim3(heads, B), dim3(kBwdThreads), (size_t)smem, st, tmQ, tmD, p);
  trace_print("attn_bwd_tc", p.trace, heads * B);
  return check_launch("attn_bwd_tc_kernel");
}

void attn_set_impl(int impl) { g_attn_impl = impl; }

}  // namespace vb
