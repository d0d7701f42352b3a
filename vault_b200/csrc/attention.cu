// Fused masked-softmax attention, forward and backward, head_dim 64, sequence <= 512 with arbitrary per-key validity
// (text padding sits in the MIDDLE of the ViLT sequence: [text | text pad | image CLS | patches | patch pad]).
//
// Flash-style: scores never touch HBM.  One CTA = (sample, head, 64-row tile), 4 warps x 16 rows, K/V (or Q/dO) of the whole
// sequence resident in XOR-swizzled shared memory, fp32 online softmax in registers, log-sum-exp saved for the backward.
// Tensor work runs on mma.sync m16n8k16 bf16 (HMMA): attention is 4-8 % of the path's FLOPs, the dense GEMMs own tcgen05.
#include "common.cuh"

namespace vb {

constexpr int kHd = 64;       // head dim
constexpr int kTile = 64;     // rows per CTA / keys per chunk
constexpr int kAttnThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// 64-wide bf16 tiles: row r lives at r*128 bytes, its eight 16-byte chunks XOR-swizzled by (r & 7)
__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int col) {
  return base + (uint32_t)row * 128u + (uint32_t)((((col >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}
// rows [0, n_total) of a tile; rows >= n_valid are zero-filled.  src row stride in elements.
__device__ __forceinline__ void load_tile(uint32_t sbase, uint8_t* sgen, const bf16* src, long long stride, int n_valid, int n_total) {
  for (int idx = threadIdx.x; idx < n_total * 8; idx += kAttnThreads) {
    const int r = idx >> 3, ch = idx & 7;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
    if (r < n_valid) cp_async16(sbase + off, src + (long long)r * stride + ch * 8);
    else *reinterpret_cast<uint4*>(sgen + off) = make_uint4(0, 0, 0, 0);
  }
}
// Per-lane pieces of the swizzled ldmatrix addresses, computed once: every fragment load below is then ONE integer add
// (tile base + 2 KB * block + lane row + pre-XORed 16-byte chunk), which matters because these kernels are issue-bound.
struct LaneOff {
  uint32_t row_n;   // (lane&7) + (lane>>4)*8   rows of a non-transposed B fragment pair, * 128 B
  uint32_t row_t;   // (lane&7) + ((lane>>3)&1)*8   rows of a transposed B fragment pair, * 128 B
  uint32_t row_a;   // lane & 15   rows of an A fragment, * 128 B
  uint32_t xn[4];   // ((2*ks + ((lane>>3)&1)) ^ (lane&7)) << 4
  uint32_t xt[4];   // ((2*j  + (lane>>4))     ^ (lane&7)) << 4   (transposed B fragments and A fragments)
  __device__ __forceinline__ explicit LaneOff(int lane) {
    row_n = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * 128);
    row_t = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * 128);
    row_a = (uint32_t)((lane & 15) * 128);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xn[k] = (uint32_t)(((2 * k + ((lane >> 3) & 1)) ^ (lane & 7)) << 4);
      xt[k] = (uint32_t)(((2 * k + (lane >> 4)) ^ (lane & 7)) << 4);
    }
  }
};
// A fragments (16 rows x 64 cols) of a warp from a swizzled tile starting at row0 (multiple of 16)
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int row0, const LaneOff& lo, uint32_t (&a)[4][4]) {
  const uint32_t base = sbase + (uint32_t)row0 * 128u + lo.row_a;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(base + lo.xt[ks], a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}
// C[16 x 64] += A[16 x 64(d)] * T[64 rows x 64(d)]^T  where T rows are the n index (non-transposed ldmatrix): S = Q K^T
__device__ __forceinline__ void mma_a_tn(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int trow0, const LaneOff& lo) {
  const uint32_t base = sbase + (uint32_t)trow0 * 128u + lo.row_n;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int nb2 = 0; nb2 < 4; ++nb2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(base + nb2 * 2048 + lo.xn[ks], b0, b1, b2, b3);
      mma16816(c[2 * nb2], a[ks], b0, b1);
      mma16816(c[2 * nb2 + 1], a[ks], b2, b3);
    }
  }
}
// C[16 x 64(d)] += P[16 x 64(k)] * T[64 rows(k) x 64(d)]  (transposed ldmatrix): O = P V, dQ = dS K, dV = P^T dO, dK = dS^T Q
__device__ __forceinline__ void mma_p_t(float (&c)[8][4], const uint32_t (&p)[4][4], uint32_t sbase, int trow0, const LaneOff& lo) {
  const uint32_t base = sbase + (uint32_t)trow0 * 128u + lo.row_t;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int db2 = 0; db2 < 4; ++db2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(base + kk * 2048 + lo.xt[db2], b0, b1, b2, b3);
      mma16816(c[2 * db2], p[kk], b0, b1);
      mma16816(c[2 * db2 + 1], p[kk], b2, b3);
    }
  }
}
// accumulator layout (8 n-blocks x 4) -> bf16 A fragments for the next MMA
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
    a[kk][2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[kk][3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
}

// Dropout on attention probabilities: 16 random bits per (b,h,q,k), keep iff bits16 >= p * 65536; indexed only by logical
// coordinates so forward and both backward kernels agree.  One Philox call = 128 bits = 8 elements, and the element <-> bit assignment
// follows the MMA fragment layout so that no call is wasted: for query row q and 64-key chunk c, call (t, h) -- t = 0..3 the thread's
// column pair inside an 8-key block, h = 0..1 the half of the chunk -- holds in word i (low / high 16 bits) the keys
// 64c + 8(4h+i) + 2t (+1), i.e. exactly the 8 keys one thread of the forward / dQ kernels owns for that row and half.  The dK/dV
// kernel (keys as rows) needs per query one call for BOTH of its keys (they are 8 apart: adjacent words of the same call).
struct ProbDropout {
  unsigned long long seed;
  const unsigned long long* seed_dev;
  unsigned site;
  uint32_t thr16;
  float scale;
  int nchunk;  // ceil(S / 64)
  __device__ __forceinline__ unsigned long long effective_seed() const { return seed + (seed_dev ? *seed_dev : 0ull); }
  __device__ __forceinline__ uint4 bits(unsigned long long seed_eff, long long bh, int q, int chunk, int t, int half) const {
    return Philox(seed_eff)((((unsigned long long)((bh * 0x10000LL + q) * (long long)nchunk + chunk)) << 3) + (unsigned)(t * 2 + half), site);
  }
  __device__ __forceinline__ static uint32_t word(const uint4& r, int i) { return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w)); }
  // keep flags of the key pair of 8-key block j (0..7) of the chunk, given the two calls (halves) of this thread's row
  __device__ __forceinline__ void keep_pair(const uint4& h0, const uint4& h1, int j, bool& k0, bool& k1) const {
    const uint32_t w = word(j < 4 ? h0 : h1, j & 3);
    k0 = (w & 0xFFFFu) >= thr16;
    k1 = (w >> 16) >= thr16;
  }
};

struct AttnParams {
  const bf16* qkv;
  const uint8_t* key_mask;
  bf16* ctx;
  float* lse;
  const bf16* dctx;
  float* delta;
  bf16* dqkv;
  int B, S, heads, S_pad;
  float scale_log2;  // (1/sqrt(dh)) * log2(e)
  float scale;       // 1/sqrt(dh)
  ProbDropout drop;
};

// ================================================== forward ==================================================
template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  const int S = p.S, S_pad = p.S_pad, H = p.heads * kHd;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const LaneOff lo(lane);
  uint8_t* sQg = smem;
  uint8_t* sKg = sQg + kTile * 128;
  uint8_t* sVg = sKg + S_pad * 128;
  uint8_t* sMg = sVg + S_pad * 128;
  const uint32_t sQ = smem_u32(sQg), sK = smem_u32(sKg), sV = smem_u32(sVg);
  const long long ld = 3LL * H;
  const bf16* base = p.qkv + (long long)b * S * ld + h * kHd;
  const int q0 = qt * kTile;
  load_tile(sQ, sQg, base + (long long)q0 * ld, ld, min(kTile, S - q0), kTile);
  load_tile(sK, sKg, base + H, ld, S, S_pad);
  load_tile(sV, sVg, base + 2 * H, ld, S, S_pad);
  for (int i = threadIdx.x; i < S_pad; i += kAttnThreads) sMg[i] = i < S ? p.key_mask[(long long)b * S + i] : (uint8_t)0;
  cp_async_wait_all();
  __syncthreads();

  uint32_t aQ[4][4];
  load_a_frags(sQ, warp * 16, lo, aQ);
  float o[8][4];
  zero_acc(o);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int qr0 = q0 + warp * 16 + g, qr1 = qr0 + 8;
  const long long bh = (long long)b * p.heads + h;
  const unsigned long long seed_eff = DROP ? p.drop.effective_seed() : 0ull;

  for (int kc = 0; kc < S_pad; kc += kTile) {
    float s[8][4];
    zero_acc(s);
    mma_a_tn(s, aQ, sK, kc, lo);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kc + 8 * j + 2 * t;
      const bool v0 = sMg[key] != 0, v1 = sMg[key + 1] != 0;
      s[j][0] = v0 ? s[j][0] * p.scale_log2 : -INFINITY;
      s[j][1] = v1 ? s[j][1] * p.scale_log2 : -INFINITY;
      s[j][2] = v0 ? s[j][2] * p.scale_log2 : -INFINITY;
      s[j][3] = v1 ? s[j][3] * p.scale_log2 : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float mb0 = mn0 == -INFINITY ? 0.f : mn0, mb1 = mn1 == -INFINITY ? 0.f : mn1;
    const float al0 = fast_exp2(m0 - mb0), al1 = fast_exp2(m1 - mb1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = fast_exp2(s[j][0] - mb0); s[j][1] = fast_exp2(s[j][1] - mb0);
      s[j][2] = fast_exp2(s[j][2] - mb1); s[j][3] = fast_exp2(s[j][3] - mb1);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
      o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1;
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
    if (DROP) {
      const uint4 a0 = p.drop.bits(seed_eff, bh, qr0, kc >> 6, t, 0), a1 = p.drop.bits(seed_eff, bh, qr0, kc >> 6, t, 1);
      const uint4 b0 = p.drop.bits(seed_eff, bh, qr1, kc >> 6, t, 0), b1 = p.drop.bits(seed_eff, bh, qr1, kc >> 6, t, 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bool k0, k1;
        p.drop.keep_pair(a0, a1, j, k0, k1);
        s[j][0] = k0 ? s[j][0] * p.drop.scale : 0.f;
        s[j][1] = k1 ? s[j][1] * p.drop.scale : 0.f;
        p.drop.keep_pair(b0, b1, j, k0, k1);
        s[j][2] = k0 ? s[j][2] * p.drop.scale : 0.f;
        s[j][3] = k1 ? s[j][3] * p.drop.scale : 0.f;
      }
    }
    uint32_t aP[4][4];
    c_to_a(s, aP);
    mma_p_t(o, aP, sV, kc, lo);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  bf16* crow0 = p.ctx + ((long long)b * S + qr0) * H + h * kHd;
  bf16* crow1 = p.ctx + ((long long)b * S + qr1) * H + h * kHd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (qr0 < S) *reinterpret_cast<uint32_t*>(crow0 + 8 * j + 2 * t) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    if (qr1 < S) *reinterpret_cast<uint32_t*>(crow1 + 8 * j + 2 * t) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
  if (t == 0 && p.lse != nullptr) {
    if (qr0 < S) p.lse[bh * S + qr0] = (m0 == -INFINITY ? 0.f : m0) * kLn2 + __logf(fmaxf(l0, 1e-30f));
    if (qr1 < S) p.lse[bh * S + qr1] = (m1 == -INFINITY ? 0.f : m1) * kLn2 + __logf(fmaxf(l1, 1e-30f));
  }
}

// ================================================== backward: dQ (+ delta) ==================================================
// CTA = (b, h, 64 query rows).  delta[q] = sum_d dO[q,d] * O[q,d];  dQ = scale * sum_k dS[q,k] K[k,:]
template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dq_kernel(const AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  const int S = p.S, S_pad = p.S_pad, H = p.heads * kHd;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const LaneOff lo(lane);
  uint8_t* sQg = smem;
  uint8_t* sdOg = sQg + kTile * 128;
  uint8_t* sOg = sdOg + kTile * 128;
  uint8_t* sKg = sOg + kTile * 128;
  uint8_t* sVg = sKg + S_pad * 128;
  uint8_t* sMg = sVg + S_pad * 128;
  const uint32_t sQ = smem_u32(sQg), sdO = smem_u32(sdOg), sO = smem_u32(sOg), sK = smem_u32(sKg), sV = smem_u32(sVg);
  const long long ld = 3LL * H;
  const bf16* base = p.qkv + (long long)b * S * ld + h * kHd;
  const int q0 = qt * kTile;
  const int nq = min(kTile, S - q0);
  load_tile(sQ, sQg, base + (long long)q0 * ld, ld, nq, kTile);
  load_tile(sdO, sdOg, p.dctx + ((long long)b * S + q0) * H + h * kHd, H, nq, kTile);
  load_tile(sO, sOg, p.ctx + ((long long)b * S + q0) * H + h * kHd, H, nq, kTile);
  load_tile(sK, sKg, base + H, ld, S, S_pad);
  load_tile(sV, sVg, base + 2 * H, ld, S, S_pad);
  for (int i = threadIdx.x; i < S_pad; i += kAttnThreads) sMg[i] = i < S ? p.key_mask[(long long)b * S + i] : (uint8_t)0;
  cp_async_wait_all();
  __syncthreads();

  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const int qr0 = q0 + r0, qr1 = q0 + r1;
  const long long bh = (long long)b * p.heads + h;
  // delta for this thread's two rows
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = 8 * j + 2 * t;
    uint32_t x, y;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(sw_addr(sdO, r0, col)));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(y) : "r"(sw_addr(sO, r0, col)));
    float2 a = unpack_bf16x2(x), c = unpack_bf16x2(y);
    d0 += a.x * c.x + a.y * c.y;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(sw_addr(sdO, r1, col)));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(y) : "r"(sw_addr(sO, r1, col)));
    a = unpack_bf16x2(x); c = unpack_bf16x2(y);
    d1 += a.x * c.x + a.y * c.y;
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
  d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  if (t == 0) {
    if (qr0 < S) p.delta[bh * S + qr0] = d0;
    if (qr1 < S) p.delta[bh * S + qr1] = d1;
  }
  const float lse0 = qr0 < S ? p.lse[bh * S + qr0] * kLog2e : 0.f;
  const float lse1 = qr1 < S ? p.lse[bh * S + qr1] * kLog2e : 0.f;

  uint32_t aQ[4][4], adO[4][4];
  load_a_frags(sQ, warp * 16, lo, aQ);
  load_a_frags(sdO, warp * 16, lo, adO);
  float dq[8][4];
  zero_acc(dq);
  const unsigned long long seed_eff = DROP ? p.drop.effective_seed() : 0ull;
  for (int kc = 0; kc < S_pad; kc += kTile) {
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_tn(s, aQ, sK, kc, lo);
    mma_a_tn(dp, adO, sV, kc, lo);
    uint4 a0, a1, b0, b1;
    if (DROP) {
      a0 = p.drop.bits(seed_eff, bh, qr0, kc >> 6, t, 0); a1 = p.drop.bits(seed_eff, bh, qr0, kc >> 6, t, 1);
      b0 = p.drop.bits(seed_eff, bh, qr1, kc >> 6, t, 0); b1 = p.drop.bits(seed_eff, bh, qr1, kc >> 6, t, 1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kc + 8 * j + 2 * t;
      const bool v0 = sMg[key] != 0, v1 = sMg[key + 1] != 0;
      const float p00 = v0 ? fast_exp2(s[j][0] * p.scale_log2 - lse0) : 0.f;
      const float p01 = v1 ? fast_exp2(s[j][1] * p.scale_log2 - lse0) : 0.f;
      const float p10 = v0 ? fast_exp2(s[j][2] * p.scale_log2 - lse1) : 0.f;
      const float p11 = v1 ? fast_exp2(s[j][3] * p.scale_log2 - lse1) : 0.f;
      if (DROP) {
        bool k0, k1;
        p.drop.keep_pair(a0, a1, j, k0, k1);
        dp[j][0] = k0 ? dp[j][0] * p.drop.scale : 0.f;
        dp[j][1] = k1 ? dp[j][1] * p.drop.scale : 0.f;
        p.drop.keep_pair(b0, b1, j, k0, k1);
        dp[j][2] = k0 ? dp[j][2] * p.drop.scale : 0.f;
        dp[j][3] = k1 ? dp[j][3] * p.drop.scale : 0.f;
      }
      s[j][0] = p00 * (dp[j][0] - d0);
      s[j][1] = p01 * (dp[j][1] - d0);
      s[j][2] = p10 * (dp[j][2] - d1);
      s[j][3] = p11 * (dp[j][3] - d1);
    }
    uint32_t aS[4][4];
    c_to_a(s, aS);
    mma_p_t(dq, aS, sK, kc, lo);
  }
  bf16* drow0 = p.dqkv + ((long long)b * S + qr0) * ld + h * kHd;
  bf16* drow1 = p.dqkv + ((long long)b * S + qr1) * ld + h * kHd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (qr0 < S) *reinterpret_cast<uint32_t*>(drow0 + 8 * j + 2 * t) = pack_bf16x2(dq[j][0] * p.scale, dq[j][1] * p.scale);
    if (qr1 < S) *reinterpret_cast<uint32_t*>(drow1 + 8 * j + 2 * t) = pack_bf16x2(dq[j][2] * p.scale, dq[j][3] * p.scale);
  }
}

// ================================================== backward: dK, dV ==================================================
// CTA = (b, h, 64 keys), warp = 16 keys; loops over 64-query chunks with Q, dO of the whole sequence in smem.
//   S^T = K Q^T ; P^T = exp2(S^T*c - lse[q]) ; dV += P_drop^T dO ; dP^T = V dO^T ; dS^T = P^T (dP^T_drop - delta[q]) ; dK += scale dS^T Q
template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dkv_kernel(const AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  const int S = p.S, S_pad = p.S_pad, H = p.heads * kHd;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const LaneOff lo(lane);
  uint8_t* sKg = smem;
  uint8_t* sVg = sKg + kTile * 128;
  uint8_t* sQg = sVg + kTile * 128;
  uint8_t* sdOg = sQg + S_pad * 128;
  float* sLse = reinterpret_cast<float*>(sdOg + S_pad * 128);
  float* sDelta = sLse + S_pad;
  const uint32_t sK = smem_u32(sKg), sV = smem_u32(sVg), sQ = smem_u32(sQg), sdO = smem_u32(sdOg);
  const long long ld = 3LL * H;
  const bf16* base = p.qkv + (long long)b * S * ld + h * kHd;
  const int k0 = kt * kTile;
  const int nk = min(kTile, S - k0);
  const long long bh = (long long)b * p.heads + h;
  load_tile(sK, sKg, base + H + (long long)k0 * ld, ld, nk, kTile);
  load_tile(sV, sVg, base + 2 * H + (long long)k0 * ld, ld, nk, kTile);
  load_tile(sQ, sQg, base, ld, S, S_pad);
  load_tile(sdO, sdOg, p.dctx + (long long)b * S * H + h * kHd, H, S, S_pad);
  for (int i = threadIdx.x; i < S_pad; i += kAttnThreads) {
    sLse[i] = i < S ? p.lse[bh * S + i] * kLog2e : 0.f;
    sDelta[i] = i < S ? p.delta[bh * S + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  const int kr0 = k0 + warp * 16 + g, kr1 = kr0 + 8;  // this thread's two keys (rows of the transposed problem)
  const bool kv0 = kr0 < S && p.key_mask[(long long)b * S + kr0] != 0;
  const bool kv1 = kr1 < S && p.key_mask[(long long)b * S + kr1] != 0;
  uint32_t aK[4][4], aV[4][4];
  load_a_frags(sK, warp * 16, lo, aK);
  load_a_frags(sV, warp * 16, lo, aV);
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const unsigned long long seed_eff = DROP ? p.drop.effective_seed() : 0ull;
  for (int qc = 0; qc < S_pad; qc += kTile) {
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_tn(s, aK, sQ, qc, lo);    // [16 keys x 64 queries]
    mma_a_tn(dp, aV, sdO, qc, lo);  // dP^T
    float pt[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = qc + 8 * j + 2 * t;  // columns = queries q, q+1
      const bool qv0 = q < S, qv1 = q + 1 < S;
      const float l0 = sLse[q], l1 = sLse[q + 1];
      const float dl0 = sDelta[q], dl1 = sDelta[q + 1];
      float p00 = (kv0 && qv0) ? fast_exp2(s[j][0] * p.scale_log2 - l0) : 0.f;
      float p01 = (kv0 && qv1) ? fast_exp2(s[j][1] * p.scale_log2 - l1) : 0.f;
      float p10 = (kv1 && qv0) ? fast_exp2(s[j][2] * p.scale_log2 - l0) : 0.f;
      float p11 = (kv1 && qv1) ? fast_exp2(s[j][3] * p.scale_log2 - l1) : 0.f;
      float e00 = dp[j][0], e01 = dp[j][1], e10 = dp[j][2], e11 = dp[j][3];
      float pd00 = p00, pd01 = p01, pd10 = p10, pd11 = p11;
      if (DROP) {
        // element (q, key): this thread's keys kr0 and kr1 = kr0 + 8 sit in 8-key blocks 2*warp and 2*warp + 1 of the CTA's chunk, at
        // column pair t' = g >> 1, slot g & 1: adjacent words of ONE call of the forward's layout (see ProbDropout)
        const int wsel = (warp & 1) * 2, sh = (g & 1) * 16;
        const uint4 r0 = p.drop.bits(seed_eff, bh, q, kt, g >> 1, warp >> 1);
        const uint4 r1 = p.drop.bits(seed_eff, bh, q + 1, kt, g >> 1, warp >> 1);
        bool keep = ((ProbDropout::word(r0, wsel) >> sh) & 0xFFFFu) >= p.drop.thr16;
        pd00 = keep ? p00 * p.drop.scale : 0.f; e00 = keep ? e00 * p.drop.scale : 0.f;
        keep = ((ProbDropout::word(r1, wsel) >> sh) & 0xFFFFu) >= p.drop.thr16;
        pd01 = keep ? p01 * p.drop.scale : 0.f; e01 = keep ? e01 * p.drop.scale : 0.f;
        keep = ((ProbDropout::word(r0, wsel + 1) >> sh) & 0xFFFFu) >= p.drop.thr16;
        pd10 = keep ? p10 * p.drop.scale : 0.f; e10 = keep ? e10 * p.drop.scale : 0.f;
        keep = ((ProbDropout::word(r1, wsel + 1) >> sh) & 0xFFFFu) >= p.drop.thr16;
        pd11 = keep ? p11 * p.drop.scale : 0.f; e11 = keep ? e11 * p.drop.scale : 0.f;
      }
      pt[j][0] = pd00; pt[j][1] = pd01; pt[j][2] = pd10; pt[j][3] = pd11;
      s[j][0] = p00 * (e00 - dl0);
      s[j][1] = p01 * (e01 - dl1);
      s[j][2] = p10 * (e10 - dl0);
      s[j][3] = p11 * (e11 - dl1);
    }
    uint32_t aP[4][4];
    c_to_a(pt, aP);
    mma_p_t(dv, aP, sdO, qc, lo);
    c_to_a(s, aP);
    mma_p_t(dk, aP, sQ, qc, lo);
  }
  bf16* krow0 = p.dqkv + ((long long)b * S + kr0) * ld + H + h * kHd;
  bf16* krow1 = p.dqkv + ((long long)b * S + kr1) * ld + H + h * kHd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = 8 * j + 2 * t;
    if (kr0 < S) {
      *reinterpret_cast<uint32_t*>(krow0 + col) = pack_bf16x2(dk[j][0] * p.scale, dk[j][1] * p.scale);
      *reinterpret_cast<uint32_t*>(krow0 + H + col) = pack_bf16x2(dv[j][0], dv[j][1]);
    }
    if (kr1 < S) {
      *reinterpret_cast<uint32_t*>(krow1 + col) = pack_bf16x2(dk[j][2] * p.scale, dk[j][3] * p.scale);
      *reinterpret_cast<uint32_t*>(krow1 + H + col) = pack_bf16x2(dv[j][2], dv[j][3]);
    }
  }
}

static int fill_params(AttnParams& p, int B, int S, int heads, float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site) {
  VB_REQUIRE(B > 0 && S > 0 && heads > 0, "attention: bad shape B=%d S=%d heads=%d", B, S, heads);
  VB_REQUIRE(S <= 512, "attention: S=%d > 512 not supported (whole-sequence K/V live in shared memory)", S);
  VB_REQUIRE(S < 65536, "attention: S too large for the dropout counter");
  VB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attention: dropout_p=%f", dropout_p);
  p.B = B; p.S = S; p.heads = heads;
  p.S_pad = (S + kTile - 1) / kTile * kTile;
  p.scale = 0.125f;  // 1/sqrt(64)
  p.scale_log2 = 0.125f * kLog2e;
  p.drop.seed = seed; p.drop.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev); p.drop.site = site;
  p.drop.thr16 = (uint32_t)(dropout_p * 65536.0f + 0.5f);
  p.drop.scale = 1.0f / (1.0f - dropout_p);
  p.drop.nchunk = p.S_pad / kTile;
  return VAULT_OK;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "attention: cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
  return VAULT_OK;
}

// tensor-memory (tcgen05) kernels of attention_tc.cu
bool attn_tc_fwd_ok(int S, float dropout_p);
bool attn_tc_bwd_ok(int S, float dropout_p);
int attn_fwd_tc(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int B, int S, int heads, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta, void* dqkv, int B, int S,
                int heads, cudaStream_t st);
void attn_set_impl(int impl);
// pipelined tcgen05 kernels of attention_sm100.cu (S <= 384, no dropout)
bool attn_sm100_ok(int S, float dropout_p);
void attn_sm100_enable(int on);
void attn_sm100_fwd_pp(int on);
int attn_fwd_sm100(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int B, int S, int heads, float dropout_p, unsigned long long seed,
                   const unsigned long long* seed_dev, unsigned site, cudaStream_t st);
int attn_bwd_sm100(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta, void* dqkv, int B, int S,
                   int heads, float dropout_p, unsigned long long seed, const unsigned long long* seed_dev, unsigned site, cudaStream_t st);

}  // namespace vb

extern "C" int vault_attn_set_impl(int32_t impl) {
  using namespace vb;
  VB_REQUIRE(impl >= 0 && impl <= 4, "vault_attn_set_impl: impl=%d (0 auto, 1 mma.sync only, 2 whole-row tcgen05 kernels where the shape allows, "
             "3 pipelined tcgen05 kernels for every S <= 384, 4 = 3 with the one-tile-per-warpgroup forward)", impl);
  attn_set_impl(impl >= 3 ? 0 : impl);
  attn_sm100_enable(impl == 0 ? 1 : (impl >= 3 ? 2 : 0));
  attn_sm100_fwd_pp(impl == 4 ? 1 : 0);
  return VAULT_OK;
}

extern "C" int vault_attn_fwd(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int32_t B, int32_t S, int32_t heads,
                              float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site, void* stream) {
  using namespace vb;
  VB_REQUIRE(qkv && key_mask && ctx, "attn_fwd: null pointer");
  AttnParams p{};
  int rc = fill_params(p, B, S, heads, dropout_p, seed, seed_dev, site);
  if (rc) return rc;
  if (attn_sm100_ok(S, dropout_p))
    return attn_fwd_sm100(qkv, key_mask, ctx, lse, B, S, heads, dropout_p, seed, reinterpret_cast<const unsigned long long*>(seed_dev), site,
                          reinterpret_cast<cudaStream_t>(stream));
  if (attn_tc_fwd_ok(S, dropout_p)) return attn_fwd_tc(qkv, key_mask, ctx, lse, B, S, heads, reinterpret_cast<cudaStream_t>(stream));
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse;
  const int smem = (kTile + 2 * p.S_pad) * 128 + p.S_pad;
  dim3 grid(p.S_pad / kTile, heads, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dropout_p > 0.f) {
    if ((rc = set_smem(attn_fwd_kernel<true>, smem))) return rc;
    launch(attn_fwd_kernel<true>, dim3(grid), dim3(kAttnThreads), smem, st, p);
  } else {
    if ((rc = set_smem(attn_fwd_kernel<false>, smem))) return rc;
    launch(attn_fwd_kernel<false>, dim3(grid), dim3(kAttnThreads), smem, st, p);
  }
  return check_launch("attn_fwd_kernel");
}

extern "C" int vault_attn_bwd(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse, float* delta,
                              void* dqkv, int32_t B, int32_t S, int32_t heads, float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                              void* stream) {
  using namespace vb;
  VB_REQUIRE(qkv && key_mask && ctx && dctx && lse && delta && dqkv, "attn_bwd: null pointer");
  AttnParams p{};
  int rc = fill_params(p, B, S, heads, dropout_p, seed, seed_dev, site);
  if (rc) return rc;
  if (attn_sm100_ok(S, dropout_p))
    return attn_bwd_sm100(qkv, key_mask, ctx, dctx, lse, delta, dqkv, B, S, heads, dropout_p, seed, reinterpret_cast<const unsigned long long*>(seed_dev), site,
                          reinterpret_cast<cudaStream_t>(stream));
  if (attn_tc_bwd_ok(S, dropout_p))
    return attn_bwd_tc(qkv, key_mask, ctx, dctx, lse, delta, dqkv, B, S, heads, reinterpret_cast<cudaStream_t>(stream));
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.key_mask = key_mask; p.ctx = reinterpret_cast<bf16*>(const_cast<void*>(ctx));
  p.lse = const_cast<float*>(lse); p.dctx = reinterpret_cast<const bf16*>(dctx); p.delta = delta; p.dqkv = reinterpret_cast<bf16*>(dqkv);
  dim3 grid(p.S_pad / kTile, heads, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int smem_dq = (3 * kTile + 2 * p.S_pad) * 128 + p.S_pad;
  const int smem_dkv = (2 * kTile + 2 * p.S_pad) * 128 + 2 * p.S_pad * 4;
  if (dropout_p > 0.f) {
    if ((rc = set_smem(attn_bwd_dq_kernel<true>, smem_dq))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<true>, smem_dkv))) return rc;
    launch(attn_bwd_dq_kernel<true>, dim3(grid), dim3(kAttnThreads), smem_dq, st, p);
    if ((rc = check_launch("attn_bwd_dq_kernel"))) return rc;
    launch(attn_bwd_dkv_kernel<true>, dim3(grid), dim3(kAttnThreads), smem_dkv, st, p);
  } else {
    if ((rc = set_smem(attn_bwd_dq_kernel<false>, smem_dq))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<false>, smem_dkv))) return rc;
    launch(attn_bwd_dq_kernel<false>, dim3(grid), dim3(kAttnThreads), smem_dq, st, p);
    if ((rc = check_launch("attn_bwd_dq_kernel"))) return rc;
    launch(attn_bwd_dkv_kernel<false>, dim3(grid), dim3(kAttnThreads), smem_dkv, st, p);
  }
  return check_launch("attn_bwd_dkv_kernel");
}
