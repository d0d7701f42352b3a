// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (accumulators in TMEM,
// double-buffered) -> 8 epilogue warps (tcgen05.ld, smem transpose, fused epilogue, coalesced global stores).
//
//   C[M,N] = sum_k A[m,k] * B[n,k]     A, B bf16, fp32 accumulate
//
// Either operand may be stored K-major (k contiguous) or MN-major (m/n contiguous); the only differences are the TMA box
// walk, the UMMA smem-descriptor strides and two bits of the instruction descriptor, all run-time values.  That one
// kernel therefore serves forward (K,K), dgrad (K,MN: W is read un-transposed) and wgrad (MN,MN: dy and x read
// un-transposed, contraction over tokens, optional split-K with fp32 atomics).
//
// Warp roles (384 threads, 1 CTA/SM):  warp0 TMA producer | warp1 MMA issuer | warp2 TMEM alloc | warp3 idle | warps4-11 epilogue
// (VB_EPI_WARPS=16: 640 threads, four epilogue warps per TMEM lane quadrant working in 16-column chunks)
#include "common.cuh"

#ifndef VB_EPI_DIRECT
#define VB_EPI_DIRECT 0  // 1 = no smem transpose: every thread applies the epilogue to its own row (the TMEM lane) and stores 16-byte pieces
#endif
#ifndef VB_GELU_PACKED
#define VB_GELU_PACKED 0  // 1 = GELU / GELU' epilogue arithmetic on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2): 26 % fewer instructions per chunk,
                          // yet measured EQUAL for the forward (64.6 vs 64.4 us at 11808x3072x768) and 6 us SLOWER for GELU' (71.9 -> 78.1 us):
                          // these epilogues are not instruction-bound -- the forward writes two bf16 streams (145 MB per launch, 2.3 TB/s of
                          // pure writes next to the operand reads), i.e. it sits on the HBM write roof, not on the issue slots
#endif
#ifndef VB_EPI_WARPS
#define VB_EPI_WARPS 8  // 16 (four warps per TMEM lane quadrant, 16-column chunks) measured 3-4 % SLOWER on every shape: kept for A/B builds
#endif

namespace vb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = VB_EPI_WARPS;
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps: 8 or 16");
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kRingBytes = 196608;                 // smem ring for A/B stages
constexpr int kCW = kEpiWarps == 8 ? 32 : 16;      // columns of one transposed chunk (32 rows x kCW fp32 per warp)
constexpr int kLPR = kCW / 4;                      // lanes per row after the transpose (one float4 each)
constexpr int kRPP = 32 / kLPR;                    // rows per pass
constexpr int kNP = 32 / kRPP;                     // passes per chunk
constexpr int kStagingBytes = kEpiWarps * 32 * kCW * 4;  // per-warp 32 x kCW fp32 transpose buffers
// XOR swizzle of the 16-byte slots of a staging row (row = 128 B for 32-column chunks, 64 B for 16-column chunks): conflict-free
// for the row-per-lane writes and for the kRPP-rows-per-pass reads
__device__ __forceinline__ int stg_swz(int row) { return kCW == 32 ? (row & 7) : ((row >> 1) & 3); }
constexpr int kSchedDepth = 4;
constexpr int kSmemBytes = kRingBytes + kStagingBytes + 1024 /*align slack*/ + 384 /*barriers + tile ring*/;

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int num_m_blocks, num_n_blocks, num_k_blocks, split_k, kb_per_split;
  int* sched;  // optional {next tile, finished CTAs} counters: dynamic tile scheduling (robust to SMs taken by other kernels)
  int cl;  // 0: no cluster; 1: CTA pair along M (B tile multicast); 2: CTA pair along N (A tile multicast)
  const float* bias;
  const float* resid;
  long long ldr;
  const bf16* aux;
  long long ldaux;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  float dropout_p;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  unsigned site;
  float* a_colsum;  // optional (MN-major A, fp32 epilogues): [M] += sum_k A[m,k], taken from the A tiles as they pass through the smem ring
};

template <int EPI> constexpr bool kGeluFwd = EPI == VAULT_EPI_BIAS_GELU_BF16 || EPI == VAULT_EPI_BIAS_GELU_GRAD_BF16;  // +bias, GELU, optional second bf16 stream
template <int EPI> constexpr bool kAuxMul = EPI == VAULT_EPI_DGELU_BF16 || EPI == VAULT_EPI_MUL_AUX_BF16;              // bf16 side input per output element

// ---- fused epilogues -------------------------------------------------------------------------------------------------
// After the smem transpose each lane owns 4 consecutive columns of kNP rows (row stride kRPP) of a 32 x kCW chunk.  Pointers are
// formed once per chunk and advanced by a constant row stride; full tiles skip every bounds check (GUARD=false).
template <int EPI>
struct EpiPtrs {
  char* out;          // bf16 or fp32
  char* out2;         // bf16 (GELU pre-activation), may be null
  const char* side;   // fp32 residual or bf16 pre-activation
  long long out_step, out2_step, side_step;  // bytes per kRPP rows
};

template <int EPI>
__device__ __forceinline__ EpiPtrs<EPI> make_ptrs(const GemmParams& p, long long row, int col) {
  constexpr bool kOutF32 = EPI == VAULT_EPI_BIAS_RESID_F32 || EPI == VAULT_EPI_ATOMIC_F32 || EPI == VAULT_EPI_BIAS_F32 || EPI == VAULT_EPI_STORE_F32 ||
                           EPI == VAULT_EPI_ATOMIC_BIAS_DROP_F32;
  constexpr int osz = kOutF32 ? 4 : 2;
  EpiPtrs<EPI> e;
  e.out = reinterpret_cast<char*>(p.out) + (row * p.ldo + col) * osz;
  e.out_step = kRPP * p.ldo * osz;
  e.out2 = nullptr; e.out2_step = 0; e.side = nullptr; e.side_step = 0;
  if constexpr (kGeluFwd<EPI>) {
    if (p.out2) { e.out2 = reinterpret_cast<char*>(p.out2) + (row * p.ldo2 + col) * 2; e.out2_step = 2 * kRPP * p.ldo2; }
  }
  if constexpr (EPI == VAULT_EPI_BIAS_RESID_F32) { e.side = reinterpret_cast<const char*>(p.resid) + (row * p.ldr + col) * 4; e.side_step = 4 * kRPP * p.ldr; }
  if constexpr (kAuxMul<EPI>) { e.side = reinterpret_cast<const char*>(p.aux) + (row * p.ldaux + col) * 2; e.side_step = 2 * kRPP * p.ldaux; }
  return e;
}

template <int EPI>
__device__ __forceinline__ float4 load_side(const char* ptr) {
  if constexpr (EPI == VAULT_EPI_BIAS_RESID_F32) {
    return __ldg(reinterpret_cast<const float4*>(ptr));
  } else if constexpr (kAuxMul<EPI>) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(ptr));
    const float2 a01 = unpack_bf16x2(a.x), a23 = unpack_bf16x2(a.y);
    return make_float4(a01.x, a01.y, a23.x, a23.y);
  } else {
    return make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue4(const GemmParams& p, float4 acc, const float4& b4, const float4& side, unsigned long long seed,
                                          unsigned long long drop_idx, char* out, char* out2) {
  if constexpr (EPI == VAULT_EPI_BIAS_BF16) {
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(acc.x + b4.x, acc.y + b4.y), pack_bf16x2(acc.z + b4.z, acc.w + b4.w));
  } else if constexpr (EPI == VAULT_EPI_BIAS_GELU_BF16 /*exact*/) {
#if VB_GELU_PACKED
    const float2 xa = add2(make_float2(acc.x, acc.y), make_float2(b4.x, b4.y)), xb = add2(make_float2(acc.z, acc.w), make_float2(b4.z, b4.w));
    if (out2) *reinterpret_cast<uint2*>(out2) = make_uint2(pack_bf16x2(xa.x, xa.y), pack_bf16x2(xb.x, xb.y));
    const float2 ga = gelu_erf2(xa), gb = gelu_erf2(xb);
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(ga.x, ga.y), pack_bf16x2(gb.x, gb.y));
#else
    const float x0 = acc.x + b4.x, x1 = acc.y + b4.y, x2 = acc.z + b4.z, x3 = acc.w + b4.w;
#ifndef VB_DIAG_NO_OUT2
    if (out2) *reinterpret_cast<uint2*>(out2) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
#endif
#ifdef VB_DIAG_NO_GELU
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
#else
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(gelu_erf(x0), gelu_erf(x1)), pack_bf16x2(gelu_erf(x2), gelu_erf(x3)));
#endif
#endif
  } else if constexpr (EPI == VAULT_EPI_BIAS_RESID_F32) {
    float x0 = acc.x + b4.x, x1 = acc.y + b4.y, x2 = acc.z + b4.z, x3 = acc.w + b4.w;
    if (p.dropout_p > 0.f) {
      const uint32_t thr = dropout_threshold(p.dropout_p);
      const float sc = 1.0f / (1.0f - p.dropout_p);
      const uint4 bits = dropout_bits4(seed, p.site, drop_idx);
      x0 = bits.x >= thr ? x0 * sc : 0.f;
      x1 = bits.y >= thr ? x1 * sc : 0.f;
      x2 = bits.z >= thr ? x2 * sc : 0.f;
      x3 = bits.w >= thr ? x3 * sc : 0.f;
    }
    *reinterpret_cast<float4*>(out) = make_float4(side.x + x0, side.y + x1, side.z + x2, side.w + x3);
  } else if constexpr (EPI == VAULT_EPI_PLAIN_BF16) {
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
  } else if constexpr (EPI == VAULT_EPI_DGELU_BF16) {
#if VB_GELU_PACKED
    const float2 da = mul2(make_float2(acc.x, acc.y), gelu_erf_grad2(make_float2(side.x, side.y)));
    const float2 db = mul2(make_float2(acc.z, acc.w), gelu_erf_grad2(make_float2(side.z, side.w)));
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(da.x, da.y), pack_bf16x2(db.x, db.y));
#else
#ifdef VB_DIAG_NO_GELU
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(acc.x * side.x, acc.y * side.y), pack_bf16x2(acc.z * side.z, acc.w * side.w));
#else
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(acc.x * gelu_erf_grad(side.x), acc.y * gelu_erf_grad(side.y)),
                                                pack_bf16x2(acc.z * gelu_erf_grad(side.z), acc.w * gelu_erf_grad(side.w)));
#endif
#endif
  } else if constexpr (EPI == VAULT_EPI_BIAS_GELU_GRAD_BF16) {
    // forward of a TRAINING step: out = gelu(x), out2 = gelu'(x) -- the derivative costs three more instructions here (it shares the
    // exponential and the tail polynomial with the value) and saves the whole evaluation in the backward epilogue, which then only
    // multiplies (VAULT_EPI_MUL_AUX_BF16): 72.7 -> 60.0 us for the 11808 x 3072 x 768 dgrad
    const float x[4] = {acc.x + b4.x, acc.y + b4.y, acc.z + b4.z, acc.w + b4.w};
    float g[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gelu_erf_both(x[i], g[i], d[i]);
    if (out2) *reinterpret_cast<uint2*>(out2) = make_uint2(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]));
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
  } else if constexpr (EPI == VAULT_EPI_MUL_AUX_BF16) {
    *reinterpret_cast<uint2*>(out) = make_uint2(pack_bf16x2(acc.x * side.x, acc.y * side.y), pack_bf16x2(acc.z * side.z, acc.w * side.w));
  } else if constexpr (EPI == VAULT_EPI_ATOMIC_F32) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
  } else if constexpr (EPI == VAULT_EPI_ATOMIC_BIAS_DROP_F32) {
    // split-K form of BIAS_RESID: dropout is a per-element scale, hence linear over the K splits (b4 is zero except in split 0)
    float x0 = acc.x + b4.x, x1 = acc.y + b4.y, x2 = acc.z + b4.z, x3 = acc.w + b4.w;
    if (p.dropout_p > 0.f) {
      const uint32_t thr = dropout_threshold(p.dropout_p);
      const float sc = 1.0f / (1.0f - p.dropout_p);
      const uint4 bits = dropout_bits4(seed, p.site, drop_idx);
      x0 = bits.x >= thr ? x0 * sc : 0.f;
      x1 = bits.y >= thr ? x1 * sc : 0.f;
      x2 = bits.z >= thr ? x2 * sc : 0.f;
      x3 = bits.w >= thr ? x3 * sc : 0.f;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
  } else if constexpr (EPI == VAULT_EPI_BIAS_F32) {
    *reinterpret_cast<float4*>(out) = make_float4(acc.x + b4.x, acc.y + b4.y, acc.z + b4.z, acc.w + b4.w);
  } else {  // VAULT_EPI_STORE_F32
    *reinterpret_cast<float4*>(out) = acc;
  }
}

// one 32 x kCW chunk: side inputs of all row groups are fetched before any store so the loads are in flight together
template <int EPI, bool GUARD>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const float4* stg, int lane, const float4& b4, unsigned long long seed,
                                               long long row0, int col) {
  const int cg = lane % kLPR, rsub = lane / kLPR;
  const long long rbase = row0 + rsub;
  if (GUARD && col >= p.N) return;
  EpiPtrs<EPI> e = make_ptrs<EPI>(p, rbase, col);
  float4 side[kNP];
  if constexpr (EPI == VAULT_EPI_BIAS_RESID_F32 || kAuxMul<EPI>) {
#pragma unroll
    for (int i = 0; i < kNP; ++i) {
      if (!GUARD || rbase + kRPP * i < p.M) side[i] = load_side<EPI>(e.side + i * e.side_step);
    }
  }
  const unsigned long long drop0 = (unsigned long long)(rbase * p.N + col) >> 2;
  const unsigned long long drop_step = (unsigned long long)p.N * (kRPP / 4);  // kRPP rows * N / 4
#pragma unroll
  for (int i = 0; i < kNP; ++i) {
    const int rl = kRPP * i + rsub;
    const float4 acc = stg[rl * kLPR + (cg ^ stg_swz(rl))];
    if (!GUARD || rbase + kRPP * i < p.M)
      epilogue4<EPI>(p, acc, b4, side[i], seed, drop0 + i * drop_step, e.out + i * e.out_step, e.out2 ? e.out2 + i * e.out2_step : nullptr);
  }
}

// Direct form (VB_EPI_DIRECT): thread = output row, kCW consecutive columns in registers; 16-byte loads / stores per thread, the
// halves of a 32-byte sector arrive from consecutive instructions of the same thread.  No staging buffer, no warp synchronisation.
template <int EPI, bool GUARD>
__device__ __forceinline__ void epilogue_row(const GemmParams& p, const uint32_t (&r)[kCW], long long row, int col0, int split, unsigned long long seed) {
  constexpr bool kOutF32 = EPI == VAULT_EPI_BIAS_RESID_F32 || EPI == VAULT_EPI_ATOMIC_F32 || EPI == VAULT_EPI_BIAS_F32 || EPI == VAULT_EPI_STORE_F32 ||
                           EPI == VAULT_EPI_ATOMIC_BIAS_DROP_F32;
  constexpr bool kBias = EPI == VAULT_EPI_BIAS_BF16 || kGeluFwd<EPI> || EPI == VAULT_EPI_BIAS_RESID_F32 || EPI == VAULT_EPI_BIAS_F32 ||
                         EPI == VAULT_EPI_ATOMIC_BIAS_DROP_F32;
  if (GUARD && row >= p.M) return;
  const bool has_bias = kBias && p.bias != nullptr && (EPI != VAULT_EPI_ATOMIC_BIAS_DROP_F32 || split == 0);
  const unsigned long long drop0 = (unsigned long long)(row * p.N + col0) >> 2;
  if constexpr (kOutF32) {
    float* out = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
    float4 side[kCW / 4];
    if constexpr (EPI == VAULT_EPI_BIAS_RESID_F32) {
      const float* rs = p.resid + row * p.ldr + col0;
#pragma unroll
      for (int j = 0; j < kCW / 4; ++j)
        if (!GUARD || col0 + 4 * j < p.N) side[j] = __ldg(reinterpret_cast<const float4*>(rs + 4 * j));
    }
#pragma unroll
    for (int j = 0; j < kCW / 4; ++j) {
      if (GUARD && col0 + 4 * j >= p.N) break;
      const float4 acc = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      const float4 b4 = has_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      epilogue4<EPI>(p, acc, b4, side[j], seed, drop0 + j, reinterpret_cast<char*>(out + 4 * j), nullptr);
    }
  } else {
    bf16* out = reinterpret_cast<bf16*>(p.out) + row * p.ldo + col0;
    bf16* out2 = (kGeluFwd<EPI> && p.out2) ? reinterpret_cast<bf16*>(p.out2) + row * p.ldo2 + col0 : nullptr;
    uint4 aux[kCW / 8];
    if constexpr (kAuxMul<EPI>) {
      const bf16* ax = p.aux + row * p.ldaux + col0;
#pragma unroll
      for (int j = 0; j < kCW / 8; ++j)
        if (!GUARD || col0 + 8 * j < p.N) aux[j] = __ldg(reinterpret_cast<const uint4*>(ax + 8 * j));
    }
#pragma unroll
    for (int j = 0; j < kCW / 8; ++j) {  // N is a multiple of 8: an 8-column group is inside or outside as a whole
      if (GUARD && col0 + 8 * j >= p.N) break;
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(r[8 * j + e]);
      if (has_bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j + 4));
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
      if constexpr (kGeluFwd<EPI>) {
        float d[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = EPI == VAULT_EPI_BIAS_GELU_GRAD_BF16 ? gelu_erf_grad(x[e]) : x[e];
        if (out2) *reinterpret_cast<uint4*>(out2 + 8 * j) = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = gelu_erf(x[e]);
      }
      if constexpr (kAuxMul<EPI>) {
        const uint32_t w[4] = {aux[j].x, aux[j].y, aux[j].z, aux[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = unpack_bf16x2(w[e]);
          x[2 * e] *= EPI == VAULT_EPI_MUL_AUX_BF16 ? a.x : gelu_erf_grad(a.x);
          x[2 * e + 1] *= EPI == VAULT_EPI_MUL_AUX_BF16 ? a.y : gelu_erf_grad(a.y);
        }
      }
      *reinterpret_cast<uint4*>(out + 8 * j) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int kAStage = BM * BK * 2;
  constexpr int kBStage = BN * BK * 2;
  constexpr int kStage = kAStage + kBStage;
  constexpr int kStages = kRingBytes / kStage;
  constexpr uint32_t kTmemCols = BN == 192 ? 512u : 2u * BN;  // two accumulator buffers; the allocation is a power of two >= 32

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring = smem_base;
  uint8_t* staging_gen = smem_gen + kRingBytes;
  const uint32_t bars = smem_base + kRingBytes + kStagingBytes;
  // barrier slots (8 bytes each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kRingBytes + kStagingBytes + 8 * (2 * kStages + 4));
  // dynamic scheduler ring: sfull[4] | sempty[4] | tile ids[4]
  auto sfull_bar = [&](int s) { return bars + 8u * (2 * kStages + 5 + s); };
  auto sempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 5 + kSchedDepth + s); };
  volatile int* tile_ring = reinterpret_cast<volatile int*>(smem_gen + kRingBytes + kStagingBytes + 8 * (2 * kStages + 5 + 2 * kSchedDepth));
  // A-operand column sums (weight-gradient launches: the bias gradient).  csfull[kStages]: the MMA thread passes on "this slot holds a k-block
  // you sum" (one arrival); csdone[kStages]: one arrival per epilogue warp once it has read the slot.  Both advance only on the k-blocks that
  // are summed and each such k-block is held in its slot until csdone completes, so neither can run more than one phase ahead of its waiters.
  constexpr bool kColsum = EPI == VAULT_EPI_ATOMIC_F32 || EPI == VAULT_EPI_STORE_F32;
  auto csfull_bar = [&](int s) { return bars + 8u * (2 * kStages + 5 + 2 * kSchedDepth + 2 + s); };
  auto csdone_bar = [&](int s) { return bars + 8u * (2 * kStages + 5 + 2 * kSchedDepth + 2 + kStages + s); };
  static_assert(8 * (2 * kStages + 5 + 2 * kSchedDepth + 2 + 2 * kStages) <= 384, "barrier block overflows its 384 bytes");
  const bool colsum = kColsum && p.a_colsum != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.cl ? 2 : 1);  // a multicast-filled slot is released by BOTH consumers
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps);
    }
    for (int s = 0; s < kSchedDepth; ++s) {
      mbar_init(sfull_bar(s), 1);
      mbar_init(sempty_bar(s), 1 + kEpiWarps);  // MMA thread + 8 epilogue warps each take every tile id the producer publishes
    }
    if constexpr (kColsum) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(csfull_bar(s), 1);
        mbar_init(csdone_bar(s), kEpiWarps);
      }
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.cl) cluster_sync_all();  // the peer's barriers must exist before a multicast load or commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_enter();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  // Work units.  Without a cluster a unit is one 128 x BN tile.  With a CTA pair a unit is two tiles that share an operand:
  // cl=1 -> (m_blk 2u+rank, same n_blk), the B tile is loaded half by each CTA and multicast to both;
  // cl=2 -> (same m_blk, n_blk 2u+rank), the A tile is shared.  Odd tails give one CTA a phantom tile: it still takes part in
  // the loads and commits, TMA zero-fills its out-of-range operand and the epilogue's bounds checks drop its stores.
  const int rank = p.cl ? (int)cluster_ctarank() : 0;
  const int unit0 = p.cl ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = p.cl ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = p.cl == 1 ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
  const int n_units = p.cl == 2 ? (p.num_n_blocks + 1) / 2 : p.num_n_blocks;
  const int tiles_mn = m_units * n_units;
  const int total_tiles = tiles_mn * p.split_k;
  // Tile sequence of this CTA: static round-robin, or (p.sched) claimed just in time by the TMA producer with an atomic
  // counter -- one tile of look-ahead, so the claim's round trip hides under the current tile's loads and no CTA hoards
  // tiles -- and broadcast to the MMA / epilogue warps through a 4-deep smem ring.  A CTA that starts late (its SM was busy
  // with another stream's kernel or a collective) then simply takes fewer tiles instead of stretching the whole launch.
  const bool dyn = p.sched != nullptr && p.cl == 0;
  struct TileIter {
    int t, step, total, slot;
    uint32_t ph;
  };
  auto iter_begin = [&]() { return TileIter{unit0 - unit_step, unit_step, total_tiles, 0, 0u}; };
  auto iter_next = [&](TileIter& it, bool arrive_lane, bool whole_warp) -> bool {
    if (!dyn) {
      it.t += it.step;
      return it.t < it.total;
    }
    mbar_wait(sfull_bar(it.slot), it.ph);
    it.t = tile_ring[it.slot];
    if (whole_warp) __syncwarp();  // every lane has read the id before lane 0 hands the slot back
    if (arrive_lane) mbar_arrive(sempty_bar(it.slot));
    if (++it.slot == kSchedDepth) { it.slot = 0; it.ph ^= 1u; }
    return it.t < it.total;
  };
  auto tile_m0 = [&](int rem) { return ((rem / n_units) * (p.cl == 1 ? 2 : 1) + (p.cl == 1 ? rank : 0)) * BM; };
  auto tile_n0 = [&](int rem) { return ((rem % n_units) * (p.cl == 2 ? 2 : 1) + (p.cl == 2 ? rank : 0)) * BN; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int t = unit0, pslot = 0;
      uint32_t pph = 0;
      uint32_t cs_pending = 0, cs_phase = 0;  // per-stage bits: slot last filled for a column-sum tile / parity of its csdone barrier
      for (;;) {
        int t_next = t + unit_step;
        if (dyn) {
          t_next = atomicAdd(p.sched, 1) + (int)gridDim.x;  // result is first read after this tile's k-loop
          mbar_wait(sempty_bar(pslot), pph ^ 1u);
          tile_ring[pslot] = t;
          mbar_arrive(sfull_bar(pslot));  // release: MMA and epilogue warps may read the id
          if (++pslot == kSchedDepth) { pslot = 0; pph ^= 1u; }
        }
        if (t >= total_tiles) break;
        const int split = t / tiles_mn;
        const int rem = t - split * tiles_mn;
        const int m0 = tile_m0(rem);
        const int n0 = tile_n0(rem);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        // the n_units tiles of one m block share its A rows: tile j sums the k-blocks with kb % n_units == j
        int cs_next = colsum ? kb0 + ((rem % n_units - kb0 % n_units) + n_units) % n_units : -1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if constexpr (kColsum) {
            // a slot that fed the column sums is also read by the epilogue warps: refill only after they have let go of it
            if ((cs_pending >> stage) & 1u) {
              mbar_wait(csdone_bar(stage), (cs_phase >> stage) & 1u);
              cs_phase ^= 1u << stage;
              cs_pending &= ~(1u << stage);
            }
            if (kb == cs_next) {
              cs_pending |= 1u << stage;
              cs_next += n_units;
            }
          }
          const uint32_t sA = ring + stage * kStage;
          const uint32_t sB = sA + kAStage;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, kStage);
          const int k0 = kb * BK;
          if (p.cl == 2) {
            // A shared along N: this CTA fetches its half (64 rows / one 64-wide m box) and multicasts it to the pair
            if (!p.a_mn) tma_load_2d_mc(sA + rank * 8192, &tmA, fb, k0, m0 + 64 * rank, (uint16_t)3);  // box {64 k, 64 rows}
            else tma_load_2d_mc(sA + rank * 8192, &tmA, fb, m0 + 64 * rank, k0, (uint16_t)3);
          } else if (!p.a_mn) {
            tma_load_2d(sA, &tmA, fb, k0, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, fb, m0 + 64 * j, k0);  // box {64 m, 64 k-rows}
          }
          if (p.cl == 1) {
            // B shared along M: half of the tile's n range per CTA, multicast to both
            if (!p.b_mn) {
              tma_load_2d_mc(sB + rank * (BN / 2) * 128, &tmB, fb, k0, n0 + rank * (BN / 2), (uint16_t)3);  // box {64 k, BN/2 rows}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = rank * (BN / 128) + j;
                tma_load_2d_mc(sB + jj * 8192, &tmB, fb, n0 + 64 * jj, k0, (uint16_t)3);
              }
            }
          } else if (!p.b_mn) {
            tma_load_2d(sB, &tmB, fb, k0, n0);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, fb, n0 + 64 * j, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        t = t_next;
      }
      if (dyn) {
        // This CTA will not touch the counter again.  The last CTA to get here re-arms both counters for the next launch
        // that uses this slot -- off the critical path, while the MMA / epilogue warps are still working.
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
          __threadfence();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(1u, BM, BN, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128u) : (UMMA_K * 2u);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      TileIter it = iter_begin();
      while (iter_next(it, true, false)) {
        const int t = it.t;
        const int split = t / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        int cs_next = -1;
        if constexpr (kColsum) {
          if (colsum) {
            const int rem = t - split * tiles_mn;
            cs_next = kb0 + ((rem % n_units - kb0 % n_units) + n_units) % n_units;
          }
        }
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if constexpr (kColsum) {
            if (kb == cs_next) {  // the k-block has landed: the epilogue warps may add it up while the MMAs below read it
              mbar_arrive(csfull_bar(stage));
              cs_next += n_units;
            }
          }
          tc_fence_after();
          const uint32_t sA = ring + stage * kStage;
          const uint32_t sB = sA + kAStage;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = umma_desc_sw128(sA + k * a_kstep, a_lbo, 1024u);
            const uint64_t bd = umma_desc_sw128(sB + k * b_kstep, b_lbo, 1024u);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (p.cl) tc_commit_mc(empty_bar(stage), (uint16_t)3);  // both CTAs' producers write into this slot of both CTAs
          else tc_commit(empty_bar(stage));               // smem slot free once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));  // accumulator complete
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int ew = warp - 4;
    const int q = warp & 3;        // TMEM lane quadrant this warp may access
    const int part = ew >> 2;      // which slice of the tile's columns
    constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
    static_assert(kColsPerWarp % kCW == 0, "tile columns must split into whole chunks per epilogue warp");
    float4* stg = reinterpret_cast<float4*>(staging_gen + ew * (32 * kCW * 4));
    const unsigned long long seed = p.seed + ((p.dropout_p > 0.f && p.seed_dev) ? *p.seed_dev : 0ull);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t gkb = 0;  // k-blocks this CTA has been through so far: ring slot = gkb % kStages, as in the producer
    uint32_t csf_phase = 0;  // per-slot parity of csfull
    TileIter it = iter_begin();
    while (iter_next(it, lane == 0, true)) {
      const int t = it.t;
      const int split = t / tiles_mn;
      const int rem = t - split * tiles_mn;
      const int m0 = tile_m0(rem);
      const int n0 = tile_n0(rem);
      if constexpr (kColsum) {
        const int kb0 = split * p.kb_per_split;
        const int nkb = min(kb0 + p.kb_per_split, p.num_k_blocks) - kb0;
        if (colsum) {
          // Bias gradient for free: while the MMAs of this tile run, the (otherwise idle) epilogue warps add up MN-major A tiles
          // (64 k-rows x two 128-byte m boxes, 128B-swizzled) along k.  The n_units tiles of an m block see the same A rows, so tile j
          // takes every n_units-th k-block (kb % n_units == j): the extra shared-memory reads are spread over all CTAs of the launch.
          // Warp -> (m box, k-row group), lane -> 2 adjacent m: a warp reads one whole 128-byte row per instruction (conflict-free
          // whatever the swizzle phase of the row).
          constexpr int kRowsPerWarp = 64 / (kEpiWarps / 2);
          const int box = ew & 1, r0 = (ew >> 1) * kRowsPerWarp;
          const int nblk = rem % n_units;
          float s0 = 0.f, s1 = 0.f;
          for (int kb = kb0 + ((nblk - kb0 % n_units) + n_units) % n_units; kb < kb0 + nkb; kb += n_units) {
            const int stage = (int)((gkb + (uint32_t)(kb - kb0)) % (uint32_t)kStages);
            mbar_wait(csfull_bar(stage), (csf_phase >> stage) & 1u);
            csf_phase ^= 1u << stage;
            const uint32_t sA = ring + stage * kStage + box * 8192;
#pragma unroll
            for (int r = r0; r < r0 + kRowsPerWarp; ++r) {
              const float2 v = unpack_bf16x2(lds_u32(sA + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4));
              s0 += v.x;
              s1 += v.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(csdone_bar(stage));
          }
          const int m = m0 + box * 64 + 2 * lane;  // M is even (checked on the host)
          if (m < p.M) {
            atomicAdd(p.a_colsum + m, s0);
            atomicAdd(p.a_colsum + m + 1, s1);
          }
        }
        gkb += (uint32_t)nkb;
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const bool full_tile = (m0 + BM <= p.M) && (n0 + BN <= p.N);
#pragma unroll 1
      for (int c = part * kColsPerWarp; c < (part + 1) * kColsPerWarp; c += kCW) {
        uint32_t r[kCW];
        tmem_ld_cols(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c), r);
        // the chunk's bias columns are requested while the accumulators are on their way from TMEM (L1 / L2 latency off the chunk's chain)
        const int col = n0 + c + (lane % kLPR) * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EPI == VAULT_EPI_BIAS_BF16 || kGeluFwd<EPI> || EPI == VAULT_EPI_BIAS_RESID_F32 || EPI == VAULT_EPI_BIAS_F32) {
          if (p.bias != nullptr && col < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        }
        if constexpr (EPI == VAULT_EPI_ATOMIC_BIAS_DROP_F32) {
          if (p.bias != nullptr && col < p.N && split == 0) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        }
        tmem_ld_wait();
#if VB_EPI_DIRECT
        if (full_tile) epilogue_row<EPI, false>(p, r, (long long)m0 + q * 32 + lane, n0 + c, split, seed);
        else epilogue_row<EPI, true>(p, r, (long long)m0 + q * 32 + lane, n0 + c, split, seed);
        continue;
#endif
        // transpose through smem: thread = row -> (kRPP rows x kLPR column-groups) per pass, XOR-swizzled 16B slots
#pragma unroll
        for (int j = 0; j < kLPR; ++j) {
          stg[lane * kLPR + (j ^ stg_swz(lane))] =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        __syncwarp();
        const long long row0 = (long long)m0 + q * 32;
        if (full_tile) epilogue_chunk<EPI, false>(p, stg, lane, b4, seed, row0, col);
        else epilogue_chunk<EPI, true>(p, stg, lane, b4, seed, row0, col);
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cl) cluster_sync_all();  // no CTA may leave while its peer can still multicast into it or arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- grouped weight-gradient launch ------------------------------------------------------------------------------------
// The four weight gradients of a transformer layer (QKV, attention output, MLP-1, MLP-2) have the same contraction (the layer's tokens)
// and nothing in common otherwise.  As four launches each pays its own head (barrier set-up, first loads: ~3 us) and tail (the last
// epilogue cannot overlap anything: most of these launches are ONE tile per CTA) -- 8-10 us on 20-50 us kernels.  Here the tiles of
// all problems form one list walked by one persistent grid: a CTA's epilogue of tile i runs under the main loop of its tile i+1
// whichever problem that belongs to, and the 216 x split-K tiles of a layer fill the 148 SMs in whole waves (432 = 2.92 waves) where the
// single launches left 4-40 SMs idle.  Same pipeline as gemm_bf16_kernel, specialised: both operands MN-major (dy and x read
// un-transposed), 128x256 tiles, fp32 red.global.add epilogue, optional fused bias gradient (a_colsum) per problem.
constexpr int kMaxGroup = 4;
struct GroupProblem {
  CUtensorMap tmA, tmB;
  int M, N;                                                     // dW is [M, N]: M = output features of the Linear, N = its input features
  int num_m_blocks, num_n_blocks, num_k_blocks, split_k, kb_per_split;
  int tile_begin;                                               // first tile of this problem in the launch's tile list
  float* out;
  long long ldo;
  float* a_colsum;
};
struct GroupedParams {
  int n, total_tiles;
  GroupProblem pr[kMaxGroup];
};

__global__ void __launch_bounds__(kThreads, 1) gemm_wgrad_grouped_kernel(const __grid_constant__ GroupedParams gp) {
  constexpr int BN = 256, EPI = VAULT_EPI_ATOMIC_F32;
  constexpr int kAStage = BM * BK * 2, kBStage = BN * BK * 2, kStage = kAStage + kBStage;
  constexpr int kStages = kRingBytes / kStage;
  constexpr uint32_t kTmemCols = 2u * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring = smem_base;
  uint8_t* staging_gen = smem_gen + kRingBytes;
  const uint32_t bars = smem_base + kRingBytes + kStagingBytes;
  // barrier slots: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem ptr | csfull[kStages] | csdone[kStages]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kRingBytes + kStagingBytes + 8 * (2 * kStages + 4));
  auto csfull_bar = [&](int s) { return bars + 8u * (2 * kStages + 5 + s); };
  auto csdone_bar = [&](int s) { return bars + 8u * (3 * kStages + 5 + s); };
  static_assert(8 * (4 * kStages + 5) <= 384, "barrier block overflows its 384 bytes");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int q = 0; q < gp.n; ++q) {
      tma_prefetch_desc(&gp.pr[q].tmA);
      tma_prefetch_desc(&gp.pr[q].tmB);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(csfull_bar(s), 1);
      mbar_init(csdone_bar(s), kEpiWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_enter();

  // tile t of the launch -> (problem, split, m block, n block): the same arithmetic in the three roles
  struct Tile {
    int q, split, m0, n0, kb0, kb1, cs_first, n_units;
  };
  auto decode = [&](int t) {
    int q = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroup; ++i)
      if (i < gp.n && t >= gp.pr[i].tile_begin) q = i;
    const GroupProblem& P = gp.pr[q];
    const int tl = t - P.tile_begin, tiles_mn = P.num_m_blocks * P.num_n_blocks;
    Tile r;
    r.q = q;
    r.split = tl / tiles_mn;
    const int rem = tl - r.split * tiles_mn;
    r.n_units = P.num_n_blocks;
    r.m0 = (rem / P.num_n_blocks) * BM;
    const int nblk = rem % P.num_n_blocks;
    r.n0 = nblk * BN;
    r.kb0 = r.split * P.kb_per_split;
    r.kb1 = min(r.kb0 + P.kb_per_split, P.num_k_blocks);
    // fused bias gradient: tile j of an m block adds up the A tiles of the k-blocks with kb % n_units == j (-1: none)
    r.cs_first = P.a_colsum != nullptr ? r.kb0 + ((nblk - r.kb0 % P.num_n_blocks) + P.num_n_blocks) % P.num_n_blocks : -1;
    return r;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, cs_pending = 0, cs_phase = 0;
      for (int t = (int)blockIdx.x; t < gp.total_tiles; t += (int)gridDim.x) {
        const Tile T = decode(t);
        const CUtensorMap* ta = &gp.pr[T.q].tmA;
        const CUtensorMap* tb = &gp.pr[T.q].tmB;
        int cs_next = T.cs_first;
        for (int kb = T.kb0; kb < T.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if ((cs_pending >> stage) & 1u) {  // the slot fed the column sums: the epilogue warps must have let go of it too
            mbar_wait(csdone_bar(stage), (cs_phase >> stage) & 1u);
            cs_phase ^= 1u << stage;
            cs_pending &= ~(1u << stage);
          }
          if (kb == cs_next) {
            cs_pending |= 1u << stage;
            cs_next += T.n_units;
          }
          const uint32_t sA = ring + stage * kStage, sB = sA + kAStage, fb = full_bar(stage);
          mbar_expect_tx(fb, kStage);
          const int k0 = kb * BK;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, ta, fb, T.m0 + 64 * j, k0);  // box {64 m, 64 k-rows}
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, tb, fb, T.n0 + 64 * j, k0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(1u, BM, BN, 1u, 1u);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = (int)blockIdx.x; t < gp.total_tiles; t += (int)gridDim.x) {
        const Tile T = decode(t);
        int cs_next = T.cs_first;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = T.kb0; kb < T.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (kb == cs_next) {
            mbar_arrive(csfull_bar(stage));
            cs_next += T.n_units;
          }
          tc_fence_after();
          const uint32_t sA = ring + stage * kStage, sB = sA + kAStage;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_f16(d_tmem, umma_desc_sw128(sA + k * (UMMA_K * 128u), 8192u, 1024u), umma_desc_sw128(sB + k * (UMMA_K * 128u), 8192u, 1024u), idesc,
                       (kb > T.kb0 || k > 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (+ the bias-gradient sums during the main loop) =====================
    const int ew = warp - 4, q4 = warp & 3, part = ew >> 2;
    constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
    float4* stg = reinterpret_cast<float4*>(staging_gen + ew * (32 * kCW * 4));
    int as = 0;
    uint32_t aphase = 0, gkb = 0, csf_phase = 0;
    for (int t = (int)blockIdx.x; t < gp.total_tiles; t += (int)gridDim.x) {
      const Tile T = decode(t);
      const GroupProblem& P = gp.pr[T.q];
      if (T.cs_first >= 0) {
        constexpr int kRowsPerWarp = 64 / (kEpiWarps / 2);
        const int box = ew & 1, r0 = (ew >> 1) * kRowsPerWarp;
        float s0 = 0.f, s1 = 0.f;
        for (int kb = T.cs_first; kb < T.kb1; kb += T.n_units) {
          const int stage = (int)((gkb + (uint32_t)(kb - T.kb0)) % (uint32_t)kStages);
          mbar_wait(csfull_bar(stage), (csf_phase >> stage) & 1u);
          csf_phase ^= 1u << stage;
          const uint32_t sA = ring + stage * kStage + box * 8192;
#pragma unroll
          for (int r = r0; r < r0 + kRowsPerWarp; ++r) {
            const float2 v = unpack_bf16x2(lds_u32(sA + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4));
            s0 += v.x;
            s1 += v.y;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(csdone_bar(stage));
        }
        const int m = T.m0 + box * 64 + 2 * lane;
        if (m < P.M) {
          atomicAdd(P.a_colsum + m, s0);
          atomicAdd(P.a_colsum + m + 1, s1);
        }
      }
      gkb += (uint32_t)(T.kb1 - T.kb0);
      GemmParams lp{};  // what the shared epilogue code reads of the problem
      lp.M = P.M; lp.N = P.N; lp.out = P.out; lp.ldo = P.ldo;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const bool full_tile = (T.m0 + BM <= P.M) && (T.n0 + BN <= P.N);
#pragma unroll 1
      for (int c = part * kColsPerWarp; c < (part + 1) * kColsPerWarp; c += kCW) {
        uint32_t r[kCW];
        tmem_ld_cols(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * BN + c), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kLPR; ++j)
          stg[lane * kLPR + (j ^ stg_swz(lane))] =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int col = T.n0 + c + (lane % kLPR) * 4;
        const float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long row0 = (long long)T.m0 + q4 * 32;
        if (full_tile) epilogue_chunk<EPI, false>(lp, stg, lane, b4, 0ull, row0, col);
        else epilogue_chunk<EPI, true>(lp, stg, lane, b4, 0ull, row0, col);
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D bf16 tensor [outer rows, inner contiguous], 128B swizzle, zero fill out of bounds
int encode_tmap_2d_sw(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle);
int encode_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d_sw(tm, dt, elem_bytes, base, inner, outer, ld_elems, box_inner, box_outer, CU_TENSOR_MAP_SWIZZLE_128B);
}
int encode_tmap_2d_sw(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VAULT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * elem_bytes) & 15) != 0)
    return fail(VAULT_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, ld %llu)", base, (unsigned long long)ld_elems);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VAULT_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu ld %llu box %ux%u", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
  return VAULT_OK;
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%d): %s", kSmemBytes, cudaGetErrorString(e));
    attr_set = true;
  }
  if (p.cl) launch_cluster(gemm_bf16_kernel<BN, EPI>, dim3(grid), dim3(kThreads), kSmemBytes, st, 2u, tmA, tmB, p);
  else launch(gemm_bf16_kernel<BN, EPI>, dim3(grid), dim3(kThreads), kSmemBytes, st, tmA, tmB, p);
  return check_launch("gemm_bf16_kernel");
}

template <int EPI>
int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t st) {
  switch (bn) {
    case 256: return launch_gemm<256, EPI>(tmA, tmB, p, grid, st);
    case 192: return launch_gemm<192, EPI>(tmA, tmB, p, grid, st);
    case 128: return launch_gemm<128, EPI>(tmA, tmB, p, grid, st);
    case 64: return launch_gemm<64, EPI>(tmA, tmB, p, grid, st);
  }
  return fail(VAULT_ERR_INVALID, "block_n must be 64, 128, 192 or 256 (got %d)", bn);
}

// Tile-N choice: fill whole waves of SMs; wider tiles amortise the A-tile smem traffic (N=256 runs the MMA at full rate with
// 96 B/clk of smem reads, N=128 needs 128 B/clk, N=64 is smem-bound).  N=192 exists for the 768-wide outputs at 4,096 tokens (the LM stack
// of the target shape): 32 x 4 = 128 tiles fill 86 % of the SMs in one wave where 128x256 tiles (96) fill 65 %.
static bool no_bn192() {  // VAULT_B200_NO_BN192=1: A/B switch
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VAULT_B200_NO_BN192");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}
static int pick_block_n(int M, int N, int split_k, int sms) {
  const int cands[4] = {256, 192, 128, 64};
  const float rate[4] = {1.0f, 0.93f, 0.85f, 0.55f};
  float best = -1.f;
  int best_bn = 128;
  const int mb = (M + BM - 1) / BM;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N < bn) continue;
    if (bn == 192 && (N % 192 != 0 || no_bn192())) continue;
    const int nb = (N + bn - 1) / bn;
    const long long tiles = (long long)mb * nb * split_k;
    const long long waves = (tiles + sms - 1) / sms;
    const float util = (float)((double)M * N * split_k / ((double)waves * sms * BM * bn));
    const float score = util * rate[i];
    if (score > best) { best = score; best_bn = bn; }
  }
  return best_bn;
}

}  // namespace vb

extern "C" int vault_gemm_bf16(const vault_gemm_args* a, void* stream) {
  using namespace vb;
  VB_REQUIRE(a != nullptr, "vault_gemm_bf16: null args");
  VB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "vault_gemm_bf16: empty problem %dx%dx%d", a->M, a->N, a->K);
  VB_REQUIRE(a->N % 8 == 0, "vault_gemm_bf16: N=%d must be a multiple of 8", a->N);
  VB_REQUIRE(a->A && a->B && a->out, "vault_gemm_bf16: null operand");
  VB_REQUIRE(a->split_k >= 1, "vault_gemm_bf16: split_k must be >= 1");
  VB_REQUIRE(a->split_k == 1 || a->epilogue == VAULT_EPI_ATOMIC_F32 || a->epilogue == VAULT_EPI_ATOMIC_BIAS_DROP_F32,
             "vault_gemm_bf16: split_k>1 needs an ATOMIC epilogue");
  if (a->epilogue == VAULT_EPI_BIAS_RESID_F32) VB_REQUIRE(a->resid != nullptr, "vault_gemm_bf16: resid required");
  if (a->epilogue == VAULT_EPI_DGELU_BF16 || a->epilogue == VAULT_EPI_MUL_AUX_BF16) VB_REQUIRE(a->aux != nullptr, "vault_gemm_bf16: aux required");
  VB_REQUIRE(a->ldo % 4 == 0, "vault_gemm_bf16: ldo=%lld must be a multiple of 4", (long long)a->ldo);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int sms = a->max_ctas > 0 ? a->max_ctas : device_sm_count();
  int bn = a->block_n > 0 ? a->block_n : pick_block_n(a->M, a->N, a->split_k, sms);

  int cl = a->cluster > 0 ? a->cluster : 0;  // 0 / negative: no cluster (the CTA-pair multicast measured +2 % at best, so there is no automatic choice)
  VB_REQUIRE(cl <= 2, "vault_gemm_bf16: cluster mode %d (0/-1 off, 1 pair along M, 2 pair along N)", a->cluster);
  if (cl == 1 && a->b_mn && bn < 128) cl = 0;  // an MN-major B tile of one 64-wide box cannot be split across the pair
  if (bn == 192) cl = 0;                       // three 64-wide boxes do not split evenly across a CTA pair
  if (a->a_colsum) {
    VB_REQUIRE(a->a_mn == 1 && a->M % 2 == 0, "vault_gemm_bf16: a_colsum needs an MN-major A operand and an even M");
    VB_REQUIRE(a->epilogue == VAULT_EPI_ATOMIC_F32 || a->epilogue == VAULT_EPI_STORE_F32, "vault_gemm_bf16: a_colsum needs EPI_ATOMIC_F32 / EPI_STORE_F32");
    cl = 0;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = encode_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, cl == 2 ? 64 : BM);
  else rc = encode_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
  if (rc) return rc;
  if (!a->b_mn) rc = encode_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK,
                                    (uint32_t)(cl == 1 ? bn / 2 : bn));
  else rc = encode_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
  if (rc) return rc;

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.num_m_blocks = (a->M + BM - 1) / BM;
  p.num_n_blocks = (a->N + bn - 1) / bn;
  p.num_k_blocks = (a->K + BK - 1) / BK;
  p.split_k = a->split_k < p.num_k_blocks ? a->split_k : p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + p.split_k - 1) / p.split_k;
  p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.bias = a->bias; p.resid = a->resid; p.ldr = a->ldr;
  p.aux = reinterpret_cast<const bf16*>(a->aux); p.ldaux = a->ldaux;
  p.out = a->out; p.ldo = a->ldo; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.dropout_p = a->dropout_p; p.seed = a->seed; p.seed_dev = reinterpret_cast<const unsigned long long*>(a->seed_dev); p.site = a->site;
  p.a_colsum = a->a_colsum;
  p.cl = cl;
  p.sched = cl ? nullptr : a->sched;
  int grid;
  if (cl) {
    const long long mu = cl == 1 ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks, nu = cl == 2 ? (p.num_n_blocks + 1) / 2 : p.num_n_blocks;
    const long long units = mu * nu * p.split_k;
    const long long pairs = sms / 2;
    grid = 2 * (int)(units < pairs ? units : pairs);
  } else {
    const long long total = (long long)p.num_m_blocks * p.num_n_blocks * p.split_k;
    grid = (int)(total < sms ? total : sms);
  }

  switch (a->epilogue) {
    case VAULT_EPI_BIAS_BF16: return dispatch_bn<VAULT_EPI_BIAS_BF16>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_BIAS_GELU_BF16: return dispatch_bn<VAULT_EPI_BIAS_GELU_BF16>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_BIAS_RESID_F32: return dispatch_bn<VAULT_EPI_BIAS_RESID_F32>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_PLAIN_BF16: return dispatch_bn<VAULT_EPI_PLAIN_BF16>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_DGELU_BF16: return dispatch_bn<VAULT_EPI_DGELU_BF16>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_ATOMIC_F32: return dispatch_bn<VAULT_EPI_ATOMIC_F32>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_BIAS_F32: return dispatch_bn<VAULT_EPI_BIAS_F32>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_STORE_F32: return dispatch_bn<VAULT_EPI_STORE_F32>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_ATOMIC_BIAS_DROP_F32: return dispatch_bn<VAULT_EPI_ATOMIC_BIAS_DROP_F32>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_BIAS_GELU_GRAD_BF16: return dispatch_bn<VAULT_EPI_BIAS_GELU_GRAD_BF16>(bn, tmA, tmB, p, grid, st);
    case VAULT_EPI_MUL_AUX_BF16: return dispatch_bn<VAULT_EPI_MUL_AUX_BF16>(bn, tmA, tmB, p, grid, st);
  }
  return fail(VAULT_ERR_INVALID, "vault_gemm_bf16: unknown epilogue %d", a->epilogue);
}

extern "C" int vault_gemm_wgrad_grouped(const vault_gemm_args* args, int32_t n, void* stream) {
  using namespace vb;
  VB_REQUIRE(args != nullptr && n >= 1 && n <= kMaxGroup, "vault_gemm_wgrad_grouped: 1..%d problems (got %d)", kMaxGroup, n);
  GroupedParams gp;  // ~1.4 KB of kernel parameters (the tensor maps travel as __grid_constant__ data)
  gp.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const vault_gemm_args* a = args + i;
    VB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->A && a->B && a->out, "vault_gemm_wgrad_grouped: problem %d: empty or null", i);
    VB_REQUIRE(a->a_mn == 1 && a->b_mn == 1 && a->epilogue == VAULT_EPI_ATOMIC_F32,
               "vault_gemm_wgrad_grouped: problem %d: needs MN-major operands and EPI_ATOMIC_F32 (the weight-gradient form)", i);
    VB_REQUIRE(a->N % 8 == 0 && a->M % 2 == 0 && a->ldo % 4 == 0 && a->split_k >= 1, "vault_gemm_wgrad_grouped: problem %d: bad shape", i);
    GroupProblem& P = gp.pr[i];
    int rc = encode_tmap_2d(&P.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
    if (rc) return rc;
    rc = encode_tmap_2d(&P.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
    if (rc) return rc;
    P.M = a->M; P.N = a->N;
    P.num_m_blocks = (a->M + BM - 1) / BM;
    P.num_n_blocks = (a->N + 255) / 256;
    P.num_k_blocks = (a->K + BK - 1) / BK;
    int split = a->split_k < P.num_k_blocks ? a->split_k : P.num_k_blocks;
    P.kb_per_split = (P.num_k_blocks + split - 1) / split;
    P.split_k = (P.num_k_blocks + P.kb_per_split - 1) / P.kb_per_split;
    P.tile_begin = tiles;
    tiles += P.num_m_blocks * P.num_n_blocks * P.split_k;
    P.out = reinterpret_cast<float*>(a->out); P.ldo = a->ldo; P.a_colsum = a->a_colsum;
  }
  gp.total_tiles = tiles;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%d): %s", kSmemBytes, cudaGetErrorString(e));
    attr_set = true;
  }
  const int sms = args[0].max_ctas > 0 ? args[0].max_ctas : device_sm_count();
  launch(gemm_wgrad_grouped_kernel, dim3((unsigned)(tiles < sms ? tiles : sms)), dim3(kThreads), kSmemBytes, reinterpret_cast<cudaStream_t>(stream), gp);
  return check_launch("gemm_wgrad_grouped_kernel");
}
