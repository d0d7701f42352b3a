// Patch embedding as an im2col-free GEMM: Conv2d(C, N, kernel = stride = P = 32) == [B*gh*gw, C*P*P] x [N, C*P*P]^T.
// The A tiles are never materialised: a 5-D TMA tensor map over the NCHW fp32 pixels
//     (kw:32 | pw:gw | kh:32 | ph:gh | b*C+c)   box (32, gw, 1, phBox, nb images of one channel via elementStride C)
// drops, for one (c, kh), the 32 contiguous kw of up to 128 patches straight into 128-byte swizzled smem rows -- exactly the
// K-major operand tcgen05 wants for a 32-deep k-block.  The pixels stay fp32: the MMA runs as kind::tf32 (K=8 per
// instruction, fp32 weights), which is also more accurate than rounding raw pixels to bf16.
// One tile per CTA (96 k-blocks each, so prologue/epilogue are amortised), warp roles as in gemm_sm100.cu.
#include "common.cuh"

namespace vb {

int encode_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer);  // gemm_sm100.cu
int encode_tmap_2d_sw(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

constexpr int kPeBN = 128;
constexpr int kPeBK = 32;  // fp32 elements = 128 bytes
constexpr int kPeStage = 128 * 128 + kPeBN * 128;
constexpr int kPeStages = 6;
constexpr int kPeThreads = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM, warp3 idle, warps 4-7 epilogue
constexpr int kPeSmem = kPeStages * kPeStage + 4 * 4096 + 1024 + 256;

struct PatchParams {
  int B, C, gh, gw, N;
  int phBox, nb, rows_per_tile;
  int tiles_ph, tiles_b, tiles_n;
  int num_k_blocks;  // C * 32
  const float* bias;
  float* out;
};

__global__ void __launch_bounds__(kPeThreads, 1)
patch_embed_tf32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const PatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring = smem_base;
  uint8_t* staging_gen = smem_gen + kPeStages * kPeStage;
  const uint32_t bars = smem_base + kPeStages * kPeStage + 4 * 4096;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kPeStages + s); };
  const uint32_t acc_bar = bars + 8u * (2 * kPeStages);
  const uint32_t tmem_slot = bars + 8u * (2 * kPeStages + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kPeStages * kPeStage + 4 * 4096 + 8 * (2 * kPeStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPeStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kPeBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_enter();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  // tile -> (n block, image group, patch-row group)
  const int t = blockIdx.x;
  const int n_blk = t % p.tiles_n;
  const int rem = t / p.tiles_n;
  const int ph0 = (rem % p.tiles_ph) * p.phBox;
  const int b0 = (rem / p.tiles_ph) * p.nb;
  const int n0 = n_blk * kPeBN;
  const uint32_t a_bytes = (uint32_t)p.rows_per_tile * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sA = ring + stage * kPeStage, sB = sA + 128 * 128, fb = full_bar(stage);
        mbar_expect_tx(fb, a_bytes + kPeBN * 128);
        const int c = kb >> 5, kh = kb & 31;
        tma_load_5d(sA, &tmX, fb, 0, 0, kh, ph0, b0 * p.C + c);
        tma_load_2d(sB, &tmW, fb, kb * kPeBK, n0);
        if (++stage == kPeStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(2u /*tf32*/, 128, kPeBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sA = ring + stage * kPeStage, sB = sA + 128 * 128;
#pragma unroll
        for (int k = 0; k < kPeBK / 8; ++k) {
          const uint64_t ad = umma_desc_sw128(sA + k * 32, 16u, 1024u);
          const uint64_t bd = umma_desc_sw128(sB + k * 32, 16u, 1024u);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(stage));
        if (++stage == kPeStages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(acc_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    float4* stg = reinterpret_cast<float4*>(staging_gen + q * 4096);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int per_img = p.phBox * p.gw;
#pragma unroll 1
    for (int c = 0; c < kPeBN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        stg[lane * 8 + (j ^ (lane & 7))] =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      __syncwarp();
      const int cg = lane & 7;
      const int col = n0 + c + cg * 4;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && col < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + (lane >> 3);
        const int rt = q * 32 + rl;  // row of the tile = (image, patch row, patch col) in box order
        const int ib = rt / per_img, rr = rt % per_img;
        const int b = b0 + ib;
        if (rt < p.rows_per_tile && b < p.B && col < p.N) {
          const float4 a = stg[rl * 8 + (cg ^ (rl & 7))];
          const long long orow = ((long long)b * p.gh + ph0 + rr / p.gw) * p.gw + rr % p.gw;
          *reinterpret_cast<float4*>(p.out + orow * p.N + col) = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPeBN);
  }
}


// ---- weight gradient, im2col-free ------------------------------------------------------------------------------------------------
//   dW[n, k] += sum_m dY[m, n] * X[m, k]      m = patch (b, ph, pw),  k = (c, kh, kw),  dY fp32 [B*gh*gw, N],  X = the NCHW fp32 pixels
// Both operands are read as they lie in memory, contraction over the patches:
//   B operand: the forward's 5-D pixel map (here with the 32-byte-atom flavour of the 128-byte swizzle, the one an MN-major operand of
//              4-byte elements needs), box (32 kw, gw, 1 kh, pb patch rows, 1 image): one box = [rows = pb*gw patches] x [32 kw]
//              in 128-byte swizzled rows = one 32-wide N atom of an MN-major operand (N = kw contiguous, K = patches); the 8 boxes of
//              kh0 .. kh0+7 sit `rows * 128` bytes apart (the descriptor's leading byte offset) and form N = 256.
//   A operand: dY rows of the same patches, box (32 n, rows) x 4 = M = 128 output channels, MN-major as well (n contiguous, K = patches).
// tcgen05.mma kind::tf32, M = 128, N = 256, K = 8 patches per instruction; one CTA = one (n block, channel, 8-kh group) tile over a slice of the
// patches (split over the contraction so that the launch fills the SMs), partial sums added into the zero-filled fp32 gradient with
// red.global.add.v4.f32.  No patch matrix is ever materialised.
constexpr int kWgThreads = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM, warp3 idle, warps 4-7 epilogue
constexpr int kWgSmemRing = 196608;
constexpr int kWgSmem = kWgSmemRing + 4 * 4096 + 1024 + 256;

struct PatchWgParams {
  int B, C, gh, gw, N, K;
  int pb, rows, rows_pad, stages, stage_bytes;
  int kblocks_total, kb_per_split, split;
  float* dW;
};

__global__ void __launch_bounds__(kWgThreads, 1)
patch_embed_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const PatchWgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring = smem_base;
  uint8_t* staging_gen = smem_gen + kWgSmemRing;
  const uint32_t bars = smem_base + kWgSmemRing + 4 * 4096;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (8 + s); };
  const uint32_t acc_bar = bars + 8u * 16;
  const uint32_t tmem_slot = bars + 8u * 17;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kWgSmemRing + 4 * 4096 + 8 * 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256u);
  const uint32_t slab = (uint32_t)p.rows_pad * 128u;  // one 32-wide atom column: rows_pad x 128 B (a multiple of 1024)
  if (p.rows_pad > p.rows) {
    // k-blocks of `rows` patches are padded to a multiple of the 8-patch MMA depth: the pad rows of every slab are never written by
    // TMA (its boxes hold exactly `rows` rows), so zero-filling them once makes their products vanish for the whole kernel
    const int pad16 = (p.rows_pad - p.rows) * 8;  // 16-byte pieces per slab
    const int nslab = p.stages * 12;
    for (int i = threadIdx.x; i < nslab * pad16; i += kWgThreads) {
      const uint32_t a = ring + (uint32_t)(i / pad16) * slab + (uint32_t)p.rows * 128u + (uint32_t)(i % pad16) * 16u;
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_enter();

  // CTA -> (split slice, kh group, channel, n block)
  int t = blockIdx.x;
  const int n_blk = t % (p.N / 128); t /= (p.N / 128);
  const int khg = t % 4; t /= 4;
  const int c = t % p.C; t /= p.C;
  const int sp = t;
  const int kb0 = sp * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kblocks_total);
  const int n0 = n_blk * 128, kh0 = khg * 8;
  const int per_img = p.gh / p.pb;                // k-blocks per image

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sA = ring + (uint32_t)stage * (uint32_t)p.stage_bytes, sB = sA + 4u * slab, fb = full_bar(stage);
        mbar_expect_tx(fb, 12u * (uint32_t)p.rows * 128u);
        const int b = kb / per_img, ph0 = (kb % per_img) * p.pb;
        const int m0 = (b * p.gh + ph0) * p.gw;  // first patch row of this k-block in dY
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(sA + (uint32_t)j * slab, &tmDY, fb, n0 + 32 * j, m0);
#pragma unroll
        for (int j = 0; j < 8; ++j) tma_load_5d(sB + (uint32_t)j * slab, &tmX, fb, 0, 0, kh0 + j, ph0, b * p.C + c);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(2u /*tf32*/, 128, 256, 1u, 1u);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sA = ring + (uint32_t)stage * (uint32_t)p.stage_bytes, sB = sA + 4u * slab;
        for (int k = 0; k < p.rows_pad / 8; ++k) {  // 8 patches (= 8 smem rows = 1024 B) per instruction
          const uint64_t ad = umma_desc_sw128_base32(sA + (uint32_t)k * 1024u, slab, 512u);
          const uint64_t bd = umma_desc_sw128_base32(sB + (uint32_t)k * 1024u, slab, 512u);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(acc_bar);
    }
  } else if (warp >= 4 && kb1 > kb0) {
    const int q = warp & 3;
    float4* stg = reinterpret_cast<float4*>(staging_gen + q * 4096);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const long long kcol0 = (long long)c * 1024 + kh0 * 32;  // 256 consecutive k of this tile: (kh0 .. kh0+7) x 32 kw
#pragma unroll 1
    for (int cc = 0; cc < 256; cc += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        stg[lane * 8 + (j ^ (lane & 7))] =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      __syncwarp();
      const int cg = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + (lane >> 3);
        const int n = n0 + q * 32 + rl;
        if (n < p.N) {
          const float4 a = stg[rl * 8 + (cg ^ (rl & 7))];
          float* dst = p.dW + (long long)n * p.K + kcol0 + cc + cg * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

// fp32 patch-gradient rows for the wgrad above + their column sums (the projection's bias gradient):
//   dpatch[b, i*gw+j] = dX[b, T+1+i*w_b+j] inside the sample's valid rectangle, zero elsewhere;  dbias[n] += sum_rows dpatch[., n]
__global__ void __launch_bounds__(256) patch_grad_rows_f32_kernel(const float* __restrict__ dX, const int* __restrict__ hw, float* __restrict__ dpatch,
                                                                  float* __restrict__ dbias, int B, int T, int Pmax, int gh, int gw, int H, int rows_per_cta) {
  pdl_enter();
  const int S = T + 1 + Pmax;
  const long long cells = (long long)B * gh * gw;
  const long long cell0 = (long long)blockIdx.x * rows_per_cta;
  for (int c4 = threadIdx.x; c4 < H / 4; c4 += 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rows_per_cta; ++r) {
      const long long cell = cell0 + r;
      if (cell >= cells) break;
      const int b = (int)(cell / (gh * gw)), ij = (int)(cell % (gh * gw));
      const int i = ij / gw, j = ij % gw;
      const int h = hw[2 * b], w = hw[2 * b + 1];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < h && j < w) v = __ldg(reinterpret_cast<const float4*>(dX + ((long long)b * S + T + 1 + i * w + j) * H) + c4);
      reinterpret_cast<float4*>(dpatch + cell * H)[c4] = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (dbias != nullptr) {
      atomicAdd(dbias + 4 * c4, acc.x); atomicAdd(dbias + 4 * c4 + 1, acc.y); atomicAdd(dbias + 4 * c4 + 2, acc.z); atomicAdd(dbias + 4 * c4 + 3, acc.w);
    }
  }
}

}  // namespace vb

extern "C" int vault_patch_embed_fwd(const float* pixels, const float* weight, const float* bias, float* out, int32_t B, int32_t C, int32_t Hi,
                                     int32_t Wi, int32_t P, int32_t N, void* stream) {
  using namespace vb;
  VB_REQUIRE(pixels && weight && out, "patch_embed_fwd: null pointer");
  VB_REQUIRE(P == 32, "patch_embed_fwd: patch size %d (the 128-byte swizzle row is one 32-float patch row)", P);
  VB_REQUIRE(B > 0 && C > 0 && Hi % P == 0 && Wi % P == 0 && N % 4 == 0, "patch_embed_fwd: bad shape B=%d C=%d %dx%d N=%d", B, C, Hi, Wi, N);
  const int gh = Hi / P, gw = Wi / P;
  VB_REQUIRE(gw <= 128, "patch_embed_fwd: image too wide (%d patches per row)", gw);
  // rows per tile = nb * phBox * gw <= 128 with phBox | gh: take the fullest tile
  int best_rows = 0, phBox = 1, nb = 1;
  for (int pb = 1; pb <= gh; ++pb) {
    if (gh % pb || pb * gw > 128) continue;
    int n = 128 / (pb * gw);
    if (n > B) n = B;
    if ((n - 1) * C + 1 > 256) n = 255 / C + 1;
    if (n * pb * gw > best_rows) { best_rows = n * pb * gw; phBox = pb; nb = n; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VAULT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(pixels) & 15) == 0 && (Wi * 4) % 16 == 0, "patch_embed_fwd: pixels must be 16-byte aligned");
  CUtensorMap tmX, tmW;
  {
    cuuint64_t gdim[5] = {32, (cuuint64_t)gw, 32, (cuuint64_t)gh, (cuuint64_t)B * C};
    cuuint64_t gstr[4] = {32ull * 4, (cuuint64_t)Wi * 4, 32ull * Wi * 4, (cuuint64_t)Hi * Wi * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)gw, 1, (cuuint32_t)phBox, (cuuint32_t)((nb - 1) * C + 1)};
    cuuint32_t estr[5] = {1, 1, 1, 1, (cuuint32_t)C};
    VB_REQUIRE(C <= 8, "patch_embed_fwd: more than 8 channels not supported (TMA element stride)");
    CUresult r = fn(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(pixels), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VAULT_ERR_DRIVER, "patch_embed_fwd: pixel tensor map encode failed (%d)", (int)r);
  }
  const int K = C * P * P;
  int rc = encode_tmap_2d(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, weight, (uint64_t)K, (uint64_t)N, (uint64_t)K, kPeBK, kPeBN);
  if (rc) return rc;
  PatchParams p;
  p.B = B; p.C = C; p.gh = gh; p.gw = gw; p.N = N;
  p.phBox = phBox; p.nb = nb; p.rows_per_tile = nb * phBox * gw;
  p.tiles_ph = gh / phBox; p.tiles_b = (B + nb - 1) / nb; p.tiles_n = (N + kPeBN - 1) / kPeBN;
  p.num_k_blocks = C * 32;
  p.bias = bias; p.out = out;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmem);
    if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "patch_embed_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = p.tiles_ph * p.tiles_b * p.tiles_n;
  launch(patch_embed_tf32_kernel, dim3(grid), dim3(kPeThreads), kPeSmem, (cudaStream_t)stream, tmX, tmW, p);
  return check_launch("patch_embed_tf32_kernel");
}

extern "C" int vault_patch_grad_rows_f32(const float* dX, const int32_t* hw, float* dpatch_f32, float* dbias, int32_t B, int32_t T, int32_t Pmax,
                                         int32_t gh, int32_t gw, int32_t H, void* stream) {
  using namespace vb;
  VB_REQUIRE(dX && hw && dpatch_f32, "patch_grad_rows_f32: null pointer");
  VB_REQUIRE(H % 4 == 0, "patch_grad_rows_f32: H=%d must be a multiple of 4", H);
  const long long cells = (long long)B * gh * gw;
  const int rows_per_cta = 16;
  launch(patch_grad_rows_f32_kernel, dim3((unsigned)((cells + rows_per_cta - 1) / rows_per_cta)), dim3(256), 0, (cudaStream_t)stream, dX, hw, dpatch_f32, dbias, B,
         T, Pmax, gh, gw, H, rows_per_cta);
  return check_launch("patch_grad_rows_f32_kernel");
}

/* Patch rows per k-block of the wgrad kernel for a gh x gw patch grid: the largest pb | gh with pb*gw <= 64 patches (the k-block is
 * padded to a multiple of 8 with zero rows in shared memory); 0 if a single patch row already exceeds 64 patches. */
static int patch_wgrad_pb(int gh, int gw) {
  int best = 0;
  for (int pb = 1; pb <= gh; ++pb) {
    if (gh % pb || pb * gw > 64) continue;
    best = pb;
  }
  return best;
}

extern "C" int vault_patch_embed_wgrad_ok(int32_t C, int32_t Hi, int32_t Wi, int32_t P, int32_t N) {
  if (P != 32 || C < 1 || C > 8 || Hi % 32 || Wi % 32 || N % 128) return 0;
  return patch_wgrad_pb(Hi / 32, Wi / 32) > 0 ? 1 : 0;
}

extern "C" int vault_patch_embed_wgrad(const float* pixels, const float* dpatch_f32, float* dW, int32_t B, int32_t C, int32_t Hi, int32_t Wi, int32_t P,
                                       int32_t N, void* stream) {
  using namespace vb;
  VB_REQUIRE(pixels && dpatch_f32 && dW, "patch_embed_wgrad: null pointer");
  VB_REQUIRE(vault_patch_embed_wgrad_ok(C, Hi, Wi, P, N), "patch_embed_wgrad: unsupported shape C=%d %dx%d P=%d N=%d", C, Hi, Wi, P, N);
  const int gh = Hi / P, gw = Wi / P;
  const int pb = patch_wgrad_pb(gh, gw), rows = pb * gw;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(VAULT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(pixels) & 15) == 0 && (reinterpret_cast<uintptr_t>(dpatch_f32) & 15) == 0, "patch_embed_wgrad: operands must be 16-byte aligned");
  CUtensorMap tmX, tmDY;
  {
    cuuint64_t gdim[5] = {32, (cuuint64_t)gw, 32, (cuuint64_t)gh, (cuuint64_t)B * C};
    cuuint64_t gstr[4] = {32ull * 4, (cuuint64_t)Wi * 4, 32ull * Wi * 4, (cuuint64_t)Hi * Wi * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)gw, 1, (cuuint32_t)pb, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(pixels), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VAULT_ERR_DRIVER, "patch_embed_wgrad: pixel tensor map encode failed (%d)", (int)r);
  }
  int rc = encode_tmap_2d_sw(&tmDY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dpatch_f32, (uint64_t)N, (uint64_t)B * gh * gw, (uint64_t)N, 32, (uint32_t)rows,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  PatchWgParams p;
  p.B = B; p.C = C; p.gh = gh; p.gw = gw; p.N = N; p.K = C * P * P;
  p.pb = pb; p.rows = rows; p.rows_pad = (rows + 7) / 8 * 8;
  p.stage_bytes = 12 * p.rows_pad * 128;
  p.stages = kWgSmemRing / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.kblocks_total = B * (gh / pb);
  const int tiles = (N / 128) * C * 4;
  int split = device_sm_count() / tiles;
  if (split < 1) split = 1;
  if (split > p.kblocks_total) split = p.kblocks_total;
  p.kb_per_split = (p.kblocks_total + split - 1) / split;
  p.split = (p.kblocks_total + p.kb_per_split - 1) / p.kb_per_split;
  p.dW = dW;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "patch_embed_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch(patch_embed_wgrad_tf32_kernel, dim3(tiles * p.split), dim3(kWgThreads), kWgSmem, (cudaStream_t)stream, tmX, tmDY, p);
  return check_launch("patch_embed_wgrad_tf32_kernel");
}
