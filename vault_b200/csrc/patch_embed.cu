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
