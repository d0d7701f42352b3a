// Embedding assembly kernels (all HBM-bound, one warp per token row, 128-bit accesses):
//   LM word+type+position gather (BERT / RoBERTa position ids), ViLT text type(+pos) add, per-sample patch grid from the pixel
//   mask, ViLT sequence assembly (raster-ordered valid patches + bilinear position table + modality embeddings, writing text and
//   image rows straight into one [B,S,H] buffer -- no concat copy), their backward scatters, and the bf16 im2col of the pixels.
#include "common.cuh"

namespace vb {

constexpr int kEmbWarps = 8;

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void atomic_add4(float* dst, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// RoBERTa: pid = cumsum(ids != pad)[t] * (ids[t] != pad) + pad   (HF:models/roberta/modeling_roberta.py:152-170); warp-collective
__device__ __forceinline__ int roberta_pos(const int64_t* ids_row, int t, int pad, int lane) {
  int cnt = 0;
  for (int base = 0; base <= t; base += 32) {
    const int i = base + lane;
    const bool nz = (i <= t) && (ids_row[i] != pad);
    cnt += __popc(__ballot_sync(0xffffffffu, nz));
  }
  return ids_row[t] != pad ? cnt + pad : pad;
}

__global__ void __launch_bounds__(kEmbWarps * 32)
lm_embed_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt, const float* __restrict__ word, const float* __restrict__ type,
                    const float* __restrict__ pos, float* __restrict__ x, const int64_t* __restrict__ attn_mask, uint8_t* __restrict__ key_mask, int B,
                    int T, int H, int roberta_pad) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kEmbWarps + warp;
  if (row >= (long long)B * T) return;
  if (key_mask && lane == 0) key_mask[row] = attn_mask ? (attn_mask[row] != 0) : 1;
  const int b = (int)(row / T), t = (int)(row % T);
  const long long id = ids[row];
  const long long ty = tt ? tt[row] : 0;
  const int pid = roberta_pad >= 0 ? roberta_pos(ids + (long long)b * T, t, roberta_pad, lane) : t;
  const float4* w = reinterpret_cast<const float4*>(word + id * H);
  const float4* ty4 = reinterpret_cast<const float4*>(type + ty * H);
  const float4* p4 = reinterpret_cast<const float4*>(pos + (long long)pid * H);
  float4* o = reinterpret_cast<float4*>(x + row * H);
  for (int c = lane; c < H / 4; c += 32) o[c] = f4add(f4add(__ldg(w + c), __ldg(ty4 + c)), __ldg(p4 + c));
}

// Token-type rows are shared by (nearly) every token: thousands of atomics on the same 1-2 table rows serialise in the L2 (the kernel took
// 86 us for 4,096 rows).  A warp therefore walks several rows and keeps the sums of types 0 and 1 in registers; the CTA's eight warps meet in
// shared memory and issue ONE red.global.add.v4 per 4 columns and type.  Word / position rows see little contention and keep direct atomics.
constexpr int kEmbMaxV = 8;  // float4 per lane: H <= 1024 on the register path (wider rows fall back to direct atomics)

template <int NT>
__device__ __forceinline__ void type_sums_flush(float4 (&acc)[NT][kEmbMaxV], float* __restrict__ dtype, int H, int warp, int lane) {
  __shared__ float4 red[kEmbWarps][kEmbMaxV * 32];  // 32 KB: one token type at a time
  const int nv = H / 4;
#pragma unroll
  for (int ty = 0; ty < NT; ++ty) {
#pragma unroll
    for (int k = 0; k < kEmbMaxV; ++k)
      if (lane + 32 * k < nv) red[warp][lane + 32 * k] = acc[ty][k];
    __syncthreads();
    for (int c = threadIdx.x; c < nv; c += kEmbWarps * 32) {
      float4 a = red[0][c];
#pragma unroll
      for (int w = 1; w < kEmbWarps; ++w) a = f4add(a, red[w][c]);
      if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) atomic_add4(dtype + (long long)ty * H + 4 * c, a);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kEmbWarps * 32)
lm_embed_bwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt, const float* __restrict__ dx, float* __restrict__ dword,
                    float* __restrict__ dtype, float* __restrict__ dpos, int B, int T, int H, int roberta_pad, int word_pad) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool reg_path = dtype != nullptr && H <= kEmbMaxV * 128;
  float4 acc[2][kEmbMaxV];
#pragma unroll
  for (int k = 0; k < kEmbMaxV; ++k) acc[0][k] = acc[1][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < (long long)B * T; row += (long long)gridDim.x * kEmbWarps) {
    const int b = (int)(row / T), t = (int)(row % T);
    const long long id = ids[row];
    const long long ty = tt ? tt[row] : 0;
    const int pid = roberta_pad >= 0 ? roberta_pos(ids + (long long)b * T, t, roberta_pad, lane) : t;
    const float4* g = reinterpret_cast<const float4*>(dx + row * H);
    // nn.Embedding(padding_idx=...) rows receive no gradient (word table: pad_token_id; RoBERTa position table: pad id too)
    const bool do_word = dword != nullptr && id != word_pad;
    const bool do_pos = dpos != nullptr && !(roberta_pad >= 0 && pid == roberta_pad);
    const bool ty_reg = reg_path && ty < 2;  // warp-uniform
#pragma unroll
    for (int k = 0; k < kEmbMaxV; ++k) {
      const int c = lane + 32 * k;
      if (c < H / 4) {
        const float4 v = __ldg(g + c);
        if (do_word) atomic_add4(dword + id * H + 4 * c, v);
        if (do_pos) atomic_add4(dpos + (long long)pid * H + 4 * c, v);
        if (ty_reg) {
          if (ty == 0) acc[0][k] = f4add(acc[0][k], v);
          else acc[1][k] = f4add(acc[1][k], v);
        } else if (dtype) {
          atomic_add4(dtype + ty * H + 4 * c, v);
        }
      }
    }
    if (H > kEmbMaxV * 128) {  // columns beyond the register path
      for (int c = lane + 32 * kEmbMaxV; c < H / 4; c += 32) {
        const float4 v = __ldg(g + c);
        if (do_word) atomic_add4(dword + id * H + 4 * c, v);
        if (do_pos) atomic_add4(dpos + (long long)pid * H + 4 * c, v);
        if (dtype) atomic_add4(dtype + ty * H + 4 * c, v);
      }
    }
  }
  if (reg_path) type_sums_flush<2>(acc, dtype, H, warp, lane);
}

__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_text_embed_fwd_kernel(const float* __restrict__ e, const int64_t* __restrict__ tt, const float* __restrict__ type, const float* __restrict__ pos,
                           float* __restrict__ x, int B, int T, int H) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kEmbWarps + warp;
  if (row >= (long long)B * T) return;
  const int t = (int)(row % T);
  const long long ty = tt ? tt[row] : 0;
  const float4* e4 = reinterpret_cast<const float4*>(e + row * H);
  const float4* ty4 = reinterpret_cast<const float4*>(type + ty * H);
  float4* o = reinterpret_cast<float4*>(x + row * H);
  for (int c = lane; c < H / 4; c += 32) {
    float4 v = f4add(__ldg(e4 + c), __ldg(ty4 + c));
    if (pos) v = f4add(v, __ldg(reinterpret_cast<const float4*>(pos + (long long)t * H) + c));
    o[c] = v;
  }
}

__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_text_embed_bwd_kernel(const int64_t* __restrict__ tt, const float* __restrict__ dx, float* __restrict__ dtype, float* __restrict__ dpos, int B,
                           int T, int H) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool reg_path = dtype != nullptr && H <= kEmbMaxV * 128;
  float4 acc[2][kEmbMaxV];
#pragma unroll
  for (int k = 0; k < kEmbMaxV; ++k) acc[0][k] = acc[1][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < (long long)B * T; row += (long long)gridDim.x * kEmbWarps) {
    const int t = (int)(row % T);
    const long long ty = tt ? tt[row] : 0;
    const float4* g = reinterpret_cast<const float4*>(dx + row * H);
    const bool ty_reg = reg_path && ty < 2;  // warp-uniform
    for (int c = lane; c < H / 4; c += 32) {
      const float4 v = __ldg(g + c);
      const int k = c >> 5;
      if (ty_reg && k < kEmbMaxV) {
#pragma unroll
        for (int kk = 0; kk < kEmbMaxV; ++kk)
          if (kk == k) {
            if (ty == 0) acc[0][kk] = f4add(acc[0][kk], v);
            else acc[1][kk] = f4add(acc[1][kk], v);
          }
      } else if (dtype) {
        atomic_add4(dtype + ty * H + 4 * c, v);
      }
      if (dpos) atomic_add4(dpos + (long long)t * H + 4 * c, v);
    }
  }
  if (reg_path) type_sums_flush<2>(acc, dtype, H, warp, lane);
}

// h_b = #valid patch rows in patch column 0, w_b = #valid patch columns in patch row 0 (nearest down-sampling: pixel (P*i, P*j))
template <typename MT>
__global__ void patch_grid_kernel(const MT* __restrict__ mask, int* __restrict__ hw, int Hi, int Wi, int P) {
  pdl_enter();
  const int b = blockIdx.x;
  const MT* m = mask + (long long)b * Hi * Wi;
  int h = 0, w = 0;
  for (int i = threadIdx.x; i < Hi / P; i += blockDim.x) h += m[(long long)i * P * Wi] != (MT)0;
  for (int j = threadIdx.x; j < Wi / P; j += blockDim.x) w += m[(long long)j * P] != (MT)0;
  h = (int)warp_sum((float)h);
  w = (int)warp_sum((float)w);
  if (threadIdx.x == 0) { hw[2 * b] = h; hw[2 * b + 1] = w; }
}

struct Bilin {
  int i00, i01, i10, i11;  // rows of pos_table (already offset by 1)
  float w00, w01, w10, w11;
};
// F.interpolate(mode="bilinear", align_corners=True) from grid x grid to (h, w): source = dst * (grid-1)/(out-1)
__device__ __forceinline__ Bilin bilin_coeffs(int i, int j, int h, int w, int grid) {
  const float sy = h > 1 ? (float)(grid - 1) / (float)(h - 1) : 0.f;
  const float sx = w > 1 ? (float)(grid - 1) / (float)(w - 1) : 0.f;
  const float fy = sy * i, fx = sx * j;
  const int y0 = min((int)fy, grid - 1), x0 = min((int)fx, grid - 1);
  const int y1 = min(y0 + 1, grid - 1), x1 = min(x0 + 1, grid - 1);
  const float ly = fy - y0, lx = fx - x0;
  Bilin r;
  r.i00 = 1 + y0 * grid + x0; r.i01 = 1 + y0 * grid + x1; r.i10 = 1 + y1 * grid + x0; r.i11 = 1 + y1 * grid + x1;
  r.w00 = (1.f - ly) * (1.f - lx); r.w01 = (1.f - ly) * lx; r.w10 = ly * (1.f - lx); r.w11 = ly * lx;
  return r;
}

__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_assemble_fwd_kernel(const float* __restrict__ text_ln, const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos_table,
                         const float* __restrict__ modality, const int64_t* __restrict__ attn_mask, const int* __restrict__ hw, float* __restrict__ X,
                         uint8_t* __restrict__ key_mask, int B, int T, int Pmax, int gh, int gw, int grid, int H, int img_type) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = T + 1 + Pmax;
  const long long row = (long long)blockIdx.x * kEmbWarps + warp;
  if (row >= (long long)B * S) return;
  const int b = (int)(row / S), s = (int)(row % S);
  float4* o = reinterpret_cast<float4*>(X + row * H);
  const int H4 = H / 4;
  if (s < T) {
    const float4* src = reinterpret_cast<const float4*>(text_ln + ((long long)b * T + s) * H);
    const float4* m4 = reinterpret_cast<const float4*>(modality);
    for (int c = lane; c < H4; c += 32) o[c] = f4add(__ldg(src + c), __ldg(m4 + c));
    if (lane == 0) key_mask[row] = attn_mask ? (attn_mask[(long long)b * T + s] != 0) : 1;
    return;
  }
  const float4* m4 = reinterpret_cast<const float4*>(modality + (long long)img_type * H);
  if (s == T) {
    const float4* c4 = reinterpret_cast<const float4*>(cls);
    const float4* p0 = reinterpret_cast<const float4*>(pos_table);
    for (int c = lane; c < H4; c += 32) o[c] = f4add(f4add(__ldg(c4 + c), __ldg(p0 + c)), __ldg(m4 + c));
    if (lane == 0) key_mask[row] = 1;
    return;
  }
  const int p = s - T - 1;
  const int h = hw[2 * b], w = hw[2 * b + 1];
  if (p >= h * w) {
    for (int c = lane; c < H4; c += 32) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) key_mask[row] = 0;
    return;
  }
  const int i = p / w, j = p % w;
  const Bilin bl = bilin_coeffs(i, j, h, w, grid);
  const float4* src = reinterpret_cast<const float4*>(patch + ((long long)b * gh * gw + (long long)i * gw + j) * H);
  const float4* t00 = reinterpret_cast<const float4*>(pos_table + (long long)bl.i00 * H);
  const float4* t01 = reinterpret_cast<const float4*>(pos_table + (long long)bl.i01 * H);
  const float4* t10 = reinterpret_cast<const float4*>(pos_table + (long long)bl.i10 * H);
  const float4* t11 = reinterpret_cast<const float4*>(pos_table + (long long)bl.i11 * H);
  for (int c = lane; c < H4; c += 32) {
    float4 pe = f4scale(__ldg(t00 + c), bl.w00);
    pe = f4add(pe, f4scale(__ldg(t01 + c), bl.w01));
    pe = f4add(pe, f4scale(__ldg(t10 + c), bl.w10));
    pe = f4add(pe, f4scale(__ldg(t11 + c), bl.w11));
    o[c] = f4add(f4add(__ldg(src + c), pe), __ldg(m4 + c));
  }
  if (lane == 0) key_mask[row] = 1;
}

// Backward of the assembly, pass 1 (row-parallel): text rows -> dtext_ln; CLS / patch rows -> dcls, dpos_table scatter;
// modality sums kept in registers per warp, reduced per CTA.  H <= 1024.
__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_assemble_bwd_rows_kernel(const float* __restrict__ dX, const int* __restrict__ hw, float* __restrict__ dtext_ln, float* __restrict__ dcls,
                              float* __restrict__ dpos_table, float* __restrict__ dmodality, int B, int T, int Pmax, int grid, int H, int img_type) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = T + 1 + Pmax;
  const int H4 = H / 4;
  float4 acc_t[8], acc_i[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc_t[k] = acc_i[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < (long long)B * S; row += (long long)gridDim.x * kEmbWarps) {
    const int b = (int)(row / S), s = (int)(row % S);
    const float4* g = reinterpret_cast<const float4*>(dX + row * H);
    if (s < T) {
      float4* o = dtext_ln ? reinterpret_cast<float4*>(dtext_ln + ((long long)b * T + s) * H) : nullptr;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = lane + 32 * k;
        if (c < H4) {
          const float4 v = __ldg(g + c);
          if (o) o[c] = v;
          acc_t[k] = f4add(acc_t[k], v);
        }
      }
    } else if (s == T) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = lane + 32 * k;
        if (c < H4) {
          const float4 v = __ldg(g + c);
          acc_i[k] = f4add(acc_i[k], v);
          if (dcls) atomic_add4(dcls + 4 * c, v);
          if (dpos_table) atomic_add4(dpos_table + 4 * c, v);
        }
      }
    } else {
      const int p = s - T - 1;
      const int h = hw[2 * b], w = hw[2 * b + 1];
      if (p >= h * w) continue;
      const Bilin bl = bilin_coeffs(p / w, p % w, h, w, grid);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = lane + 32 * k;
        if (c < H4) {
          const float4 v = __ldg(g + c);
          acc_i[k] = f4add(acc_i[k], v);
          if (dpos_table) {
            if (bl.w00 != 0.f) atomic_add4(dpos_table + (long long)bl.i00 * H + 4 * c, f4scale(v, bl.w00));
            if (bl.w01 != 0.f) atomic_add4(dpos_table + (long long)bl.i01 * H + 4 * c, f4scale(v, bl.w01));
            if (bl.w10 != 0.f) atomic_add4(dpos_table + (long long)bl.i10 * H + 4 * c, f4scale(v, bl.w10));
            if (bl.w11 != 0.f) atomic_add4(dpos_table + (long long)bl.i11 * H + 4 * c, f4scale(v, bl.w11));
          }
        }
      }
    }
  }
  if (dmodality == nullptr) return;
  __shared__ float4 red[kEmbWarps][256 + 1];
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][lane + 32 * k] = pass == 0 ? acc_t[k] : acc_i[k];
    __syncthreads();
    float* dst = dmodality + (long long)(pass == 0 ? 0 : img_type) * H;
    for (int c = threadIdx.x; c < H4; c += kEmbWarps * 32) {
      float4 a = red[0][c];
#pragma unroll
      for (int wv = 1; wv < kEmbWarps; ++wv) a = f4add(a, red[wv][c]);
      atomic_add4(dst + 4 * c, a);
    }
  }
}

// pass 2 (grid-cell-parallel): dpatch[b, i*gw+j] = bf16(dX[b, T+1+i*w_b+j]) if (i,j) inside the sample's valid rectangle else 0
__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_assemble_bwd_patch_kernel(const float* __restrict__ dX, const int* __restrict__ hw, bf16* __restrict__ dpatch, int B, int T, int Pmax, int gh, int gw,
                               int H) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = T + 1 + Pmax;
  const long long cell = (long long)blockIdx.x * kEmbWarps + warp;
  if (cell >= (long long)B * gh * gw) return;
  const int b = (int)(cell / (gh * gw)), ij = (int)(cell % (gh * gw));
  const int i = ij / gw, j = ij % gw;
  const int h = hw[2 * b], w = hw[2 * b + 1];
  uint2* o = reinterpret_cast<uint2*>(dpatch + cell * H);
  if (i < h && j < w) {
    const float4* g = reinterpret_cast<const float4*>(dX + ((long long)b * S + T + 1 + i * w + j) * H);
    for (int c = lane; c < H / 4; c += 32) {
      const float4 v = __ldg(g + c);
      o[c] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  } else {
    for (int c = lane; c < H / 4; c += 32) o[c] = make_uint2(0u, 0u);
  }
}

// ---- pre-embedded image tokens (HF ViltEmbeddings.forward with image_embeds=..., HF:models/vilt/modeling_vilt.py:196-201; the TomViLT
// path ref:vault/models/tomvilt/model.py:281-287): X = [text_ln + modality[0] | image_embeds + modality[img_type]], no CLS, no position
// table, key validity = [attention_mask | image_mask].
__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_assemble_embeds_fwd_kernel(const float* __restrict__ text_ln, const float* __restrict__ img, const float* __restrict__ modality,
                                const int64_t* __restrict__ attn_mask, const uint8_t* __restrict__ img_mask, float* __restrict__ X,
                                uint8_t* __restrict__ key_mask, int B, int T, int P, int H, int img_type) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = T + P;
  const long long row = (long long)blockIdx.x * kEmbWarps + warp;
  if (row >= (long long)B * S) return;
  const int b = (int)(row / S), s = (int)(row % S);
  float4* o = reinterpret_cast<float4*>(X + row * H);
  const bool text = s < T;
  const float4* src = reinterpret_cast<const float4*>(text ? text_ln + ((long long)b * T + s) * H : img + ((long long)b * P + (s - T)) * H);
  const float4* m4 = reinterpret_cast<const float4*>(modality + (long long)(text ? 0 : img_type) * H);
  for (int c = lane; c < H / 4; c += 32) o[c] = f4add(__ldg(src + c), __ldg(m4 + c));
  if (lane == 0) {
    if (text) key_mask[row] = attn_mask ? (attn_mask[(long long)b * T + s] != 0) : 1;
    else key_mask[row] = img_mask ? (img_mask[(long long)b * P + (s - T)] != 0) : 1;
  }
}

__global__ void __launch_bounds__(kEmbWarps * 32)
vilt_assemble_embeds_bwd_kernel(const float* __restrict__ dX, float* __restrict__ dtext_ln, float* __restrict__ dimg, float* __restrict__ dmodality, int B,
                                int T, int P, int H, int img_type) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = T + P, H4 = H / 4;
  float4 acc_t[8], acc_i[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc_t[k] = acc_i[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < (long long)B * S; row += (long long)gridDim.x * kEmbWarps) {
    const int b = (int)(row / S), s = (int)(row % S);
    const float4* g = reinterpret_cast<const float4*>(dX + row * H);
    const bool text = s < T;
    float* dst = text ? dtext_ln : dimg;
    float4* o = dst ? reinterpret_cast<float4*>(dst + (text ? (long long)b * T + s : (long long)b * P + (s - T)) * H) : nullptr;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      if (c < H4) {
        const float4 v = __ldg(g + c);
        if (o) o[c] = v;
        if (text) acc_t[k] = f4add(acc_t[k], v);
        else acc_i[k] = f4add(acc_i[k], v);
      }
    }
  }
  if (dmodality == nullptr) return;
  __shared__ float4 red[kEmbWarps][256 + 1];
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][lane + 32 * k] = pass == 0 ? acc_t[k] : acc_i[k];
    __syncthreads();
    float* dst = dmodality + (long long)(pass == 0 ? 0 : img_type) * H;
    for (int c = threadIdx.x; c < H4; c += kEmbWarps * 32) {
      float4 a = red[0][c];
#pragma unroll
      for (int wv = 1; wv < kEmbWarps; ++wv) a = f4add(a, red[wv][c]);
      atomic_add4(dst + 4 * c, a);
    }
  }
}

// im2col: out[(b*gh+i)*gw+j, c*P*P + kh*P + kw] = bf16(pixels[b,c,i*P+kh,j*P+kw]); one thread = 8 consecutive kw
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ px, bf16* __restrict__ out, int B, int C, int Hi, int Wi, int P) {
  pdl_enter();
  const int gh = Hi / P, gw = Wi / P;
  const int K = C * P * P;
  const long long total = (long long)B * gh * gw * (K / 8);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(idx % (K / 8));
    const long long rowi = idx / (K / 8);
    const int k = k8 * 8;
    const int c = k / (P * P), kh = (k / P) % P, kw = k % P;
    const int j = (int)(rowi % gw), i = (int)((rowi / gw) % gh), b = (int)(rowi / ((long long)gw * gh));
    const float* src = px + (((long long)b * C + c) * Hi + (i * P + kh)) * Wi + j * P + kw;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(d.x, d.y); o.w = pack_bf16x2(d.z, d.w);
    *reinterpret_cast<uint4*>(out + rowi * K + k) = o;
  }
}

static inline unsigned rows_grid(long long rows) { return (unsigned)((rows + kEmbWarps - 1) / kEmbWarps); }
// embedding backward: at most one CTA per SM, so the per-CTA token-type sums meet in <= 148 atomics per address
static inline unsigned bwd_rows_grid(long long rows) {
  const long long g = (rows + kEmbWarps - 1) / kEmbWarps, cap = device_sm_count();
  return (unsigned)(g < cap ? g : cap);
}

}  // namespace vb

using namespace vb;

extern "C" int vault_lm_embed_fwd(const int64_t* ids, const int64_t* tt, const float* word, const float* type, const float* pos, float* x_sum,
                                  const int64_t* attention_mask, uint8_t* key_mask, int32_t B, int32_t T, int32_t H, int32_t roberta_pad,
                                  void* stream) {
  VB_REQUIRE(ids && word && type && pos && x_sum, "lm_embed_fwd: null pointer");
  VB_REQUIRE(B > 0 && T > 0 && H % 4 == 0, "lm_embed_fwd: bad shape");
  launch(lm_embed_fwd_kernel, dim3(rows_grid((long long)B * T)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, ids, tt, word, type, pos, x_sum, attention_mask, key_mask, B, T, H,
                                                                                                roberta_pad);
  return check_launch("lm_embed_fwd_kernel");
}

extern "C" int vault_lm_embed_bwd(const int64_t* ids, const int64_t* tt, const float* dx, float* dword, float* dtype, float* dpos, int32_t B,
                                  int32_t T, int32_t H, int32_t roberta_pad, int32_t word_pad, void* stream) {
  VB_REQUIRE(ids && dx, "lm_embed_bwd: null pointer");
  VB_REQUIRE(B > 0 && T > 0 && H % 4 == 0, "lm_embed_bwd: bad shape");
  launch(lm_embed_bwd_kernel, dim3(bwd_rows_grid((long long)B * T)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, ids, tt, dx, dword, dtype, dpos, B, T, H, roberta_pad,
                                                                                                word_pad);
  return check_launch("lm_embed_bwd_kernel");
}

extern "C" int vault_vilt_text_embed_fwd(const float* inputs_embeds, const int64_t* tt, const float* type, const float* pos, float* x_sum,
                                         int32_t B, int32_t T, int32_t H, void* stream) {
  VB_REQUIRE(inputs_embeds && type && x_sum, "vilt_text_embed_fwd: null pointer");
  VB_REQUIRE(B > 0 && T > 0 && H % 4 == 0, "vilt_text_embed_fwd: bad shape");
  launch(vilt_text_embed_fwd_kernel, dim3(rows_grid((long long)B * T)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, inputs_embeds, tt, type, pos, x_sum, B, T, H);
  return check_launch("vilt_text_embed_fwd_kernel");
}

extern "C" int vault_vilt_text_embed_bwd(const int64_t* tt, const float* dx, float* dtype, float* dpos, int32_t B, int32_t T, int32_t H,
                                         void* stream) {
  VB_REQUIRE(dx, "vilt_text_embed_bwd: null pointer");
  if (!dtype && !dpos) return VAULT_OK;
  launch(vilt_text_embed_bwd_kernel, dim3(bwd_rows_grid((long long)B * T)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, tt, dx, dtype, dpos, B, T, H);
  return check_launch("vilt_text_embed_bwd_kernel");
}

extern "C" int vault_patch_grid(const void* pixel_mask, int32_t mask_is_f32, int32_t* hw, int32_t B, int32_t Hi, int32_t Wi, int32_t P,
                                void* stream) {
  VB_REQUIRE(pixel_mask && hw && B > 0 && Hi % P == 0 && Wi % P == 0, "patch_grid: bad arguments (Hi=%d Wi=%d P=%d)", Hi, Wi, P);
  if (mask_is_f32) launch(patch_grid_kernel<float>, dim3(B), dim3(32), 0, (cudaStream_t)stream, reinterpret_cast<const float*>(pixel_mask), hw, Hi, Wi, P);
  else launch(patch_grid_kernel<int64_t>, dim3(B), dim3(32), 0, (cudaStream_t)stream, reinterpret_cast<const int64_t*>(pixel_mask), hw, Hi, Wi, P);
  return check_launch("patch_grid_kernel");
}

extern "C" int vault_vilt_assemble_fwd(const float* text_ln, const float* patch, const float* cls, const float* pos_table, const float* modality,
                                       const int64_t* attention_mask, const int32_t* hw, float* X, uint8_t* key_mask, int32_t B, int32_t T,
                                       int32_t Pmax, int32_t gh, int32_t gw, int32_t grid, int32_t H, int32_t img_type, void* stream) {
  VB_REQUIRE(text_ln && patch && cls && pos_table && modality && hw && X && key_mask, "vilt_assemble_fwd: null pointer");
  VB_REQUIRE(H % 4 == 0 && Pmax <= gh * gw && Pmax >= 0, "vilt_assemble_fwd: bad shape");
  const long long rows = (long long)B * (T + 1 + Pmax);
  launch(vilt_assemble_fwd_kernel, dim3(rows_grid(rows)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, text_ln, patch, cls, pos_table, modality, attention_mask, hw, X,
                                                                                         key_mask, B, T, Pmax, gh, gw, grid, H, img_type);
  return check_launch("vilt_assemble_fwd_kernel");
}

extern "C" int vault_vilt_assemble_bwd(const float* dX, const int32_t* hw, float* dtext_ln, void* dpatch_bf16, float* dcls, float* dpos_table,
                                       float* dmodality, int32_t B, int32_t T, int32_t Pmax, int32_t gh, int32_t gw, int32_t grid, int32_t H,
                                       int32_t img_type, void* stream) {
  VB_REQUIRE(dX && hw, "vilt_assemble_bwd: null pointer");
  VB_REQUIRE(H % 4 == 0 && H <= 1024, "vilt_assemble_bwd: H=%d must be a multiple of 4 and <= 1024", H);
  const long long rows = (long long)B * (T + 1 + Pmax);
  long long grid1 = rows_grid(rows);
  const long long cap = (long long)device_sm_count() * 2;
  if (grid1 > cap) grid1 = cap;
  launch(vilt_assemble_bwd_rows_kernel, dim3((unsigned)grid1), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, dX, hw, dtext_ln, dcls, dpos_table, dmodality, B, T, Pmax,
                                                                                              grid, H, img_type);
  int rc = check_launch("vilt_assemble_bwd_rows_kernel");
  if (rc) return rc;
  if (dpatch_bf16) {
    launch(vilt_assemble_bwd_patch_kernel, dim3(rows_grid((long long)B * gh * gw)), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, 
        dX, hw, reinterpret_cast<bf16*>(dpatch_bf16), B, T, Pmax, gh, gw, H);
    rc = check_launch("vilt_assemble_bwd_patch_kernel");
  }
  return rc;
}

extern "C" int vault_vilt_assemble_embeds_fwd(const float* text_ln, const float* image_embeds, const float* modality, const int64_t* attention_mask,
                                              const uint8_t* image_mask, float* X, uint8_t* key_mask, int32_t B, int32_t T, int32_t P, int32_t H,
                                              int32_t img_type, void* stream) {
  VB_REQUIRE(text_ln && image_embeds && modality && X && key_mask, "vilt_assemble_embeds_fwd: null pointer");
  VB_REQUIRE(H % 4 == 0 && P > 0 && T > 0, "vilt_assemble_embeds_fwd: bad shape");
  launch(vilt_assemble_embeds_fwd_kernel, dim3(rows_grid((long long)B * (T + P))), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, text_ln, image_embeds,
         modality, attention_mask, image_mask, X, key_mask, B, T, P, H, img_type);
  return check_launch("vilt_assemble_embeds_fwd_kernel");
}

extern "C" int vault_vilt_assemble_embeds_bwd(const float* dX, float* dtext_ln, float* dimage_embeds, float* dmodality, int32_t B, int32_t T, int32_t P,
                                              int32_t H, int32_t img_type, void* stream) {
  VB_REQUIRE(dX, "vilt_assemble_embeds_bwd: null pointer");
  VB_REQUIRE(H % 4 == 0 && H <= 1024, "vilt_assemble_embeds_bwd: H=%d must be a multiple of 4 and <= 1024", H);
  long long grid1 = rows_grid((long long)B * (T + P));
  const long long cap = (long long)device_sm_count() * 2;
  if (grid1 > cap) grid1 = cap;
  launch(vilt_assemble_embeds_bwd_kernel, dim3((unsigned)grid1), dim3(kEmbWarps * 32), 0, (cudaStream_t)stream, dX, dtext_ln, dimage_embeds, dmodality, B, T,
         P, H, img_type);
  return check_launch("vilt_assemble_embeds_bwd_kernel");
}

extern "C" int vault_patchify_bf16(const float* pixels, void* out_bf16, int32_t B, int32_t C, int32_t Hi, int32_t Wi, int32_t P, void* stream) {
  VB_REQUIRE(pixels && out_bf16, "patchify: null pointer");
  VB_REQUIRE(P % 8 == 0 && Hi % P == 0 && Wi % P == 0 && Wi % 4 == 0, "patchify: bad shape Hi=%d Wi=%d P=%d", Hi, Wi, P);
  const long long total = (long long)B * (Hi / P) * (Wi / P) * (C * P * P / 8);
  long long grid = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (grid > cap) grid = cap;
  launch(patchify_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, pixels, reinterpret_cast<bf16*>(out_bf16), B, C, Hi, Wi, P);
  return check_launch("patchify_kernel");
}
