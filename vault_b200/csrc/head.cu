// Pooler / classifier head / loss / bias-gradient kernels.  Everything here is tiny (B rows) and runs in fp32 on CUDA cores:
// a [B,768]x[768,768] pooler is 38 MFLOP -- not tensor-core work.
#include "common.cuh"

namespace vb {

// y[r,n] = act(sum_k x[r*ldx+k] * W[n*K+k] + b[n]); one warp per output
__global__ void __launch_bounds__(256)
small_linear_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ y,
                        int rows, int N, int K, int act) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * 8 + warp;
  if (o >= (long long)rows * N) return;
  const int r = (int)(o / N), n = (int)(o % N);
  const float4* xr = reinterpret_cast<const float4*>(x + r * ldx);
  const float4* wr = reinterpret_cast<const float4*>(W + (long long)n * K);
  float s = 0.f;
  for (int c = lane; c < K / 4; c += 32) {
    const float4 a = __ldg(xr + c), w = __ldg(wr + c);
    s += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
  }
  s = warp_sum(s);
  if (lane == 0) {
    s += b ? b[n] : 0.f;
    y[(long long)r * N + n] = act == 1 ? tanhf(s) : s;
  }
}

__device__ __forceinline__ float act_grad(float dy, const float* y, long long idx, int act) {
  if (act == 1) {
    const float t = y[idx];
    return dy * (1.f - t * t);
  }
  return dy;
}

// dW[n,k] = sum_r g[r,n] x[r,k], db[n] = sum_r g[r,n]; one warp per (n, 128-column chunk), rows unrolled for memory parallelism
__global__ void __launch_bounds__(256)
small_linear_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, long long ldx, float* __restrict__ dW,
                          float* __restrict__ db, int rows, int N, int K, int act) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (K / 4 + 31) / 32;
  const long long o = (long long)blockIdx.x * 8 + warp;
  if (o >= (long long)N * chunks) return;
  const int n = (int)(o / chunks), ch = (int)(o % chunks);
  const int c = ch * 32 + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float g[8];
    float4 a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + j;
      g[j] = r < rows ? act_grad(dy[(long long)r * N + n], y, (long long)r * N + n, act) : 0.f;
      a[j] = (r < rows && c < K / 4) ? __ldg(reinterpret_cast<const float4*>(x + r * ldx) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bsum += g[j];
      acc.x += g[j] * a[j].x; acc.y += g[j] * a[j].y; acc.z += g[j] * a[j].z; acc.w += g[j] * a[j].w;
    }
  }
  if (dW && c < K / 4) reinterpret_cast<float4*>(dW + (long long)n * K)[c] = acc;
  if (db && ch == 0 && lane == 0) db[n] = bsum;
}

// dx[r,k] (+)= sum_n g[r,n] W[n,k]; one CTA per (r, 128-column chunk): its 8 warps split n, partial sums meet in shared memory
__global__ void __launch_bounds__(256)
small_linear_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ W, float* __restrict__ dx, long long lddx,
                          int accumulate, int rows, int N, int K, int act) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (K / 4 + 31) / 32;
  const int r = blockIdx.x / chunks, c = (blockIdx.x % chunks) * 32 + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n0 = warp; n0 < N; n0 += 8 * 4) {
    float g[4];
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 8 * j;
      g[j] = n < N ? act_grad(dy[(long long)r * N + n], y, (long long)r * N + n, act) : 0.f;
      w[j] = (n < N && c < K / 4) ? __ldg(reinterpret_cast<const float4*>(W + (long long)n * K) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc.x += g[j] * w[j].x; acc.y += g[j] * w[j].y; acc.z += g[j] * w[j].z; acc.w += g[j] * w[j].w;
    }
  }
  __shared__ float4 red[8][33];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < K / 4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int wv = 1; wv < 8; ++wv) {
      const float4 u = red[wv][lane];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float4* d = reinterpret_cast<float4*>(dx + r * lddx) + c;
    if (accumulate) {
      const float4 old = *d;
      t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
    }
    *d = t;
  }
}

__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p, unsigned long long seed,
                                   const unsigned long long* seed_dev, unsigned site) {
  pdl_enter();
  if (seed_dev) seed += *seed_dev;
  const uint32_t thr = dropout_threshold(p);
  const float sc = 1.f / (1.f - p);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (long long)gridDim.x * blockDim.x) {
    const uint4 bits = dropout_bits4(seed, site, (unsigned long long)q);
    const uint32_t bb[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = q * 4 + e;
      if (i < n) y[i] = bb[e] >= thr ? x[i] * sc : 0.f;
    }
  }
}

// mean softmax cross-entropy over rows (nn.CrossEntropyLoss default reduction); one block, thread per row (strided)
__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ loss, float* __restrict__ dlogits, int rows,
               int C, float grad_scale) {
  pdl_enter();
  __shared__ float red[8];
  float local = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* z = logits + (long long)r * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - m);
    const float lse = m + logf(se);
    const int lab = (int)labels[r];
    local += lse - z[lab];
    if (dlogits) {
      for (int c = 0; c < C; ++c) dlogits[(long long)r * C + c] = (expf(z[c] - lse) - (c == lab ? 1.f : 0.f)) * grad_scale / rows;
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    loss[0] = t / rows;
  }
}

// The other two losses of the reference's trainers, same contract (mean over rows, dlogits = d loss / d logits * grad_scale):
//   kind 1: nn.BCEWithLogitsLoss on logits [rows] vs float targets [rows]            (ref:vault/models/vault/trainer.py:42-56, n_classes = 1)
//   kind 2: 0.5 * (CE(logits[:, :C/2], labels[:, 0]) + CE(logits[:, C/2:], labels[:, 1]))   (MVSA raw annotations, ref :114-137)
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ logits, const void* __restrict__ labels, float* __restrict__ loss, float* __restrict__ dlogits, int rows,
                 int C, int kind, float grad_scale) {
  pdl_enter();
  __shared__ float red[8];
  float local = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    if (kind == 1) {
      const float x = logits[r], y = reinterpret_cast<const float*>(labels)[r];
      local += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
      if (dlogits) dlogits[r] = (1.f / (1.f + expf(-x)) - y) * grad_scale / rows;
    } else {
      const int G = C / 2;
      for (int g = 0; g < 2; ++g) {
        const float* z = logits + (long long)r * C + g * G;
        float m = -INFINITY;
        for (int c = 0; c < G; ++c) m = fmaxf(m, z[c]);
        float se = 0.f;
        for (int c = 0; c < G; ++c) se += expf(z[c] - m);
        const float lse = m + logf(se);
        const int lab = (int)reinterpret_cast<const int64_t*>(labels)[2LL * r + g];
        local += 0.5f * (lse - z[lab]);
        if (dlogits) {
          for (int c = 0; c < G; ++c)
            dlogits[(long long)r * C + g * G + c] = 0.5f * (expf(z[c] - lse) - (c == lab ? 1.f : 0.f)) * grad_scale / rows;
        }
      }
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    loss[0] = t / rows;
  }
}

// out[c] += sum_r x[r, c]  (bf16 in, fp32 atomics out); lane = 4 columns, warp = 128 columns, 8 warps stride rows
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out, long long rows, int cols) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 128 + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(x + r * ldx + c));
      const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
      acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
    }
  }
  __shared__ float4 red[8][33];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
    float4 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 u = red[w][lane];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    atomicAdd(out + c, t.x); atomicAdd(out + c + 1, t.y); atomicAdd(out + c + 2, t.z); atomicAdd(out + c + 3, t.w);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vault_small_linear_fwd(const float* x, int64_t ldx, const float* W, const float* b, float* y, int32_t rows, int32_t N, int32_t K,
                                      int32_t act, void* stream) {
  VB_REQUIRE(x && W && y, "small_linear_fwd: null pointer");
  VB_REQUIRE(rows > 0 && N > 0 && K > 0 && K % 4 == 0 && ldx % 4 == 0, "small_linear_fwd: bad shape rows=%d N=%d K=%d", rows, N, K);
  const long long outs = (long long)rows * N;
  launch(small_linear_fwd_kernel, dim3((unsigned)((outs + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, x, ldx, W, b, y, rows, N, K, act);
  return check_launch("small_linear_fwd_kernel");
}

extern "C" int vault_small_linear_bwd(const float* dy, const float* y, const float* x, int64_t ldx, const float* W, float* dx, int64_t lddx,
                                      int32_t accumulate_dx, float* dW, float* db, int32_t rows, int32_t N, int32_t K, int32_t act, void* stream) {
  VB_REQUIRE(dy && x && W, "small_linear_bwd: null pointer");
  VB_REQUIRE(act == 0 || y != nullptr, "small_linear_bwd: tanh backward needs the saved output");
  VB_REQUIRE(rows > 0 && N > 0 && K > 0 && K % 4 == 0 && ldx % 4 == 0, "small_linear_bwd: bad shape rows=%d N=%d K=%d", rows, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (dW || db) {
    const long long wwarps = (long long)N * ((K / 4 + 31) / 32);
    launch(small_linear_wgrad_kernel, dim3((unsigned)((wwarps + 7) / 8)), dim3(256), 0, st, dy, y, x, ldx, dW, db, rows, N, K, act);
    int rc = check_launch("small_linear_wgrad_kernel");
    if (rc) return rc;
  }
  if (dx) {
    VB_REQUIRE(lddx % 4 == 0, "small_linear_bwd: lddx must be a multiple of 4");
    const long long ctas = (long long)rows * ((K / 4 + 31) / 32);
    launch(small_linear_dgrad_kernel, dim3((unsigned)ctas), dim3(256), 0, st, dy, y, W, dx, lddx, accumulate_dx, rows, N, K, act);
    return check_launch("small_linear_dgrad_kernel");
  }
  return VAULT_OK;
}

extern "C" int vault_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                                 void* stream) {
  VB_REQUIRE(x && y && n >= 0 && p >= 0.f && p < 1.f, "dropout_f32: bad arguments");
  if (n == 0) return VAULT_OK;
  const long long q = (n + 3) / 4;
  launch(dropout_f32_kernel, dim3((unsigned)((q + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, y, n, p, seed, reinterpret_cast<const unsigned long long*>(seed_dev), site);
  return check_launch("dropout_f32_kernel");
}

extern "C" int vault_ce_loss(const float* logits, const int64_t* labels, float* loss, float* dlogits, int32_t rows, int32_t n_classes,
                             float grad_scale, void* stream) {
  VB_REQUIRE(logits && labels && loss && rows > 0 && n_classes > 0, "ce_loss: bad arguments");
  launch(ce_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, logits, labels, loss, dlogits, rows, n_classes, grad_scale);
  return check_launch("ce_loss_kernel");
}

extern "C" int vault_head_loss(const float* logits, const void* labels, float* loss, float* dlogits, int32_t rows, int32_t n_classes, int32_t kind,
                               float grad_scale, void* stream) {
  VB_REQUIRE(logits && labels && loss && rows > 0 && n_classes > 0, "head_loss: bad arguments");
  VB_REQUIRE(kind >= 0 && kind <= 2, "head_loss: kind=%d (0 CE, 1 BCE-with-logits, 2 two-group CE)", kind);
  if (kind == 0) return vault_ce_loss(logits, reinterpret_cast<const int64_t*>(labels), loss, dlogits, rows, n_classes, grad_scale, stream);
  VB_REQUIRE(kind != 1 || n_classes == 1, "head_loss: BCE-with-logits expects n_classes = 1 (got %d)", n_classes);
  VB_REQUIRE(kind != 2 || n_classes % 2 == 0, "head_loss: two-group CE expects an even n_classes (got %d)", n_classes);
  launch(head_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, logits, labels, loss, dlogits, rows, n_classes, kind, grad_scale);
  return check_launch("head_loss_kernel");
}

extern "C" int vault_colsum_bf16(const void* x, int64_t ldx, float* out, int64_t rows, int32_t cols, void* stream) {
  VB_REQUIRE(x && out && rows >= 0 && cols > 0 && cols % 4 == 0 && ldx % 4 == 0, "colsum: bad arguments");
  if (rows == 0) return VAULT_OK;
  long long gy = (rows + 63) / 64;
  if (gy > 64) gy = 64;
  dim3 grid((cols + 127) / 128, (unsigned)gy);
  launch(colsum_bf16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const bf16*>(x), ldx, out, rows, cols);
  return check_launch("colsum_bf16_kernel");
}
