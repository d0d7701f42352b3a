// Image pre-processing of the ViLT processor on the GPU: per-image bicubic resize (uint8, the fixed-point two-pass algorithm of the
// reference's resize backend), rescale 1/255 + normalise (mean = std = 0.5) through a 3 x 256-entry table, zero padding to the batch
// maximum and the pixel mask -- HF:models/vilt/image_processing_vilt.py (resize / rescale / normalize / pad + make_pixel_mask) with
// Pillow's ImagingResample (src/libImaging/Resample.c, 8 bits per channel) underneath, as called by transformers==4.48.0
// `image_transforms.resize(..., resample=BICUBIC, reducing_gap=None)`.
//
// The filter taps are computed on the host in double precision exactly as Pillow does (precompute_coeffs + normalize_coeffs_8bpc,
// 22 fractional bits) and uploaded as int32 tables, so the device work is pure integer arithmetic and the uint8 result is
// bit-identical to Pillow's:   out = clip8((2^21 + sum_k pixel[xmin + k] * coef[k]) >> 22), horizontal pass first (rounded to
// uint8), then vertical.  HBM-bound byte work: one coalesced read of the source, one write of the fp32 NCHW batch.
#include "common.cuh"

namespace vb {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// pass 1: tmp[b][y][x_out][c] = horizontal resample of src[b][y][:, c]       grid (x blocks, y, image)
__global__ void __launch_bounds__(128)
image_resize_h_kernel(const uint8_t* __restrict__ src, const vault_image_desc* __restrict__ descs, const int32_t* __restrict__ coefs,
                      const int32_t* __restrict__ bounds, uint8_t* __restrict__ tmp) {
  pdl_enter();
  const vault_image_desc d = descs[blockIdx.z];
  const int y = blockIdx.y, x = blockIdx.x * 128 + threadIdx.x;
  if (y >= d.h_in || x >= d.w_out) return;
  const int xmin = bounds[d.bound_h + 2 * x], n = bounds[d.bound_h + 2 * x + 1];
  const int32_t* k = coefs + d.coef_h + (long long)x * d.ksize_h;
  const uint8_t* row = src + d.src_off + ((long long)y * d.w_in + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < n; ++i) {
    const int w = k[i];
    s0 += row[3 * i] * w;
    s1 += row[3 * i + 1] * w;
    s2 += row[3 * i + 2] * w;
  }
  uint8_t* o = tmp + d.tmp_off + ((long long)y * d.w_out + x) * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// pass 2: vertical resample of tmp, table lookup (rescale + normalise), NCHW fp32 write incl. the zero padding, pixel mask
__global__ void __launch_bounds__(128)
image_resize_v_kernel(const vault_image_desc* __restrict__ descs, const int32_t* __restrict__ coefs, const int32_t* __restrict__ bounds,
                      const uint8_t* __restrict__ tmp, const float* __restrict__ lut, float* __restrict__ pixel_values, int64_t* __restrict__ pixel_mask,
                      int Hmax, int Wmax) {
  pdl_enter();
  const int b = blockIdx.z;
  const vault_image_desc d = descs[b];
  const int y = blockIdx.y, x = blockIdx.x * 128 + threadIdx.x;
  if (x >= Wmax) return;
  const long long plane = (long long)Hmax * Wmax;
  float* o = pixel_values + (long long)b * 3 * plane + (long long)y * Wmax + x;
  const bool inside = y < d.h_out && x < d.w_out;
  pixel_mask[(long long)b * plane + (long long)y * Wmax + x] = inside ? 1 : 0;
  if (!inside) {
    o[0] = 0.f; o[plane] = 0.f; o[2 * plane] = 0.f;
    return;
  }
  const int ymin = bounds[d.bound_v + 2 * y], n = bounds[d.bound_v + 2 * y + 1];
  const int32_t* k = coefs + d.coef_v + (long long)y * d.ksize_v;
  const uint8_t* col = tmp + d.tmp_off + ((long long)ymin * d.w_out + x) * 3;
  const long long stride = (long long)d.w_out * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < n; ++i) {
    const int w = k[i];
    s0 += col[i * stride] * w;
    s1 += col[i * stride + 1] * w;
    s2 += col[i * stride + 2] * w;
  }
  o[0] = lut[clip8(s0)];
  o[plane] = lut[256 + clip8(s1)];
  o[2 * plane] = lut[512 + clip8(s2)];
}

}  // namespace
}  // namespace vb

extern "C" int vault_image_preprocess(const uint8_t* src, const vault_image_desc* descs, const int32_t* coefs, const int32_t* bounds, uint8_t* tmp,
                                      const float* lut256, float* pixel_values, int64_t* pixel_mask, int32_t B, int32_t Hmax, int32_t Wmax,
                                      int32_t max_h_in, int32_t max_w_out, void* stream) {
  using namespace vb;
  VB_REQUIRE(src && descs && coefs && bounds && tmp && lut256 && pixel_values && pixel_mask, "image_preprocess: null pointer");
  VB_REQUIRE(B > 0 && Hmax > 0 && Wmax > 0 && max_h_in > 0 && max_w_out > 0 && max_h_in <= 65535 && Hmax <= 65535 && B <= 65535,
             "image_preprocess: bad shape B=%d Hmax=%d Wmax=%d max_h_in=%d", B, Hmax, Wmax, max_h_in);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  launch(image_resize_h_kernel, dim3((max_w_out + 127) / 128, max_h_in, B), dim3(128), 0, st, src, descs, coefs, bounds, tmp);
  int rc = check_launch("image_resize_h_kernel");
  if (rc) return rc;
  launch(image_resize_v_kernel, dim3((Wmax + 127) / 128, Hmax, B), dim3(128), 0, st, descs, coefs, bounds, (const uint8_t*)tmp, lut256, pixel_values,
         pixel_mask, Hmax, Wmax);
  return check_launch("image_resize_v_kernel");
}
