// Stand-alone experiment for round 2 (NOT part of libvault_b200.so, not on any product path): two-shot all-reduce of a bf16 buffer over
// NVSwitch multicast.  Every rank owns 1/world of the range: it pulls the SUM of that slice from all ranks with multimem.ld_reduce (the
// switch adds, fp32 accumulation) and pushes the result back to every rank with multimem.st.  The buffer is symmetric memory
// (torch.distributed._symmetric_memory: empty + rendezvous -> multicast_ptr); the cross-rank barriers before and after the kernel are the
// handle's stream-ordered barrier(), issued by the Python driver (tools/mc_allreduce_probe.py).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/micro/libmm_allreduce.so tools/micro/multimem_allreduce.cu
#include <cuda_runtime.h>
#include <cstdint>

namespace {

__device__ __forceinline__ uint4 mm_ld_reduce_bf16x8(const char* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}

__device__ __forceinline__ void mm_st_b128(char* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int kUnroll = 4;  // independent 16-byte reductions in flight per thread (NVLink round trips are microseconds)

__global__ void __launch_bounds__(512) mm_allreduce_bf16_kernel(char* mc_slice, long long nvec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride * kUnroll) {
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long j = i + u * stride;
      if (j < nvec) v[u] = mm_ld_reduce_bf16x8(mc_slice + 16 * j);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long j = i + u * stride;
      if (j < nvec) mm_st_b128(mc_slice + 16 * j, v[u]);
    }
  }
}

}  // namespace

// mc_ptr: multicast address of the symmetric buffer; [byte_off, byte_off + nbytes) is the range to reduce (16-byte aligned, nbytes % 16 == 0).
// Rank r reduces and re-broadcasts the r-th of `world` equal slices (in 16-byte vectors; the last slice takes the remainder).
extern "C" int mm_allreduce_bf16(uint64_t mc_ptr, long long byte_off, long long nbytes, int rank, int world, int ctas, void* stream) {
  if (mc_ptr == 0 || (byte_off & 15) || (nbytes & 15) || world <= 0 || rank < 0 || rank >= world || ctas <= 0) return -1;
  const long long nvec = nbytes / 16;
  const long long per = (nvec + world - 1) / world;
  const long long lo = per * rank, hi = lo + per < nvec ? lo + per : nvec;
  if (hi <= lo) return 0;
  char* slice = reinterpret_cast<char*>(mc_ptr) + byte_off + 16 * lo;
  mm_allreduce_bf16_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(slice, hi - lo);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
