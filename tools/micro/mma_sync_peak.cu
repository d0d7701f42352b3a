// Peak rate of legacy mma.sync.m16n8k16 bf16 on this GPU (register-resident operands, 8 independent accumulator chains per warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  float c[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int iters = 20000; dim3 grid(148 * (warps > 16 ? 2 : 1)), block(32 * (warps > 16 ? warps / 2 : warps));
    k<<<grid, block>>>(out, 100); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * (double)grid.x * (block.x / 32);
    printf("warps/SM %d: %.1f TFLOP/s (%s)\n", warps, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
