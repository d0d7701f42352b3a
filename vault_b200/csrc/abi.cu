// C-ABI plumbing: version, per-thread last-error text, device check.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace vb {
static thread_local char g_err[512] = {0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(VAULT_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return VAULT_OK;
}

int device_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}
}  // namespace vb

extern "C" {
int vault_version(void) { return 100; }

size_t vault_last_error(char* buf, size_t cap) {
  size_t n = strlen(vb::g_err);
  if (buf && cap) {
    size_t c = n < cap - 1 ? n : cap - 1;
    memcpy(buf, vb::g_err, c);
    buf[c] = 0;
  }
  return n;
}

int vault_check_device(int dev) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || dev >= n) {
    cudaGetLastError();
    return vb::fail(VAULT_ERR_ARCH, "no CUDA device %d (vault_b200 has no CPU fallback)", dev);
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return vb::fail(VAULT_ERR_ARCH, "device %d is sm_%d?, need sm_100 (B200)", dev, major * 10);
  return VAULT_OK;
}
}
