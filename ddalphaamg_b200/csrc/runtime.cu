// runtime.cu -- device memory + stream plumbing (one process per GPU).
#include "common.cuh"
#include <map>

namespace dda {

cudaStream_t g_stream = 0;
long g_launch_count = 0;
static size_t g_bytes = 0;
static std::map<void *, size_t> g_allocs;

#ifndef DDA_HOST_EMU
void *dev_alloc_bytes(size_t bytes) {
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    fprintf(stderr, "dd_alpha_amg_b200: cudaMalloc of %zu bytes failed (%s); %zu bytes already in use\n", bytes, cudaGetErrorString(e), g_bytes);
    fatal("out of device memory", __FILE__, __LINE__);
  }
  g_bytes += bytes; g_allocs[p] = bytes;
  return p;
}
void dev_free(void *p) {
  if (!p) return;
  auto it = g_allocs.find(p);
  if (it != g_allocs.end()) { g_bytes -= it->second; g_allocs.erase(it); }
  CUDA_CHECK(cudaFree(p));
}
void dev_zero(void *p, size_t bytes) { CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, g_stream)); }
void h2d(void *dst, const void *src, size_t bytes) { CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream)); CUDA_CHECK(cudaStreamSynchronize(g_stream)); }
void d2h(void *dst, const void *src, size_t bytes) { CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream)); CUDA_CHECK(cudaStreamSynchronize(g_stream)); }
void d2d(void *dst, const void *src, size_t bytes) { CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream)); }
void dev_sync() { CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaGetLastError()); }
#else
void *dev_alloc_bytes(size_t bytes) {
  if (bytes == 0) bytes = 16;
  void *p = calloc(1, bytes);
  if (!p) fatal("out of host memory (emulation build)", __FILE__, __LINE__);
  g_bytes += bytes; g_allocs[p] = bytes;
  return p;
}
void dev_free(void *p) {
  if (!p) return;
  auto it = g_allocs.find(p);
  if (it != g_allocs.end()) { g_bytes -= it->second; g_allocs.erase(it); }
  free(p);
}
void dev_zero(void *p, size_t bytes) { memset(p, 0, bytes); }
void h2d(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void d2h(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
void d2d(void *dst, const void *src, size_t bytes) { memmove(dst, src, bytes); }
void dev_sync() {}
#endif
size_t dev_bytes_in_use() { return g_bytes; }
int dev_sm_count() {
#ifndef DDA_HOST_EMU
  static int sms = 0;
  if (!sms) { int dev = 0; CUDA_CHECK(cudaGetDevice(&dev)); CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); }
  return sms;
#else
  return 1;
#endif
}

}  // namespace dda
