// cudaMalloc cost: many 1 GB allocations vs one slab of the same total (setup allocates ~60 GB in ~100 calls)
#include <cstdio>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaFree(0);
  const size_t GB = 1ull << 30;
  for (int rep = 0; rep < 2; rep++) {
    std::vector<void *> p(48);
    double t0 = now();
    for (auto &q : p) if (cudaMalloc(&q, GB) != cudaSuccess) { printf("fail\n"); return 1; }
    double t1 = now();
    for (auto q : p) cudaFree(q);
    double t2 = now();
    void *slab;
    if (cudaMalloc(&slab, 48 * GB) != cudaSuccess) { printf("fail slab\n"); return 1; }
    double t3 = now();
    cudaFree(slab);
    double t4 = now();
    printf("48 x 1 GB: malloc %.3f s, free %.3f s;  one 48 GB slab: malloc %.3f s, free %.3f s\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3);
  }
  return 0;
}
