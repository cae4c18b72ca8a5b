// cub_compare.cu -- external yardstick ONLY: CUB DeviceRadixSort::SortPairs (the library shipped with the
// CUDA toolkit) on the headline shape, timed the way sortbench times libb200sort.  Never linked into the
// product; its numbers go into profiles/ as a comparison row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cub_compare tools/cub_compare.cu
//   ./cub_compare [n] [iters]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__global__ void fill(uint64_t *k, uint64_t *p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { k[i] = mix64(12345 + i); p[i] = (uint64_t)i; }
}
template <typename K, typename V>
static void run(int64_t n, int iters, const char *name) {
  K *k0, *k1; V *v0, *v1;
  CK(cudaMalloc(&k0, n * sizeof(K))); CK(cudaMalloc(&k1, n * sizeof(K)));
  CK(cudaMalloc(&v0, n * sizeof(V))); CK(cudaMalloc(&v1, n * sizeof(V)));
  size_t tmp_bytes = 0;
  cub::DoubleBuffer<K> dk(k0, k1); cub::DoubleBuffer<V> dv(v0, v1);
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n));
  void *tmp; CK(cudaMalloc(&tmp, tmp_bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double tot = 0, best = 1e30;
  for (int it = 0; it <= iters; it++) {
    fill<<<1184, 256>>>((uint64_t *)k0, (uint64_t *)v0, n * sizeof(K) / 8 < n ? n * sizeof(K) / 8 : n);
    cub::DoubleBuffer<K> a(k0, k1); cub::DoubleBuffer<V> b(v0, v1);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, a, b, n));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it) { tot += ms; best = ms < best ? ms : best; }
  }
  const double avg = tot / iters;
  printf("CUB DeviceRadixSort::SortPairs %s n=%lld: best %.3f ms avg %.3f ms = %.3f Gpairs/s (floor 2NR: %.1f GB/s)\n", name, (long long)n, best, avg,
         n / avg * 1e-6, 2.0 * n * (sizeof(K) + sizeof(V)) / avg * 1e-6);
  cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
}
int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000000ll;
  const int iters = argc > 2 ? atoi(argv[2]) : 3;
  run<uint64_t, uint64_t>(n, iters, "u64+u64");
  run<uint32_t, uint32_t>(n, iters, "u32+u32");
  return 0;
}
