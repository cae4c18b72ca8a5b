// sortbench.cu -- native micro-benchmark / self-check driver for libb200sort.so (a development tool,
// not part of the product path).  Generates inputs on the device, sorts through the C ABI, verifies
// order + permutation checksums on the device, and prints CUDA-event timings.
//
//   sortbench --n 100000000 --key u64 --pay 8 --iters 5 --opt algo=1 --opt tile_cfg=0
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200sort.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// dist: 0 uniform bits, 1 few-unique (16 values -8..7 as signed), 2 zipf-ish over 2^20 ranks, 3 all zero,
// 4 uniform float in (-1,1) (for f32/f64 keys), 5 sorted, 6 uniform 62-bit values (at 2^28 records the
// same run density below a 32-bit cut as 1e9 uniform 64-bit keys)
__global__ void fill_keys(unsigned char *keys, int kb, uint32_t stride, int64_t n, uint64_t seed, int dist, int is_float) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t r = mix64(seed + (uint64_t)i);
    uint64_t v = r;
    if (dist == 1) v = (uint64_t)((int64_t)(r % 16) - 8);
    else if (dist == 2) {
      // rank ~ 2^(20*u^2): heavy head, long tail; mapped to a full-range value by hashing the rank
      double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
      uint64_t rank = (uint64_t)exp2(20.0 * u * u);
      v = mix64(rank * 0x1234567ull + 99);
    } else if (dist == 3) v = 0;
    else if (dist == 5) v = (uint64_t)i * 3;
    else if (dist == 6) v = r >> 2;
    if (dist == 4 || is_float) {
      double u = (double)(r >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
      if (kb == 4) { float f = (float)u; memcpy(&v, &f, 4); }
      else { memcpy(&v, &u, 8); }
    }
    unsigned char *p = keys + (size_t)i * stride;
    if (kb == 8) *(uint64_t *)p = v;
    else if (kb == 4) *(uint32_t *)p = (uint32_t)v;
    else if (kb == 2) *(uint16_t *)p = (uint16_t)v;
    else *p = (uint8_t)v;
  }
}

// payload element = function of (key bits, stream id): lets the check verify "payload moved with its key"
__global__ void fill_payload(unsigned char *pay, int eb, const unsigned char *keys, int kb, uint32_t kstride, int64_t n, int sid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = 0;
    memcpy(&k, keys + (size_t)i * kstride, kb);
    for (int b = 0; b < eb; b += 8) {
      uint64_t h = mix64(k * 31 + sid * 1000003ull + b);
      memcpy(pay + (size_t)i * eb + b, &h, eb - b < 8 ? eb - b : 8);
    }
  }
}

__device__ __forceinline__ uint64_t ordered(uint64_t raw, int kb, int kind, int asc) {
  const uint64_t mask = kb == 8 ? ~0ull : ((1ull << (8 * kb)) - 1), sign = 1ull << (8 * kb - 1);
  uint64_t u = raw & mask;
  if (kind == 1) u ^= sign;
  else if (kind == 2) u = (u & sign) ? (~u & mask) : (u ^ sign);
  if (!asc) u = ~u & mask;
  return u;
}

// errors[0] = order violations, errors[1] = payload mismatches; sums[0] = xor-free checksum of keys
__global__ void check(const unsigned char *keys, int kb, uint32_t kstride, int64_t n, int kind, int asc,
                      const unsigned char *pay0, int eb0, unsigned long long *errors, unsigned long long *sum) {
  unsigned long long local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = 0, kp = 0;
    memcpy(&k, keys + (size_t)i * kstride, kb);
    local += mix64(k);
    if (i > 0) {
      memcpy(&kp, keys + (size_t)(i - 1) * kstride, kb);
      if (ordered(kp, kb, kind, asc) > ordered(k, kb, kind, asc)) atomicAdd(&errors[0], 1ull);
    }
    if (pay0) {
      for (int b = 0; b < eb0; b += 8) {
        uint64_t h = mix64(k * 31 + 0 * 1000003ull + b), got = 0;
        const int w = eb0 - b < 8 ? eb0 - b : 8;
        memcpy(&got, pay0 + (size_t)i * eb0 + b, w);
        if (w < 8) h &= (1ull << (8 * w)) - 1;
        if (got != h) atomicAdd(&errors[1], 1ull);
      }
    }
  }
  atomicAdd(sum, local);
}

__global__ void flush_l2(uint4 *buf, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_uint4(1, 2, 3, 4);
}

int main(int argc, char **argv) {
  int64_t n = 1 << 24;
  std::string key = "u64";
  std::vector<int> pay;
  int iters = 3, dist = 0, asc = 1, aos = 0, verify = 1, prof = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (a == "--n") n = atoll(next().c_str());
    else if (a == "--key") key = next();
    else if (a == "--pay") { std::string s = next(); size_t p = 0; while (p < s.size()) { size_t q = s.find(',', p); if (q == std::string::npos) q = s.size(); if (q > p) pay.push_back(atoi(s.substr(p, q - p).c_str())); p = q + 1; } }
    else if (a == "--iters") iters = atoi(next().c_str());
    else if (a == "--dist") dist = atoi(next().c_str());
    else if (a == "--desc") asc = 0;
    else if (a == "--aos") aos = atoi(next().c_str());   // record bytes
    else if (a == "--noverify") verify = 0;
    else if (a == "--prof") { prof = 1; b200sort_set_option("profile", 1); }
    else if (a == "--opt") { std::string s = next(); size_t e = s.find('='); if (b200sort_set_option(s.substr(0, e).c_str(), atoll(s.substr(e + 1).c_str()))) { fprintf(stderr, "%s\n", b200sort_last_error()); return 2; } }
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  int kt, kb, kind;
  if (key == "u8") kt = 0, kb = 1, kind = 0; else if (key == "i8") kt = 1, kb = 1, kind = 1;
  else if (key == "u16") kt = 2, kb = 2, kind = 0; else if (key == "i16") kt = 3, kb = 2, kind = 1;
  else if (key == "u32") kt = 4, kb = 4, kind = 0; else if (key == "i32") kt = 5, kb = 4, kind = 1;
  else if (key == "u64") kt = 6, kb = 8, kind = 0; else if (key == "i64") kt = 7, kb = 8, kind = 1;
  else if (key == "f32") kt = 8, kb = 4, kind = 2; else if (key == "f64") kt = 9, kb = 8, kind = 2;
  else { fprintf(stderr, "bad key type\n"); return 2; }

  const uint32_t kstride = aos ? aos : kb;
  unsigned char *keys, *keys0;
  CK(cudaMalloc(&keys, (size_t)n * kstride));
  CK(cudaMalloc(&keys0, (size_t)n * kstride));
  std::vector<unsigned char *> pl(pay.size()), pl0(pay.size());
  std::vector<uint32_t> pb(pay.begin(), pay.end());
  size_t rec = kstride;
  for (size_t p = 0; p < pay.size(); p++) { CK(cudaMalloc(&pl[p], (size_t)n * pay[p])); CK(cudaMalloc(&pl0[p], (size_t)n * pay[p])); rec += pay[p]; }
  if (aos) CK(cudaMemset(keys0, 0x5a, (size_t)n * kstride));
  fill_keys<<<1184, 256>>>(keys0, kb, kstride, n, 12345, dist, kind == 2);
  for (size_t p = 0; p < pay.size(); p++) fill_payload<<<1184, 256>>>(pl0[p], pay[p], keys0, kb, kstride, n, (int)p);
  CK(cudaDeviceSynchronize());

  unsigned long long *d_err, *d_sum;
  CK(cudaMalloc(&d_err, 16)); CK(cudaMalloc(&d_sum, 8));
  unsigned long long sum0 = 0;
  if (verify) {
    CK(cudaMemset(d_err, 0, 16)); CK(cudaMemset(d_sum, 0, 8));
    check<<<1184, 256>>>(keys0, kb, kstride, n, kind, asc, nullptr, 0, d_err, d_sum);
    CK(cudaMemcpy(&sum0, d_sum, 8, cudaMemcpyDeviceToHost));
  }
  uint4 *flushbuf; const size_t flush_n = (size_t)256 << 20 >> 4;
  CK(cudaMalloc(&flushbuf, flush_n * 16));

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 1e30, total = 0;
  int rc_all = 0;
  for (int it = 0; it < iters + 1; it++) {  // iteration 0 is the warm-up
    CK(cudaMemcpy(keys, keys0, (size_t)n * kstride, cudaMemcpyDeviceToDevice));
    for (size_t p = 0; p < pay.size(); p++) CK(cudaMemcpy(pl[p], pl0[p], (size_t)n * pay[p], cudaMemcpyDeviceToDevice));
    flush_l2<<<1184, 256>>>(flushbuf, flush_n);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    int rc = aos ? b200sort_sort_aos(keys, kt, aos, n, asc, nullptr, nullptr, 0)
                 : b200sort_sort_soa(keys, kt, n, asc, (int)pay.size(), (void *const *)pl.data(), pb.data(), nullptr, nullptr, 0);
    CK(cudaEventRecord(e1));
    if (rc) { fprintf(stderr, "sort failed: %d %s\n", rc, b200sort_last_error()); return 1; }
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0) { best = ms < best ? ms : best; total += ms; }
    if (verify && (it == 0 || it == iters)) {
      CK(cudaMemset(d_err, 0, 16)); CK(cudaMemset(d_sum, 0, 8));
      check<<<1184, 256>>>(keys, kb, kstride, n, kind, asc, (pay.empty() || aos) ? nullptr : pl[0], pay.empty() ? 0 : pay[0], d_err, d_sum);
      unsigned long long err[2], sum1;
      CK(cudaMemcpy(err, d_err, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&sum1, d_sum, 8, cudaMemcpyDeviceToHost));
      if (err[0] || err[1] || sum1 != sum0) { printf("VERIFY FAILED: order_violations=%llu payload_mismatch=%llu checksum %s\n", err[0], err[1], sum1 == sum0 ? "ok" : "DIFFERS"); rc_all = 1; }
    }
  }
  if (prof) {
    int kinds[64]; float ms[64];
    const int np = b200sort_last_profile(kinds, ms, 64);
    static const char *names[] = {"hist", "scan", "sweep", "copyback", "segfix", "other"};
    printf("  per-kernel ms:");
    for (int i = 0; i < np; i++) printf(" %s=%.3f", names[kinds[i] < 6 ? kinds[i] : 5], ms[i]);
    printf("\n");
  }
  b200sort_stats st{};
  b200sort_last_stats(&st);
  const double avg = total / iters;
  if (st.segfix_passes) printf("  segfix moved %llu records (%.2f %%)\n", (unsigned long long)st.segfix_moved, 100.0 * st.segfix_moved / (double)n);
  printf("key=%s pay=%zu rec=%zuB n=%lld dist=%d algo=%u passes=%u hist=%u launches=%u | best %.3f ms avg %.3f ms | %.3f Gelem/s | alg %.1f GB -> %.1f GB/s | floor(2NR) %.1f GB/s %s\n",
         key.c_str(), pay.size(), rec, (long long)n, dist, st.algo, st.passes_planned, st.hist_sweeps, st.kernel_launches, best, avg,
         n / avg * 1e-6, st.algorithmic_bytes * 1e-9, st.algorithmic_bytes / avg * 1e-6, 2.0 * n * rec / avg * 1e-6, rc_all ? "FAILED" : "ok");
  return rc_all;
}
