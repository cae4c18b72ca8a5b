// oracle/ref_shim.cpp  --  TEST INFRASTRUCTURE ONLY.
//
// C-callable shim around the UNMODIFIED reference header
// /root/reference/radixSort.hpp (found through -I, never copied into this
// repository).  oracle/Makefile compiles this file once per key type
// (-DSHIM_KEY_CODE=n) plus once as the dispatcher (-DSHIM_FRONT) into
// oracle/_ref/libref_sort.so.  The library is what pins the C restatement in
// radix_oracle.c, what produced tests/golden/, and what bench.py times as the
// "reference" CPU baseline.  The product never links or loads it.
//
// Entry points instantiated (reference interface each one calls):
//   ref_sort_soa -> simd_sort::radix_sort::sort<Up>(num, K*, Ps*...)        radixSort.hpp:1780-1783
//   ref_sort_aos -> simd_sort::radix_sort::sort<Up>(num, DataElement<K,..>*) radixSort.hpp:1770-1778
// Payload streams are opaque to the reference (it only moves them), so each
// stream is instantiated as the unsigned integer of its byte width, and AoS
// records as DataElement<K, std::array<uint8_t, record_bytes - sizeof(K)>>.
#include <array>
#include <cstddef>
#include <cstdint>
#include <utility>

#if !defined(SHIM_FRONT)
#include "radixSort.hpp"
#endif

#define SHIM_CAT2(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT2(a, b)

#if defined(SHIM_FRONT)

#define DECL(n)                                                                                   \
  extern "C" int SHIM_CAT(ref_sort_soa_k, n)(void *, int64_t, int, int, void *const *,             \
                                             const uint32_t *);                                    \
  extern "C" int SHIM_CAT(ref_sort_aos_k, n)(void *, uint32_t, int64_t, int);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8) DECL(9)
#undef DECL

extern "C" int ref_sort_soa(void *keys, int key_type, int64_t num, int up, int n_payloads,
                            void *const *payloads, const uint32_t *payload_elem_bytes) {
  switch (key_type) {
#define CASE(n) case n: return SHIM_CAT(ref_sort_soa_k, n)(keys, num, up, n_payloads, payloads, payload_elem_bytes);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  return -1;
}

extern "C" int ref_sort_aos(void *records, int key_type, uint32_t record_bytes, int64_t num, int up) {
  switch (key_type) {
#define CASE(n) case n: return SHIM_CAT(ref_sort_aos_k, n)(records, record_bytes, num, up);
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
  }
  return -1;
}

// 1 when this CPU can execute the AVX-512 subset the reference is built with.
extern "C" int ref_cpu_ok(void) {
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
         __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512vbmi2");
}

#else  // one key type per translation unit

#if SHIM_KEY_CODE == 0
using Key = uint8_t;
#elif SHIM_KEY_CODE == 1
using Key = int8_t;
#elif SHIM_KEY_CODE == 2
using Key = uint16_t;
#elif SHIM_KEY_CODE == 3
using Key = int16_t;
#elif SHIM_KEY_CODE == 4
using Key = uint32_t;
#elif SHIM_KEY_CODE == 5
using Key = int32_t;
#elif SHIM_KEY_CODE == 6
using Key = uint64_t;
#elif SHIM_KEY_CODE == 7
using Key = int64_t;
#elif SHIM_KEY_CODE == 8
using Key = float;
#elif SHIM_KEY_CODE == 9
using Key = double;
#else
#error "SHIM_KEY_CODE must be 0..9"
#endif

namespace {

template <class... Ps, std::size_t... I>
void call_soa(bool up, int64_t num, Key *keys, void *const *p, std::index_sequence<I...>) {
  if (up)
    simd_sort::radix_sort::sort<true>(num, keys, static_cast<Ps *>(p[I])...);
  else
    simd_sort::radix_sort::sort<false>(num, keys, static_cast<Ps *>(p[I])...);
}

template <class... Ps>
int soa(bool up, int64_t num, void *keys, void *const *p) {
  call_soa<Ps...>(up, num, static_cast<Key *>(keys), p, std::index_sequence_for<Ps...>{});
  return 0;
}

template <std::size_t RB>
struct Rec {
  using type = simd_sort::DataElement<Key, std::array<uint8_t, RB - sizeof(Key)>>;
};
template <>
struct Rec<sizeof(Key)> {
  using type = simd_sort::DataElement<Key>;
};

template <std::size_t RB>
int aos(bool up, int64_t num, void *records) {
  if constexpr (RB < sizeof(Key)) {
    return -3;
  } else {
    using R = typename Rec<RB>::type;
    static_assert(sizeof(R) == RB && offsetof(R, key) == 0);
    if (up)
      simd_sort::radix_sort::sort<true>(num, static_cast<R *>(records));
    else
      simd_sort::radix_sort::sort<false>(num, static_cast<R *>(records));
    return 0;
  }
}

// payload-shape signature: one base-16 digit (log2(bytes)+1) per stream
constexpr uint32_t sig(std::initializer_list<uint32_t> bytes) {
  uint32_t s = 0;
  for (uint32_t b : bytes) s = s * 16 + (b == 1 ? 1 : b == 2 ? 2 : b == 4 ? 3 : b == 8 ? 4 : 15);
  return s;
}

}  // namespace

extern "C" int SHIM_CAT(ref_sort_soa_k, SHIM_KEY_CODE)(void *keys, int64_t num, int up, int n_payloads,
                                                       void *const *p, const uint32_t *bytes) {
  uint32_t s = 0;
  for (int i = 0; i < n_payloads; i++) {
    const uint32_t b = bytes[i];
    s = s * 16 + (b == 1 ? 1 : b == 2 ? 2 : b == 4 ? 3 : b == 8 ? 4 : 15);
  }
  using u8 = uint8_t; using u16 = uint16_t; using u32 = uint32_t; using u64 = uint64_t;
  switch (s) {
    case sig({}): return soa<>(up, num, keys, p);
    case sig({1}): return soa<u8>(up, num, keys, p);
    case sig({2}): return soa<u16>(up, num, keys, p);
    case sig({4}): return soa<u32>(up, num, keys, p);
    case sig({8}): return soa<u64>(up, num, keys, p);
    case sig({8, 1}): return soa<u64, u8>(up, num, keys, p);          // src/test.cpp:112-114
    case sig({8, 8}): return soa<u64, u64>(up, num, keys, p);         // src/test.cpp:115-117
    case sig({8, 8, 8}): return soa<u64, u64, u64>(up, num, keys, p); // src/test.cpp:118-119
    case sig({4, 8, 2}): return soa<u32, u64, u16>(up, num, keys, p); // BASELINE.json config 3
    case sig({1, 1, 1}): return soa<u8, u8, u8>(up, num, keys, p);    // src/test.cpp:149-150
    default: return -2;  // shape not instantiated
  }
}

extern "C" int SHIM_CAT(ref_sort_aos_k, SHIM_KEY_CODE)(void *records, uint32_t record_bytes, int64_t num,
                                                       int up) {
  switch (record_bytes) {
    case 1: return aos<1>(up, num, records);
    case 2: return aos<2>(up, num, records);
    case 4: return aos<4>(up, num, records);
    case 8: return aos<8>(up, num, records);
    case 16: return aos<16>(up, num, records);
    case 32: return aos<32>(up, num, records);
    case 64: return aos<64>(up, num, records);
    default: return -3;  // radixSort.hpp:1773-1774 static_asserts the power-of-two rule
  }
}

#endif
