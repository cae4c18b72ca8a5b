// conformance_b200.cpp -- TEST INFRASTRUCTURE: the reference's OWN test driver pointed at the GPU path.
//
// /root/reference/src/test.cpp:76-179 (test<>, testAllDistributions<>,
// testAllPayloads<>, testAllTypes<>, testAll<>) with the Data<> generators and checkData() of
// /root/reference/src/data.hpp is the reference's only test.  This file #includes that driver WHERE IT LIES
// (found through -I/root/reference/src; nothing of it is copied into this repository), turns its main()
// into a never-instantiated template, and runs testAll<> with SortMethodB200: an adapter with the facade of
// /root/reference/src/sort_methods.hpp:24-98 (name / isSupported / sort / sortThresh) whose sortThresh
// forwards to the C ABI of libb200sort.so (include/b200sort.h) -- host arrays in, host arrays out, the way
// the reference's callers use it.
//
// Built by `make -C oracle conformance` (needs /root/reference; the binary lands in oracle/_ref/ and travels
// to the GPU box with the snapshot).  Run: conformance_b200 [maxNum=10000] [seed=42]
#define main(...) b200_reference_main_is_unused(); template <int B200Unused> int b200_reference_main(__VA_ARGS__)
#include "test.cpp"  // the reference's src/test.cpp (by include path)
#undef main

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "b200sort.h"

namespace {

template <typename K>
constexpr int b200_key_code() {
  if constexpr (std::is_same_v<K, uint8_t>) return B200SORT_U8;
  else if constexpr (std::is_same_v<K, int8_t>) return B200SORT_I8;
  else if constexpr (std::is_same_v<K, uint16_t>) return B200SORT_U16;
  else if constexpr (std::is_same_v<K, int16_t>) return B200SORT_I16;
  else if constexpr (std::is_same_v<K, uint32_t>) return B200SORT_U32;
  else if constexpr (std::is_same_v<K, int32_t>) return B200SORT_I32;
  else if constexpr (std::is_same_v<K, uint64_t>) return B200SORT_U64;
  else if constexpr (std::is_same_v<K, int64_t>) return B200SORT_I64;
  else if constexpr (std::is_same_v<K, float>) return B200SORT_F32;
  else if constexpr (std::is_same_v<K, double>) return B200SORT_F64;
  else return -1;
}

void b200_check(int rc) {
  if (rc != 0) {
    std::printf("b200sort error %d: %s\n", rc, b200sort_last_error());
    std::exit(3);
  }
}

long g_sorts = 0;

}  // namespace

// facade of SortMethodRadixSort (src/sort_methods.hpp:24-98)
struct SortMethodB200 {
  static std::string name() { return "RadixB200"; }
  static constexpr bool areKeyAndPayloadSeparate = true;
  static constexpr bool hasThreshold = true;

  template <bool Up, typename K, typename... Ps>
  static constexpr bool isSupported() {
    return b200_key_code<K>() >= 0 && sizeof...(Ps) <= 63 && ((sizeof(Ps) <= 64) && ...);
  }

  template <bool Up = true, typename K, typename... Ps>
  static void sort(const simd_sort::SortIndex num, K *const keys, Ps *const... payloads) {
    sortThresh<Up>(16, num, keys, payloads...);
  }

  // separate key and payload arrays (src/test.cpp:53-59)
  template <bool Up = true, typename K, typename... Ps>
  static void sortThresh(const simd_sort::SortIndex cmpSortThresh, const simd_sort::SortIndex num, K *const keys,
                         Ps *const... payloads) {
    static_assert(isSupported<Up, K, Ps...>(), "Unsupported type combination");
    void *ptrs[sizeof...(Ps) + 1] = {static_cast<void *>(payloads)...};
    const uint32_t sizes[sizeof...(Ps) + 1] = {static_cast<uint32_t>(sizeof(Ps))...};
    g_sorts++;
    b200_check(b200sort_sort_soa_ex(keys, b200_key_code<K>(), num, Up ? 1 : 0, (int)sizeof...(Ps), ptrs, sizes, cmpSortThresh,
                                    B200SORT_CMP_INSERTION, nullptr, nullptr, 0));
  }

  // combined records (src/test.cpp:44-52)
  template <bool Up = true, typename K, typename... Ps>
  static void sortThresh(const simd_sort::SortIndex cmpSortThresh, const simd_sort::SortIndex num,
                         simd_sort::DataElement<K, Ps...> *const elements) {
    using R = simd_sort::DataElement<K, Ps...>;
    static_assert(simd_sort::is_power_of_two<sizeof(R)>, "size of DataElement<K, Ps...> must be a power of two");
    g_sorts++;
    b200_check(b200sort_sort_aos_ex(elements, b200_key_code<K>(), (uint32_t)sizeof(R), num, Up ? 1 : 0, cmpSortThresh,
                                    B200SORT_CMP_INSERTION, nullptr, nullptr, 0));
  }
};

int main(int argc, char const *argv[]) {
  const std::size_t maxNum = argc > 1 ? std::stoul(argv[1]) : 10000;
  const unsigned int seed = argc > 2 ? (unsigned)std::stoul(argv[2]) : 42u;
  bool passed = true;
  for (std::size_t num = 1; num <= maxNum; num *= 10) {  // src/test.cpp:185
    std::cout << "Testing " << num << " elements" << std::endl;
    passed &= testAll<SortMethodB200>(num, seed);
  }
  std::cout << "sorts through libb200sort: " << g_sorts << ", kernels launched: " << b200sort_launch_count() << std::endl;
  if (passed) {
    std::cout << "All tests passed" << std::endl;  // the reference driver's own pass line (src/test.cpp:217)
    return 0;
  }
  std::cout << "Tests failed, see above for details" << std::endl;
  return 1;
}
