// Uses the drop-in header exactly the way the reference's README.md:23-56 shows its own header being
// used (separate streams, combined DataElement array, ascending and descending), on host arrays.
// Exit code 0 and "DROPIN OK" on success; 3 when the sort call throws (e.g. no GPU: no CPU fallback).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "b200sort/radixSort.hpp"

using namespace simd_sort;

template <bool Up, typename K>
static bool ordered(const std::vector<K> &k) {
  for (size_t i = 1; i < k.size(); i++)
    if (Up ? k[i - 1] > k[i] : k[i - 1] < k[i]) return false;
  return true;
}

int main() {
  try {
    std::mt19937_64 gen(12345);
    const SortIndex num = 100000;
    // separate key and payload streams (README.md:23-31)
    std::vector<uint64_t> keys(num), pay64(num);
    std::vector<uint16_t> pay16(num);
    for (SortIndex i = 0; i < num; i++) { keys[i] = gen(); pay64[i] = keys[i] * 3 + 1; pay16[i] = (uint16_t)(keys[i] >> 7); }
    std::vector<uint64_t> sorted(keys);
    std::sort(sorted.begin(), sorted.end());
    radix_sort::sort(num, keys.data(), pay64.data(), pay16.data());
    if (keys != sorted) { puts("FAIL: u64 ascending keys"); return 1; }
    for (SortIndex i = 0; i < num; i++)
      if (pay64[i] != keys[i] * 3 + 1 || pay16[i] != (uint16_t)(keys[i] >> 7)) { puts("FAIL: payloads did not follow keys"); return 1; }

    // descending, float keys (README.md:45-49)
    std::vector<float> fk(num);
    std::vector<int32_t> fp(num);
    std::uniform_real_distribution<float> dist(-1.f, 1.f);
    for (SortIndex i = 0; i < num; i++) { fk[i] = dist(gen); fp[i] = (int32_t)(fk[i] * 1e6f); }
    radix_sort::sort<false>(num, fk.data(), fp.data());
    if (!ordered<false>(fk)) { puts("FAIL: float descending"); return 1; }
    for (SortIndex i = 0; i < num; i++) if (fp[i] != (int32_t)(fk[i] * 1e6f)) { puts("FAIL: float payload"); return 1; }

    // combined array (README.md:33-43): DataElement<int64_t, double> is 16 bytes
    using Rec = DataElement<int64_t, double>;
    static_assert(sizeof(Rec) == 16);
    std::vector<Rec> recs(num);
    for (SortIndex i = 0; i < num; i++) { recs[i].key = (int64_t)gen() >> 20; std::get<0>(recs[i].payloads) = (double)recs[i].key * 0.5; }
    radix_sort::sort(num, recs.data());
    for (SortIndex i = 0; i < num; i++) {
      if (i && recs[i - 1].key > recs[i].key) { puts("FAIL: AoS order"); return 1; }
      if (std::get<0>(recs[i].payloads) != (double)recs[i].key * 0.5) { puts("FAIL: AoS payload"); return 1; }
    }
    radix_sort::sort<false>(num, recs.data());
    for (SortIndex i = 1; i < num; i++) if (recs[i - 1].key < recs[i].key) { puts("FAIL: AoS descending"); return 1; }

    // the advanced overload (src/radix_sort.hpp:297-312) and the trivial sizes (src/radix_sort.hpp:276)
    std::vector<int16_t> small = {5, -3, 9, 0, -3};
    radix_sort::sort<true, radix_sort::BitSorterSIMD, CmpSorterInsertionSort>(16, (SortIndex)small.size(), small.data());
    if (!ordered<true>(small)) { puts("FAIL: thresh overload"); return 1; }
    radix_sort::sort(0, small.data());
    radix_sort::sort(1, small.data());
    puts("DROPIN OK");
    return 0;
  } catch (const std::exception &e) {
    printf("sort threw: %s\n", e.what());
    return 3;
  }
}
