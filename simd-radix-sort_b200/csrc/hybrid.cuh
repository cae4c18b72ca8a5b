// hybrid.cuh -- MSB hybrid path for 8-byte keys (placeholder until the kernels land).
#pragma once
#include "kernels.cuh"
#include "../../include/b200sort.h"

namespace b200sort {

constexpr int HYB_MIN_TILE = 4096;
constexpr int64_t HYB_MIN_N = INT64_MAX;  // auto-selection disabled

struct HybridCtrl { uint32_t flags[64]; };

struct HybridJob {
  StreamSet ss; int64_t n; KeyOrder ko; unsigned char *ws; uint64_t *ghist; uint32_t *tile_counter; Plan *plan;
  uint64_t *bin_base; uint64_t *lookback; HybridCtrl *ctrl; int sm_count; size_t smem_optin; uint32_t stage_bytes;
  cudaStream_t stream; int cfg; bool use_match;
};

static int fail(int code, const char *fmt, ...);

inline int hybrid_sort_u64(const HybridJob &, b200sort_stats *) { return B200SORT_EUNSUPPORTED; }

}  // namespace b200sort
