// mgpu.cuh -- multi-GPU sort: MSB range partition over the GPUs of one box.
// Included at the end of b200sort.cu (it uses that file's internals).
//
//   1. every rank histograms the top 16 bits of a sample of its order-mapped keys (top_hist_kernel)
//   2. ncclAllReduce(sum) of the histograms -> identical splitters on every rank (b200sort_mgpu_splitters)
//   3. exact number of records per destination (dest_count_kernel); ncclAllGather of these rows together
//      with every rank's workspace IPC handle -> receive offsets, peer mappings
//   4. partition + exchange in ONE kernel: a scatter pass of onesweep_kernel in LUT mode whose bucket d is
//      the receive region of GPU d -- the records leave the staging tile as coalesced stores straight
//      into peer memory over NVLink (cudaIpc mapping of the peer's workspace); a 1-element all-reduce is
//      the barrier after it.  Fallback when the workspaces cannot be mapped (or option mgpu_p2p = 0): the
//      same pass into local memory, then grouped ncclSend/ncclRecv per peer.
//   5. local sort of the received range, which starts in the workspace's shadow arrays (sort_device)
// The reference has nothing here (single thread); NCCL is loaded lazily with dlopen so that the
// single-GPU library has no NCCL dependency.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace b200sort {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

static NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define SYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));     \
    if (!api.field) { api.why = std::string("missing symbol ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    api.ok = true;
  });
  return api;
}

#define NCCL_TRY(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess)                                                                               \
      return fail(B200SORT_ENCCL, "%s failed: %s", #expr, nccl_api().GetErrorString(_r));                 \
  } while (0)

constexpr int MGPU_MAX_BITS = 16;

// counts of the top `bits` bits of the ordered key; warp-aggregated global atomics
struct TopHistArgs {
  const unsigned char *keys;
  uint32_t stride;
  int64_t n;
  KeyOrder ko;
  int shift;            // bin = range_bin(ordered key, lo, shift, nb)
  unsigned long long *hist;  // [nb], zeroed
  int64_t sample;       // every sample-th row of 32 keys is counted
  unsigned long long lo;
  uint32_t nb;
  unsigned long long *range;  // range_kernel: [0] = min ordered key, [1] = min of the complements (= ~max)
};

template <int KB>
__global__ void __launch_bounds__(256) top_hist_kernel(TopHistArgs a) {
  // every sample-th row of 32 consecutive keys (splitters only need proportions; the exact numbers of
  // records per destination are counted afterwards by dest_count_kernel)
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = (a.n + 31) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x / 32);
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); w * a.sample < n_rows; w += n_warps) {
    const int64_t i = w * a.sample * 32 + lane;
    const bool valid = i < a.n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t bin = range_bin((unsigned long long)to_ordered<KB, false>(load_key<KB>(a.keys, i, a.stride), a.ko), a.lo, a.shift, a.nb);
      const unsigned peers = __match_any_sync(vmask, bin);
      if ((peers & lanemask_lt()) == 0) atomicAdd(&a.hist[bin], (unsigned long long)__popc(peers));
    }
  }
}

// smallest and largest ordered key of the same sample: the histogram bins cover that range, not the whole key
// space (keys of a narrow range -- small integers, N(0,1) doubles -- would otherwise share one or two bins)
template <int KB>
__global__ void __launch_bounds__(256) range_kernel(TopHistArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = (a.n + 31) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x / 32);
  unsigned long long mn = ~0ull, nmx = ~0ull;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); w * a.sample < n_rows; w += n_warps) {
    const int64_t i = w * a.sample * 32 + lane;
    if (i < a.n) {
      const unsigned long long u = (unsigned long long)to_ordered<KB, false>(load_key<KB>(a.keys, i, a.stride), a.ko);
      mn = u < mn ? u : mn;
      nmx = ~u < nmx ? ~u : nmx;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mn, s), n2 = __shfl_xor_sync(0xffffffffu, nmx, s);
    mn = m2 < mn ? m2 : mn;
    nmx = n2 < nmx ? n2 : nmx;
  }
  if (lane == 0) {
    atomicMin(&a.range[0], mn);
    atomicMin(&a.range[1], nmx);
  }
}

static cudaError_t launch_top_hist(int kb, const TopHistArgs &a, int sm_count, cudaStream_t st, bool range_only = false) {
  const int64_t rows = ((a.n + 31) / 32 + a.sample - 1) / a.sample;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count * 16));
  ProfScope ps(PK_HIST, st);
  if (range_only) {
    switch (kb) {
      case 1: range_kernel<1><<<grid, 256, 0, st>>>(a); break;
      case 2: range_kernel<2><<<grid, 256, 0, st>>>(a); break;
      case 4: range_kernel<4><<<grid, 256, 0, st>>>(a); break;
      default: range_kernel<8><<<grid, 256, 0, st>>>(a); break;
    }
  } else {
    switch (kb) {
      case 1: top_hist_kernel<1><<<grid, 256, 0, st>>>(a); break;
      case 2: top_hist_kernel<2><<<grid, 256, 0, st>>>(a); break;
      case 4: top_hist_kernel<4><<<grid, 256, 0, st>>>(a); break;
      default: top_hist_kernel<8><<<grid, 256, 0, st>>>(a); break;
    }
  }
  g_launches++;
  return cudaGetLastError();
}

// exact number of records this rank sends to every destination (destination = lut[top bits])
struct DestCountArgs {
  const unsigned char *keys;
  uint32_t stride;
  int64_t n;
  KeyOrder ko;
  int shift;
  const uint8_t *lut;
  unsigned long long *counts;  // [RADIX], zeroed
  int world;
  const uint32_t *bounds;      // [world + 1] splitters (bin indices)
  unsigned long long lo;       // bin = range_bin(ordered key, lo, shift, nb)
  uint32_t nb;
  unsigned long long kbound[7];  // world <= 8: smallest ordered key of rank r+1 (0: every key is at or above it)
  uint32_t never;                // bit r: no key belongs to rank r+1 or higher (kbound[r] unused)
  uint32_t hi_only;              // every used kbound is nonzero with a zero low word: compare high words,
  uint32_t b32m1[7];             // key's high word > b32m1[r]  (0xffffffff: never)
};

template <int KB, int NLD>
__global__ void __launch_bounds__(HIST_THREADS, 1024 / HIST_THREADS) dest_count_kernel(DestCountArgs a) {
  using KT = KeyTile<KB, HIST_THREADS, NLD>;
  __shared__ uint32_t sh[RADIX];
  __shared__ uint32_t s_bound[32];  // upper bin boundary of rank r (exclusive)
  for (int i = threadIdx.x; i < RADIX; i += HIST_THREADS) sh[i] = 0;
  if (threadIdx.x < 32) s_bound[threadIdx.x] = (int)threadIdx.x < a.world ? a.bounds[threadIdx.x + 1] : 0u;
  __syncthreads();
  const int64_t n_tiles = (a.n + KT::TILE - 1) / KT::TILE;
  uint32_t lane_cnt = 0;
  uint32_t n_ge[7] = {0, 0, 0, 0, 0, 0, 0}, n_all = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    KT kt;
    kt.template load<false>(a.keys, a.stride, a.n, tile, a.ko);
    if (a.world <= 8) {
      // one box: lane-private counters n_ge[r] = number of keys at or above the first key of rank r+1 (the
      // splitters as 64-bit key values): 7 compares per key, no cross-lane traffic, no table
      const bool full = kt.valid == (KT::PER_THREAD >= 32 ? 0xffffffffu : ((1u << KT::PER_THREAD) - 1));
      if (a.hi_only && full) {
        // splitters with zero low words (full-width keys): one 32-bit compare per splitter
#pragma unroll
        for (int i = 0; i < KT::PER_THREAD; i++) {
          const uint32_t h = KB == 8 ? (uint32_t)((unsigned long long)kt.u[i] >> 32) : (uint32_t)kt.u[i];
#pragma unroll
          for (int r = 0; r < 7; r++) n_ge[r] += h > a.b32m1[r] ? 1u : 0u;
        }
        n_all += KT::PER_THREAD;
      } else {
#pragma unroll
        for (int i = 0; i < KT::PER_THREAD; i++) {
          const bool v = (kt.valid >> i) & 1;
          const unsigned long long u = (unsigned long long)kt.u[i];
#pragma unroll
          for (int r = 0; r < 7; r++) n_ge[r] += (v && !((a.never >> r) & 1u) && u >= a.kbound[r]) ? 1u : 0u;
          n_all += v ? 1u : 0u;
        }
      }
    } else if (a.world <= 32) {
      // few destinations: no table look-up (a random byte load per key is what would bound this kernel).
      // Lane r counts the keys below the upper boundary of rank r: one compare + ballot per rank and row;
      // the counts per destination are the differences.
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++) {
        const bool v = (kt.valid >> i) & 1;
        const uint32_t bin = v ? range_bin((unsigned long long)kt.u[i], a.lo, a.shift, a.nb) : 0xffffffffu;
        for (int r = 0; r < a.world; r++) {
          const unsigned bal = __ballot_sync(0xffffffffu, bin < s_bound[r]);
          if ((int)(threadIdx.x & 31) == r) lane_cnt += __popc(bal);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++) {
        const bool v = (kt.valid >> i) & 1;
        const unsigned vmask = __ballot_sync(0xffffffffu, v);
        if (v) hist_add<true>(sh, (uint32_t)a.lut[range_bin((unsigned long long)kt.u[i], a.lo, a.shift, a.nb)], vmask);
      }
    }
  }
  if (a.world <= 8) {
    // counts per destination are differences of the cumulative counters
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t above = r < 7 ? n_ge[r] : 0u;        // keys that belong to ranks > r
      const uint32_t from = r == 0 ? n_all : n_ge[r - 1];  // keys that belong to ranks >= r
      uint32_t cnt = from - above;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if ((threadIdx.x & 31) == 0 && r < a.world && cnt) atomicAdd(&sh[r], cnt);
    }
  } else if (a.world <= 32) {
    // lane r holds #keys below bound r+1 (cumulative): difference with the lane before gives rank r's count
    const uint32_t prev = __shfl_up_sync(0xffffffffu, lane_cnt, 1);
    const uint32_t mine = lane_cnt - ((threadIdx.x & 31) ? prev : 0u);
    if ((int)(threadIdx.x & 31) < a.world && mine) atomicAdd(&sh[threadIdx.x & 31], mine);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RADIX; i += HIST_THREADS) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&a.counts[i], (unsigned long long)c);
  }
}

static cudaError_t launch_dest_count(int kb, const DestCountArgs &a, int sm_count, cudaStream_t st) {
  const int64_t tile_keys = (int64_t)HIST_THREADS * hist_nld(kb) * (16 / kb);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a.n + tile_keys - 1) / tile_keys, (int64_t)sm_count * 8));
  ProfScope ps(PK_HIST, st);
  switch (kb) {
    case 1: dest_count_kernel<1, hist_nld(1)><<<grid, HIST_THREADS, 0, st>>>(a); break;
    case 2: dest_count_kernel<2, hist_nld(2)><<<grid, HIST_THREADS, 0, st>>>(a); break;
    case 4: dest_count_kernel<4, hist_nld(4)><<<grid, HIST_THREADS, 0, st>>>(a); break;
    default: dest_count_kernel<8, hist_nld(8)><<<grid, HIST_THREADS, 0, st>>>(a); break;
  }
  g_launches++;
  return cudaGetLastError();
}

// From the smallest / largest sampled ordered key to the binning of the splitter histogram:
// bin(u) = clamp((u - lo) >> shift, 0, 2^bits - 1)  (range_bin(), kernels.cuh).
static void choose_range_bins(unsigned long long lo_key, unsigned long long hi_key, int key_bytes, int bits,
                              unsigned long long *out_lo, int *out_shift) {
  unsigned long long lo = lo_key, hi = hi_key;
  if (lo > hi) lo = hi = 0;  // no keys anywhere
  const unsigned long long span = hi - lo;
  const int span_bits = span ? 64 - __builtin_clzll(span) : 0;
  int shift = std::max(0, span_bits - bits);
  // full-width keys: keep the bins aligned with the key's own leading bits (the local sorts then see shards
  // with constant leading bits and can shift them out)
  if (span_bits > 8 * key_bytes - 2) { lo = 0; shift = 8 * key_bytes - bits; }
  *out_lo = lo;
  *out_shift = shift;
}

// Splitters on bin boundaries.  Rank r ideally ends where the running count reaches r/world of the total; any
// boundary whose prefix is within 1/64 of a rank's share of that target is acceptable (capacities leave
// 1/8), and among those the one with the most trailing zero bits wins: the keys of a shard then agree on as
// many leading bits as possible (for uniform keys and a power-of-two world: exactly log2(world) bits),
// which is what lets the local sort shift them out (Plan::lshift).  Identical on every rank because the
// input is the reduced histogram.
static void compute_splitters(const uint64_t *hist, int bits, int world, uint32_t *bounds) {
  const uint32_t nb = 1u << bits;
  std::vector<uint64_t> prefix((size_t)nb + 1);
  prefix[0] = 0;
  for (uint32_t b = 0; b < nb; b++) prefix[b + 1] = prefix[b] + hist[b];
  const uint64_t total = prefix[nb];
  const uint64_t tol = total / (uint64_t)world / 64;
  bounds[0] = 0;
  for (int r = 1; r < world; r++) {
    const uint64_t target = (uint64_t)(((unsigned __int128)total * (unsigned)r) / (unsigned)world);
    // first boundary whose prefix reaches the target, then the closer of it and the one before
    uint32_t b = (uint32_t)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
    if (b > nb) b = nb;
    if (b > 0 && target - prefix[b - 1] < prefix[b] - target) b--;
    if (b < bounds[r - 1]) b = bounds[r - 1];
    // most aligned acceptable boundary around it
    uint32_t best = b;
    int best_tz = b ? __builtin_ctz(b) : 32;
    for (int dir = -1; dir <= 1; dir += 2) {
      for (int64_t c = (int64_t)b + dir; c >= (int64_t)bounds[r - 1] && c <= (int64_t)nb; c += dir) {
        const uint64_t pc = prefix[(size_t)c];
        const uint64_t dist = pc > target ? pc - target : target - pc;
        if (dist > tol) break;
        const int tz = c ? __builtin_ctz((uint32_t)c) : 32;
        if (tz > best_tz) { best = (uint32_t)c; best_tz = tz; }
      }
    }
    bounds[r] = best;
  }
  bounds[world] = nb;
}

}  // namespace b200sort

// what every rank tells the others before the exchange (all-gathered in one go)
struct MgpuBlob {
  cudaIpcMemHandle_t handle;        // its workspace allocation
  unsigned long long ws_bytes;      // size of that allocation's layout (equal layouts <=> equal shadow offsets)
  long long capacity;
  unsigned long long p2p;           // 1: this rank is willing to use the peer-memory path
  unsigned long long counts[b200sort::RADIX];  // records it sends to every destination
};

struct b200sort_comm {
  ncclComm_t comm = nullptr;
  int world = 0, rank = 0, dev = 0;
  unsigned long long *d_hist = nullptr;   // [2^16] local, then reduced in place
  unsigned long long *d_counts = nullptr; // [world] send counts, [world*world] gathered (NCCL path scratch, barrier word)
  uint8_t *d_lut = nullptr;               // [2^16]
  uint32_t *d_bounds = nullptr;           // [RADIX + 1] splitters
  uint64_t *d_bin_base = nullptr;         // [RADIX]
  int64_t *d_peer_delta = nullptr;        // [RADIX]
  MgpuBlob *d_blob = nullptr;             // [world]
  b200sort::Plan *d_plan = nullptr;
  // peer workspaces mapped into this process (cudaIpcOpenMemHandle), keyed by the handle they were opened from
  std::vector<cudaIpcMemHandle_t> peer_handle;
  std::vector<void *> peer_base;
  bool p2p_failed = false;                // mapping failed once: stay on the NCCL path
  cudaIpcMemHandle_t my_handle{};         // handle of this rank's workspace ...
  void *my_handle_of = nullptr;           // ... taken for this allocation
  uint64_t my_handle_gen = 0;
  bool last_p2p = false;
};

// B200SORT_MGPU_TRACE=1: per-phase device times of every multi-GPU sort on stderr (adds a stream sync at the end)
struct MgpuTrace {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<std::pair<const char *, cudaEvent_t>> ev;
  explicit MgpuTrace(cudaStream_t s) : st(s) {
    const char *e = getenv("B200SORT_MGPU_TRACE");
    on = e && *e == '1';
    mark("start");
  }
  void mark(const char *name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.emplace_back(name, e);
  }
  void report(int rank) {
    if (!on) return;
    cudaStreamSynchronize(st);
    std::string line = "[b200sort mgpu rank " + std::to_string(rank) + "]";
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      char buf[96];
      snprintf(buf, sizeof buf, " %s=%.3f", ev[i].first, ms);
      line += buf;
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ev.front().second, ev.back().second);
    fprintf(stderr, "%s total=%.3f ms\n", line.c_str(), tot);
    for (auto &e : ev) cudaEventDestroy(e.second);
    ev.clear();
  }
};

extern "C" {

int b200sort_mgpu_unique_id(void *out) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!out) return fail(B200SORT_EINVAL, "out is NULL");
  static_assert(sizeof(ncclUniqueId) == B200SORT_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(api.GetUniqueId(&id));
  memcpy(out, &id, sizeof id);
  return 0;
}

int b200sort_mgpu_comm_create(b200sort_comm **out, int world_size, int rank, const void *id_bytes) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!out || !id_bytes || world_size < 1 || world_size > RADIX || rank < 0 || rank >= world_size)
    return fail(B200SORT_EINVAL, "bad communicator arguments");
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  b200sort_comm *c = new b200sort_comm();
  c->world = world_size;
  c->rank = rank;
  CUDA_TRY(cudaGetDevice(&c->dev));
  NCCL_TRY(api.CommInitRank(&c->comm, world_size, id, rank));
  CUDA_TRY(cudaMalloc(&c->d_hist, sizeof(unsigned long long) << MGPU_MAX_BITS));
  CUDA_TRY(cudaMalloc(&c->d_counts, sizeof(unsigned long long) * ((size_t)world_size * (world_size + 1) + 8)));
  CUDA_TRY(cudaMalloc(&c->d_lut, (size_t)1 << MGPU_MAX_BITS));
  CUDA_TRY(cudaMalloc(&c->d_bin_base, sizeof(uint64_t) * RADIX));
  CUDA_TRY(cudaMalloc(&c->d_plan, sizeof(Plan)));
  CUDA_TRY(cudaMalloc(&c->d_peer_delta, sizeof(int64_t) * RADIX));
  CUDA_TRY(cudaMalloc(&c->d_bounds, sizeof(uint32_t) * (RADIX + 1)));
  CUDA_TRY(cudaMalloc(&c->d_blob, sizeof(MgpuBlob) * (size_t)world_size));
  c->peer_handle.resize(world_size);
  c->peer_base.assign(world_size, nullptr);
  for (auto &h : c->peer_handle) memset(&h, 0, sizeof h);
  *out = c;
  return 0;
}

int b200sort_mgpu_comm_destroy(b200sort_comm *c) {
  using namespace b200sort;
  if (!c) return 0;
  if (c->comm) nccl_api().CommDestroy(c->comm);
  for (int p = 0; p < c->world; p++)
    if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  cudaFree(c->d_hist); cudaFree(c->d_counts); cudaFree(c->d_lut); cudaFree(c->d_bin_base); cudaFree(c->d_plan);
  cudaFree(c->d_peer_delta); cudaFree(c->d_blob); cudaFree(c->d_bounds);
  delete c;
  return 0;
}

int b200sort_mgpu_used_p2p(const b200sort_comm *c) { return c && c->last_p2p ? 1 : 0; }

int b200sort_mgpu_range_bins(uint64_t lo_key, uint64_t hi_key, int key_bytes, int bits, uint64_t *out_lo, int *out_shift) {
  using namespace b200sort;
  if (!out_lo || !out_shift || bits < 1 || bits > MGPU_MAX_BITS || (key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8))
    return fail(B200SORT_EINVAL, "bad range arguments");
  unsigned long long lo = 0;
  choose_range_bins(lo_key, hi_key, key_bytes, std::min(bits, 8 * key_bytes), &lo, out_shift);
  *out_lo = lo;
  return 0;
}

int b200sort_mgpu_splitters(const uint64_t *global_hist, int bits, int world_size, uint32_t *out_bounds) {
  using namespace b200sort;
  if (!global_hist || !out_bounds || bits < 1 || bits > MGPU_MAX_BITS || world_size < 1 || world_size > RADIX)
    return fail(B200SORT_EINVAL, "bad splitter arguments");
  compute_splitters(global_hist, bits, world_size, out_bounds);
  return 0;
}

int b200sort_mgpu_plan(const uint64_t *local_hist, int bits, int world_size, const uint32_t *bounds,
                       uint64_t *out_send_counts) {
  using namespace b200sort;
  if (!local_hist || !bounds || !out_send_counts || bits < 1 || bits > MGPU_MAX_BITS || world_size < 1)
    return fail(B200SORT_EINVAL, "bad plan arguments");
  for (int r = 0; r < world_size; r++) {
    uint64_t s = 0;
    for (uint32_t b = bounds[r]; b < bounds[r + 1]; b++) s += local_hist[b];
    out_send_counts[r] = s;
  }
  return 0;
}

int b200sort_mgpu_sort_soa(b200sort_comm *c, void *keys, int key_type, int64_t num_local, int64_t capacity,
                           int ascending, int n_payloads, void *const *payloads, const uint32_t *payload_elem_bytes,
                           int64_t *out_num_local, void *stream_v) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!c || !out_num_local) return fail(B200SORT_EINVAL, "comm/out_num_local is NULL");
  if (capacity < num_local) return fail(B200SORT_EINVAL, "capacity smaller than num_local");
  std::vector<StreamDesc> streams;
  if (int rc = check_soa(keys, key_type, num_local, n_payloads, payloads, payload_elem_bytes, &streams)) return rc;
  if (!keys) return fail(B200SORT_EINVAL, "keys is NULL");
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int kb = key_bytes_of(key_type);
  const int world = c->world;
  const int bits = std::min(MGPU_MAX_BITS, 8 * kb);
  const uint32_t nb = 1u << bits;
  DeviceScope dev_scope;  // the communicator's device, whatever the caller's current device is
  if (dev_scope.enter(c->dev) != 0) return fail(B200SORT_ECUDA, "cannot make device %d current", c->dev);
  DevInfo di;
  if (int rc = dev_info(c->dev, &di)) return rc;
  const KeyOrder ko = make_key_order(key_type, ascending != 0);
  MgpuTrace trace(stream);

  // workspace, laid out for `capacity` records on every rank (so that equal capacities give equal layouts
  // and the local sort below finds what the peers wrote where it expects its shadow arrays)
  uint32_t stage_bytes = (uint32_t)kb;
  StreamSet ss{};
  ss.n_streams = (int)streams.size();
  for (size_t s = 0; s < streams.size(); s++) {
    const uint32_t ck = chunk_for(streams[s].ptr, streams[s].elem_bytes);
    ss.streams[s].chunk_bytes = ck;
    ss.streams[s].chunks_per_elem = streams[s].elem_bytes / ck;
    ss.streams[s].buf[0] = (unsigned char *)streams[s].ptr;
    stage_bytes = std::max(stage_bytes, ck);
  }
  const int cfg = pick_tile_cfg(kb, stage_bytes, di.smem_optin);
  const TileCfg tc = kTileCfgs[cfg];
  const int tile = tc.threads * tc.ipt;
  const int64_t n_ws = std::max<int64_t>(capacity, 1);
  // Large 8-byte-key sorts receive into a third set of arrays (landing): the local sort then runs
  // landing -> shadow -> caller -> shadow -> caller and an even number of passes ends in the caller's arrays
  // without a copy.  The choice depends on nothing but the arguments all ranks share.
  const bool landing = kb == 8 && n_ws >= ((int64_t)1 << std::min<int64_t>(std::max<int64_t>(opt_host_plan_min_log2.load(), 0), 62)) &&
                       opt_mgpu_landing.load() != 0 && opt_mgpu_p2p.load() != 0;
  Layout L;
  make_layout(streams, n_ws, std::min(tile, HYB_MIN_TILE), &L, landing);
  CacheGuard cache_guard;  // (recursive: the local sort below takes it again)
  cache_guard.acquire(c->dev, stream);
  // A rank whose cached workspace is too small is going to free and re-allocate it.  Its peers may still have
  // the old allocation mapped (cudaIpcOpenMemHandle): they have to let go of it first.  The flag travels with
  // the first all-reduce below.
  const bool will_realloc = cached_workspace_bytes(c->dev) < L.total;

  // 1: key range of a sample of all ranks' keys (min / max all-reduced), then the histogram of that range in
  //    2^bits bins on the same sample, all-reduced
  const int64_t sample = std::max<int64_t>(1, num_local >> 24);  // about 2^24 sampled keys at most
  unsigned long long range_h[3] = {~0ull, ~0ull, will_realloc ? 0ull : 1ull};
  CUDA_TRY(cudaMemcpyAsync(c->d_counts, range_h, sizeof range_h, cudaMemcpyHostToDevice, stream));
  TopHistArgs ha{(const unsigned char *)keys, (uint32_t)kb, num_local, ko, 0, c->d_hist, sample, 0ull, nb, c->d_counts};
  if (num_local > 0) CUDA_TRY(launch_top_hist(kb, ha, di.sm_count, stream, /*range_only=*/true));
  NCCL_TRY(api.AllReduce(c->d_counts, c->d_counts, 3, ncclUint64, ncclMin, c->comm, stream));
  CUDA_TRY(cudaMemcpyAsync(range_h, c->d_counts, sizeof range_h, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned long long) * nb, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (range_h[2] == 0) {
    // somebody re-allocates: every rank drops all its mappings, and nobody frees anything before all have
    for (int r = 0; r < world; r++) {
      if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
      c->peer_base[r] = nullptr;
    }
    NCCL_TRY(api.AllReduce(c->d_counts + 4, c->d_counts + 4, 1, ncclUint64, ncclSum, c->comm, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
  }
  void *ws_v = nullptr;
  if (int rc = cached_workspace(c->dev, L.total, &ws_v)) return rc;
  unsigned char *ws = (unsigned char *)ws_v;
  for (size_t s = 0; s < streams.size(); s++) ss.streams[s].buf[1] = ws + L.shadow_off[s];
  unsigned long long lo = 0;
  int shift = 0;
  choose_range_bins(range_h[0], ~range_h[1], kb, bits, &lo, &shift);
  ha.lo = lo; ha.shift = shift;
  if (num_local > 0) CUDA_TRY(launch_top_hist(kb, ha, di.sm_count, stream));
  std::vector<uint64_t> global_hist(nb);
  NCCL_TRY(api.AllReduce(c->d_hist, c->d_hist, nb, ncclUint64, ncclSum, c->comm, stream));
  CUDA_TRY(cudaMemcpyAsync(global_hist.data(), c->d_hist, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  trace.mark("hist+allreduce");

  // splitters (same on every rank) -> destination LUT
  std::vector<uint32_t> bounds(world + 1);
  compute_splitters(global_hist.data(), bits, world, bounds.data());
  std::vector<uint8_t> lut(nb);
  for (int r = 0; r < world; r++)
    for (uint32_t b = bounds[r]; b < bounds[r + 1]; b++) lut[b] = (uint8_t)r;
  CUDA_TRY(cudaMemcpyAsync(c->d_lut, lut.data(), nb, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemcpyAsync(c->d_bounds, bounds.data(), sizeof(uint32_t) * (world + 1), cudaMemcpyHostToDevice, stream));

  // 3: my blob = {workspace handle, layout size, capacity, exact send counts}; all-gather
  MgpuBlob mine{};
  const bool want_p2p = opt_mgpu_p2p.load() != 0 && !c->p2p_failed;
  if (want_p2p) {
    const uint64_t gen = cached_workspace_gen(c->dev);
    if (c->my_handle_of != ws || c->my_handle_gen != gen) {  // (a slow driver call: once per workspace allocation)
      cudaError_t e = cudaIpcGetMemHandle(&c->my_handle, ws);
      if (e != cudaSuccess) { cudaGetLastError(); c->p2p_failed = true; }
      else { c->my_handle_of = ws; c->my_handle_gen = gen; }
    }
    mine.handle = c->my_handle;
  }
  mine.ws_bytes = L.total;
  mine.capacity = capacity;
  mine.p2p = (want_p2p && !c->p2p_failed) ? 1 : 0;
  MgpuBlob *my_slot = c->d_blob + c->rank;
  CUDA_TRY(cudaMemcpyAsync(my_slot, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));  // counts zeroed with it
  trace.mark("host_prep");
  if (num_local > 0) {
    DestCountArgs da{(const unsigned char *)keys, (uint32_t)kb, num_local, ko, shift, c->d_lut, my_slot->counts, world, c->d_bounds, lo, nb, {}, 0u, 1u, {}};
    for (int r = 0; r < 7; r++) {
      // range_bin(u) >= bounds[r+1]  <=>  u >= lo + (bounds[r+1] << shift), with the two clamps of range_bin
      const uint32_t ubr = r + 1 < world ? bounds[r + 1] : nb;
      da.kbound[r] = ubr >= nb ? ~0ull : (ubr == 0 ? 0ull : lo + ((unsigned long long)ubr << shift));
      if (ubr >= nb) da.never |= 1u << r;
      // fast form: high word of the key (the whole key for <= 4-byte keys) against the splitter's high word
      const unsigned long long kbv = da.kbound[r];
      const uint32_t hw = kb == 8 ? (uint32_t)(kbv >> 32) : (uint32_t)kbv;
      if (ubr >= nb) da.b32m1[r] = 0xffffffffu;
      else if (kbv == 0 || hw == 0 || (kb == 8 && (uint32_t)kbv != 0)) da.hi_only = 0;
      else da.b32m1[r] = hw - 1u;
    }
    CUDA_TRY(launch_dest_count(kb, da, di.sm_count, stream));
  }
  trace.mark("count");
  NCCL_TRY(api.AllGather(my_slot, c->d_blob, sizeof(MgpuBlob), ncclUint8, c->comm, stream));
  std::vector<MgpuBlob> blobs(world);
  CUDA_TRY(cudaMemcpyAsync(blobs.data(), c->d_blob, sizeof(MgpuBlob) * world, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  trace.mark("allgather");

  auto m = [&](int src, int dst) -> uint64_t { return blobs[src].counts[dst]; };
  std::vector<uint64_t> send(world);
  for (int r = 0; r < world; r++) send[r] = m(c->rank, r);
  int64_t recv_total = 0;
  std::vector<int64_t> recv_off(world), send_off(world);
  for (int s = 0; s < world; s++) { recv_off[s] = recv_total; recv_total += (int64_t)m(s, c->rank); }
  {
    int64_t o = 0;
    for (int r = 0; r < world; r++) { send_off[r] = o; o += (int64_t)send[r]; }
  }
  // every rank can evaluate every rank's receive total, so all ranks fail together
  for (int r = 0; r < world; r++) {
    int64_t t = 0;
    for (int s = 0; s < world; s++) t += (int64_t)m(s, r);
    if (t > blobs[r].capacity)
      return fail(B200SORT_ENOMEM, "rank %d would receive %lld records, its capacity is %lld", r, (long long)t, (long long)blobs[r].capacity);
  }

  // peer-memory path: everybody willing, equal layouts; map the workspaces that changed since the last sort
  bool p2p = true;
  for (int r = 0; r < world; r++) p2p = p2p && blobs[r].p2p == 1 && blobs[r].ws_bytes == L.total && blobs[r].capacity == capacity;
  if (p2p) {
    bool ok = true;
    for (int r = 0; r < world; r++) {
      if (r == c->rank) { c->peer_base[r] = ws; continue; }
      if (c->peer_base[r] && memcmp(&c->peer_handle[r], &blobs[r].handle, sizeof(cudaIpcMemHandle_t)) == 0) continue;
      if (c->peer_base[r]) { cudaIpcCloseMemHandle(c->peer_base[r]); c->peer_base[r] = nullptr; }
      void *pp = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&pp, blobs[r].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); ok = false; continue; }
      c->peer_base[r] = pp;
      c->peer_handle[r] = blobs[r].handle;
    }
    // (the set of changed handles is the same on every rank up to a rank's own, so all ranks that see a
    //  change agree below; a rank that sees none contributes "ok")
    // agreement: one failed mapping anywhere sends everybody to the NCCL path for good
    unsigned long long flag = ok ? 0ull : 1ull;
    CUDA_TRY(cudaMemcpyAsync(c->d_counts, &flag, sizeof flag, cudaMemcpyHostToDevice, stream));
    NCCL_TRY(api.AllReduce(c->d_counts, c->d_counts, 1, ncclUint64, ncclSum, c->comm, stream));
    CUDA_TRY(cudaMemcpyAsync(&flag, c->d_counts, sizeof flag, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (flag != 0) { c->p2p_failed = true; p2p = false; }
  }
  trace.mark("map");
  c->last_p2p = p2p;

  // 4: partition by destination: one scatter pass; bucket d = what goes to rank d.
  //    peer path : caller arrays -> shadow arrays OF RANK d, at the offset where this rank's records belong
  //    NCCL path : caller arrays -> own shadow arrays, then send/recv
  {
    CUDA_TRY(cudaMemsetAsync(ws + L.ctrl_off, 0, L.ctrl_bytes, stream));
    uint64_t bin_base[RADIX] = {0};
    int64_t peer_delta[RADIX] = {0};
    for (int r = 0; r < RADIX; r++) {
      if (r < world) {
        if (p2p) {
          int64_t o = 0;
          for (int s2 = 0; s2 < c->rank; s2++) o += (int64_t)m(s2, r);  // my region inside rank r's receive range
          bin_base[r] = (uint64_t)o;
          peer_delta[r] = (int64_t)((intptr_t)c->peer_base[r] - (intptr_t)ws);
        } else {
          bin_base[r] = (uint64_t)send_off[r];
        }
      } else {
        bin_base[r] = (uint64_t)num_local;
      }
    }
    Plan plan{};
    plan.final_sel = 1; plan.n_exec = 1;
    CUDA_TRY(cudaMemcpyAsync(c->d_bin_base, bin_base, sizeof bin_base, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_peer_delta, peer_delta, sizeof peer_delta, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_plan, &plan, sizeof plan, cudaMemcpyHostToDevice, stream));
    if (num_local > 0) {
      SweepArgs wa{};
      wa.ss = ss; wa.n = num_local; wa.ko = ko; wa.pass = 0; wa.shift = 0;
      if (p2p && landing)
        for (size_t s2 = 0; s2 < streams.size(); s2++) wa.ss.streams[s2].buf[1] = ws + L.land_off[s2];
      wa.bin_base = c->d_bin_base; wa.lookback = (uint64_t *)(ws + L.lookback_off);
      wa.tile_counter = (uint32_t *)(ws + L.tilectr_off); wa.plan = c->d_plan; wa.tag = 1; wa.stage_bytes = stage_bytes;
      wa.lut = c->d_lut; wa.lut_shift = shift; wa.lut_lo = lo; wa.lut_bins = nb; wa.lut_world = world;
      wa.peer_delta = p2p ? c->d_peer_delta : nullptr;
      CUDA_TRY(launch_sweep(kb, cfg, wa, (num_local + tile - 1) / tile, di.smem_optin, di.sm_count, stream));
    }
    // the host arrays above are stack memory; the copies are stream-ordered but pageable: wait before
    // they go out of scope (this also bounds how far the host runs ahead of the exchange)
    CUDA_TRY(cudaStreamSynchronize(stream));
  }
  trace.mark(p2p ? "partition+exchange" : "partition");

  *out_num_local = recv_total;
  if (p2p) {
    // barrier: nobody reads its shadow arrays before every rank's scatter kernel has finished (the
    // all-reduce is ordered after the local kernel on each rank's stream and completes when all joined)
    NCCL_TRY(api.AllReduce(c->d_counts, c->d_counts, 1, ncclUint64, ncclSum, c->comm, stream));
    trace.mark("barrier");
    // 5: local sort, input in the shadow arrays, result in the caller's arrays
    if (recv_total > 0) {
      // leading key bits this rank's range [bounds[rank], bounds[rank+1]) of top-`bits` values has in common
      int lead = 0;
      if (kb == 8 && lo == 0 && shift == 8 * kb - bits && bounds[c->rank + 1] > bounds[c->rank]) {
        const uint32_t x = bounds[c->rank] ^ (bounds[c->rank + 1] - 1);
        lead = x ? __builtin_clz(x) - (32 - bits) : bits;
      }
      DevSortOpts so;
      so.start_sel = landing ? 0 : 1;
      so.layout_n = n_ws;
      so.hint_lead_bits = lead;
      so.layout_landing = landing;
      so.landing_input = landing;
      int rc = sort_device(key_type, ascending != 0, recv_total, streams, stream, nullptr, 0, so);
      if (rc != 0) return rc;
    }
  } else {
    // all-to-all-v, shadow -> caller arrays
    NCCL_TRY(api.GroupStart());
    for (size_t s = 0; s < streams.size(); s++) {
      const size_t eb = streams[s].elem_bytes;
      for (int p = 0; p < world; p++) {
        const size_t sb = (size_t)send[p] * eb, rb = (size_t)m(p, c->rank) * eb;
        if (sb) NCCL_TRY(api.Send(ss.streams[s].buf[1] + (size_t)send_off[p] * eb, sb, ncclUint8, p, c->comm, stream));
        if (rb) NCCL_TRY(api.Recv(ss.streams[s].buf[0] + (size_t)recv_off[p] * eb, rb, ncclUint8, p, c->comm, stream));
      }
    }
    NCCL_TRY(api.GroupEnd());
    trace.mark("exchange");
    if (recv_total > 1) {
      DevSortOpts so;
      so.layout_n = n_ws;
      so.layout_landing = landing;
      int rc = sort_device(key_type, ascending != 0, recv_total, streams, stream, nullptr, 0, so);
      if (rc != 0) return rc;
    }
  }
  trace.mark("local_sort");
  trace.report(c->rank);
  return 0;
}

}  // extern "C"
