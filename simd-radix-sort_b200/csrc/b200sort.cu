// b200sort.cu -- host side of libb200sort.so: the C ABI of include/b200sort.h, workspace management,
// host-array staging and the launch sequence of a sort.
//
// Replaces (paths under /root/reference): the public sort<> overloads and radixRecursion,
// src/radix_sort.hpp:270-337.  There is deliberately no CPU code path in this file: if CUDA is not
// usable every entry point fails.
#include "../../include/b200sort.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <utility>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "sweep_select.cuh"
#include "hybrid.cuh"

namespace b200sort {

// ------------------------------------------------------------------------------------------------
// errors, options, counters
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static thread_local b200sort_stats g_last_stats{};
static thread_local bool g_have_stats = false;
static std::atomic<uint64_t> g_launches{0};

static std::atomic<int64_t> opt_algo{0};       // 0 auto, 1 LSD, 2 hybrid
static std::atomic<int64_t> opt_tile_cfg{-1};  // -1 auto
static std::atomic<int64_t> opt_allow_skip{1};
static std::atomic<int64_t> opt_allow_reduce{1};
static std::atomic<int64_t> opt_spin_ns{0};
static std::atomic<int64_t> opt_fix_in_pass{1};  // order tile-local segments in the last pass + junction fix instead of the full finish
static std::atomic<int64_t> opt_hist_match{0};
static std::atomic<int64_t> opt_margin_bits{2};
static std::atomic<int64_t> opt_probe_guess{1};
static std::atomic<int64_t> opt_probe_sample{16};  // large sorts: every this-many-th tile feeds the sampled histograms
static std::atomic<int64_t> opt_allow_lshift{1};
static std::atomic<int64_t> opt_junction_table{1};  // junction kernel compares fingerprints left by the last pass
static std::atomic<int64_t> opt_host_pipeline{1};  // host SoA arrays: sort keys + index while the payloads upload
static std::atomic<int64_t> opt_mgpu_landing{1};  // multi-GPU: receive into a third set of arrays (saves the final copy)
static std::atomic<int64_t> opt_mgpu_p2p{1};  // multi-GPU: scatter straight into peer memory (0: NCCL send/recv)
static std::atomic<int64_t> opt_mgpu_refine{1};   // multi-GPU: refine heavy splitter bins / split heavy key values (0: fail with ENOMEM)
static std::atomic<int64_t> opt_mgpu_overlap{0};  // multi-GPU: chunked exchange overlapped with the receivers' first pass (measured: no gain)
static std::atomic<int64_t> opt_mgpu_chunks{4};   // ... in this many chunks
static std::atomic<int64_t> opt_mgpu_wide{1};     // multi-GPU: 16-byte peer stores of element pairs
static std::atomic<int64_t> opt_mgpu_cons_smem_kb{0};   // overlapped first pass: shared memory requested per CTA (bounds its CTAs per SM; 0 = what it needs)
static std::atomic<int64_t> opt_mgpu_persist_x2{0};     // overlapped exchange: CTAs per SM of the looping partition kernel, times two
static std::atomic<int64_t> opt_mgpu_chunk_min_log2{24};  // ... when a rank holds at least 2^this records
static std::atomic<int64_t> opt_host_plan_min_log2{24};  // hybrid sorts of at least 2^this records read the plan back

// optional per-kernel timing (option "profile"): CUDA events around every launch of the last sort
enum ProfKind { PK_HIST = 0, PK_SCAN = 1, PK_SWEEP = 2, PK_COPYBACK = 3, PK_SEGFIX = 4, PK_OTHER = 5 };
struct ProfEntry { int kind; cudaEvent_t e0, e1; int dev; };
static std::atomic<int64_t> opt_profile{0};
static thread_local std::vector<ProfEntry> g_prof;
static thread_local std::map<int, std::vector<cudaEvent_t>> g_prof_pool;  // per device: events belong to one

static void prof_reset() {
  for (auto &p : g_prof) { g_prof_pool[p.dev].push_back(p.e0); g_prof_pool[p.dev].push_back(p.e1); }
  g_prof.clear();
}
static cudaEvent_t prof_event(int dev) {
  auto &pool = g_prof_pool[dev];
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
  bool on; ProfEntry pe; cudaStream_t st;
  ProfScope(int kind, cudaStream_t s) : on(opt_profile.load() != 0), st(s), kind_(kind) {
    if (on) { pe.dev = 0; cudaGetDevice(&pe.dev); pe.kind = kind; pe.e0 = prof_event(pe.dev); pe.e1 = prof_event(pe.dev); cudaEventRecord(pe.e0, st); }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(pe.e1, st); g_prof.push_back(pe); }
    static const bool dbg = getenv("B200SORT_DEBUG_SYNC") != nullptr;  // development: wait for every kernel, say which one
    if (dbg) {
      static const char *names[] = {"hist", "scan", "sweep", "copyback", "segfix", "other"};
      fprintf(stderr, "[b200sort] %s launched ...", names[kind_ < 6 ? kind_ : 5]);
      const cudaError_t e = cudaStreamSynchronize(st);
      fprintf(stderr, " done (%s)\n", cudaGetErrorString(e));
    }
  }
  int kind_ = 5;
};

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(B200SORT_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static int key_bytes_of(int kt) {
  switch (kt) {
    case B200SORT_U8: case B200SORT_I8: return 1;
    case B200SORT_U16: case B200SORT_I16: return 2;
    case B200SORT_U32: case B200SORT_I32: case B200SORT_F32: return 4;
    case B200SORT_U64: case B200SORT_I64: case B200SORT_F64: return 8;
    default: return 0;
  }
}

KeyOrder make_key_order(int key_type, bool ascending) {
  const int kb = key_bytes_of(key_type);
  const uint64_t mask = kb == 8 ? ~0ull : ((1ull << (8 * kb)) - 1);
  const uint64_t sign = 1ull << (8 * kb - 1);
  const bool is_signed = key_type == B200SORT_I8 || key_type == B200SORT_I16 || key_type == B200SORT_I32 ||
                         key_type == B200SORT_I64;
  const bool is_float = key_type == B200SORT_F32 || key_type == B200SORT_F64;
  KeyOrder ko{0, 0, 0, 0, 0};
  if (is_signed || is_float) ko.xor_const = sign;
  if (is_float) ko.neg_xor = mask ^ sign;
  if (!ascending) ko.xor_const ^= mask;
  return ko;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// tile geometries of the scatter kernel
// ------------------------------------------------------------------------------------------------
static size_t sweep_smem_bytes(const TileCfg &c, uint32_t stage_bytes, int nstage = 1, bool fix = false, bool lut = false) {
  const size_t tile = (size_t)c.threads * c.ipt;
  return (size_t)nstage * tile * stage_bytes + (size_t)(c.threads / 32) * RADIX * 4 + RADIX * 8 + RADIX * 4 + 32 * 4 + tile * 3 +
         (fix ? tile : 0) +       // FIX: per-slot displacement
         (lut ? RADIX * 8 : 0);   // LUT: peer byte offsets
}

static std::atomic<int64_t> opt_nstage{0};  // 0 auto, 1 single staging buffer, 2 double-buffered columns
static std::atomic<int64_t> opt_prefetch_cols{1};  // the tile's part of the payload columns is prefetched into L2 when the tile starts
static std::atomic<int64_t> opt_max_chunk{16};     // largest chunk a stream is moved in (ablation: 8 = AoS records as 8-byte columns)
static std::atomic<int64_t> opt_wide_tiles{1};    // 4-byte keys, all chunks <= 4 bytes: 8192-key tiles
static std::atomic<int64_t> opt_tma_keys{1};      // key tiles arrive by one TMA bulk copy per tile (cp.async.bulk + mbarrier)
static std::atomic<int64_t> opt_bytewise{1};      // lean kernels when the host knows the plan (no range reduction / shift)
static std::atomic<int64_t> opt_first_atomic{1};  // first executed pass of a large sort: unstable atomicAdd ranking

inline SweepFn sweep_fn(int kb, int cfg, const SweepSel &s) {
  switch (kb * 2 + cfg) {
    case 2: return sweep_fn_inst<1, 0>(s);
    case 3: return sweep_fn_inst<1, 1>(s);
    case 4: return sweep_fn_inst<2, 0>(s);
    case 5: return sweep_fn_inst<2, 1>(s);
    case 8: return sweep_fn_inst<4, 0>(s);
    case 9: return sweep_fn_inst<4, 1>(s);
    case 10: return sweep_fn_inst<4, 2>(s);
    case 16: return sweep_fn_inst<8, 0>(s);
    default: return sweep_fn_inst<8, 1>(s);
  }
}

// first_pass_unordered: the caller knows that this launch is the first executed pass of the sort and that its
// digit's histogram is not skewed -- the unstable ranking may be used
// smem_floor: request at least this much dynamic shared memory (bounds how many CTAs of this launch an SM holds:
// the overlapped first pass of the multi-GPU sort must leave room for the partition kernel's CTAs)
static cudaError_t launch_sweep(int kb, int cfg, const SweepArgs &a, int64_t n_tiles, size_t smem_optin, int sm_count, cudaStream_t st,
                                bool first_pass_unordered = false, size_t smem_floor = 0, int64_t grid_cap = 0) {
  bool any = false;  // a stream with 1- or 2-byte chunks in the move loop needs the ANYCHUNK instantiation
  int n_cols = 0;
  const bool soa = a.ss.streams[0].chunk_bytes * a.ss.streams[0].chunks_per_elem == (uint32_t)kb;
  for (int s = soa ? 1 : 0; s < a.ss.n_streams; s++) {
    any = any || a.ss.streams[s].chunk_bytes < 4;
    n_cols += (int)a.ss.streams[s].chunks_per_elem;
  }
  n_cols += soa ? 1 : 0;
  const TileCfg tc = kTileCfgs[cfg];
  const bool lut = a.part.key != nullptr;
  const bool fix = !lut && a.fix_cut != 0;  // (the caller only sets fix_cut where the FIX instantiation exists)
  int rank = RANK_BALLOT;
  // digits are whole bytes of the raw key when the host knows that the plan has no range reduction and no shift
  const bool bytewise = !lut && a.plan_in_args != 0 && a.arg_sub == 0 && a.arg_lshift == 0 && opt_bytewise.load() != 0;
  if (first_pass_unordered && a.plan_in_args != 0 && !lut && !fix && opt_first_atomic.load() != 0) rank = RANK_ATOMIC;
  // double-buffer the columns when there is more than one and minb CTAs still fit on an SM
  int nstage = (int)opt_nstage.load();
  if (nstage != 1 && nstage != 2)
    nstage = (n_cols >= 2 && (sweep_smem_bytes(tc, a.stage_bytes, 2) + 1024) * tc.minb <= smem_optin + 1024) ? 2 : 1;
  if (nstage == 2 && sweep_smem_bytes(tc, a.stage_bytes, 2) > smem_optin) nstage = 1;
  if (lut || fix) nstage = 1;
  SweepSel sel{nstage, any || lut, lut, fix, rank, bytewise};
  SweepArgs a2 = a;
  // TMA bulk load of the key tile: SoA keys whose arrays (both sides of the ping-pong) are 16-byte aligned
  a2.tma_keys = (opt_tma_keys.load() != 0 && soa && (((uintptr_t)a.ss.streams[0].buf[0] | (uintptr_t)a.ss.streams[0].buf[1]) & 15) == 0 &&
                 ((size_t)tc.threads * tc.ipt * kb) % 16 == 0) ? 1u : 0u;
  SweepFn k = sweep_fn(kb, cfg, sel);
  const size_t smem = std::min(std::max(sweep_smem_bytes(tc, a.stage_bytes, nstage, fix, lut), smem_floor), smem_optin);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // one shared-memory carve-out for every instantiation: kernels that want different L1 / shared-memory splits
  // cannot be resident on an SM together (the overlapped exchange runs two of them side by side)
  // (only for the kernels of the overlapped exchange: a maximal carve-out leaves the other sorts 28 KB of L1,
  //  which costs them 4 % -- 38.2 vs 36.8 ms at 1e9 records)
  if (smem_floor != 0 || grid_cap != 0 || opt_mgpu_overlap.load() != 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  ProfScope ps(PK_SWEEP, st);
  (void)sm_count;
  // (grid_cap: the partition pass of the overlapped exchange runs with fewer, looping CTAs)
  // L2 prefetch of the tile's part of the first column after the keys (16-byte aligned on both sides)
  a2.prefetch_cols = 0;
  {
    const int s2 = a2.tma_keys ? 1 : 0;
    if (opt_prefetch_cols.load() != 0 && s2 < a.ss.n_streams) {
      const Stream &st2 = a.ss.streams[s2];
      const size_t bytes = (size_t)tc.threads * tc.ipt * st2.chunk_bytes * st2.chunks_per_elem;
      if ((((uintptr_t)st2.buf[0] | (uintptr_t)st2.buf[1] | bytes) & 15) == 0) {
        a2.prefetch_cols = 1; a2.prefetch_bytes = (uint32_t)bytes;
        a2.prefetch_ptr[0] = st2.buf[0]; a2.prefetch_ptr[1] = st2.buf[1];
      }
    }
  }
  k<<<(unsigned)(grid_cap > 0 ? std::min<int64_t>(n_tiles, grid_cap) : n_tiles), tc.threads, smem, st>>>(a2);
  g_launches++;
  return cudaGetLastError();
}

constexpr int HIST_THREADS = 256;  // key-only sweeps: 256 threads x NLD x 16 B per tile, 4 CTAs per SM
constexpr int hist_nld(int kb) { return kb == 1 ? 2 : (kb >= 4 ? 8 : 4); }  // at most 32 keys per thread; 128 bytes in flight per thread for 8-byte keys

template <int KB, int NLD>
static cudaError_t launch_hist_t(const HistArgs &a, int grid, bool use_match, int probe, cudaStream_t st) {
  ProfScope ps(PK_HIST, st);
  if (probe == 2) {
    minmax_kernel<KB, HIST_THREADS, NLD><<<grid, HIST_THREADS, 0, st>>>(a);
  } else if (probe) {
    if (use_match) probe_kernel<KB, HIST_THREADS, NLD, true><<<grid, HIST_THREADS, 0, st>>>(a);
    else probe_kernel<KB, HIST_THREADS, NLD, false><<<grid, HIST_THREADS, 0, st>>>(a);
  } else {
    if (use_match) hist_kernel<KB, HIST_THREADS, NLD, true><<<grid, HIST_THREADS, 0, st>>>(a);
    else hist_kernel<KB, HIST_THREADS, NLD, false><<<grid, HIST_THREADS, 0, st>>>(a);
  }
  g_launches++;
  return cudaGetLastError();
}

// keys per tile of the key-only sweeps over this key array
static int64_t hist_tile_keys(int kb, const void *keys, uint32_t stride) {
  // (8-byte keys in a dense, 16-byte aligned array: eight 16-byte vectors in flight per thread -- probe 1.66 ->
  //  1.33 ms at 1e9 keys; records with the key inside are read one key per load, where more per thread is slower)
  const bool dense = stride == (uint32_t)kb && (((uintptr_t)keys) & 15) == 0;
  const int nld = (kb == 8 && !dense) ? 4 : hist_nld(kb);  // (4-byte keys: 8 vectors = 32 keys per thread either way)
  return (int64_t)HIST_THREADS * nld * (16 / kb);
}

cudaError_t launch_hist(int kb, const HistArgs &a, int sm_count, int probe, cudaStream_t st) {
  const int64_t tile_keys = hist_tile_keys(kb, a.keys, a.stride);
  const int64_t tiles = (a.n + tile_keys - 1) / tile_keys;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count * 8);
  const bool m = opt_hist_match.load() != 0;
  switch (kb) {
    case 1: return launch_hist_t<1, hist_nld(1)>(a, grid, m, probe, st);
    case 2: return launch_hist_t<2, hist_nld(2)>(a, grid, m, probe, st);
    case 4: return launch_hist_t<4, hist_nld(4)>(a, grid, m, probe, st);
    default:
      return tile_keys == (int64_t)HIST_THREADS * 4 * 2 ? launch_hist_t<8, 4>(a, grid, m, probe, st)
                                                        : launch_hist_t<8, hist_nld(8)>(a, grid, m, probe, st);
  }
}

// ------------------------------------------------------------------------------------------------
// device properties + per-device workspace cache
// ------------------------------------------------------------------------------------------------
struct DevInfo { int sm_count = 0; size_t smem_optin = 0; bool ok = false; };
static std::mutex g_mu;
static std::map<int, DevInfo> g_dev;
struct CacheEntry { void *ptr = nullptr; size_t bytes = 0; uint64_t gen = 0; };  // gen: bumped by every (re)allocation
static std::map<int, CacheEntry> g_cache;

// The library-owned caches (workspace, host staging) are one allocation per device.  Sorts that use them are
// serialised: a per-device (recursive) mutex covers the launching of a sort, and the stream of the next sort
// waits for an event recorded at the end of the previous one, so that two host threads, or two streams, never
// have kernels in flight on the same scratch memory.  Callers who want concurrent sorts pass their own workspace.
struct CacheGuardState { std::recursive_mutex mu; cudaEvent_t last_use = nullptr; };
static CacheGuardState &cache_guard_state(int dev) {
  static std::map<int, CacheGuardState> g_guard;
  std::lock_guard<std::mutex> lk(g_mu);
  return g_guard[dev];  // (std::map nodes are stable)
}
struct CacheGuard {
  CacheGuardState *st = nullptr;
  cudaStream_t stream = nullptr;
  void acquire(int dev, cudaStream_t s) {
    st = &cache_guard_state(dev);
    stream = s;
    st->mu.lock();
    if (st->last_use == nullptr) cudaEventCreateWithFlags(&st->last_use, cudaEventDisableTiming);
    else cudaStreamWaitEvent(stream, st->last_use, 0);
  }
  ~CacheGuard() {
    if (!st) return;
    cudaEventRecord(st->last_use, stream);
    st->mu.unlock();
  }
};

static int dev_info(int dev, DevInfo *out) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevInfo &d = g_dev[dev];
  if (!d.ok) {
    int v = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    d.sm_count = v;
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    d.smem_optin = (size_t)v;
    d.ok = true;
  }
  *out = d;
  return 0;
}

static int cached_workspace(int dev, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(g_mu);
  CacheEntry &c = g_cache[dev];
  if (c.bytes < bytes) {
    if (c.ptr) {
      CUDA_TRY(cudaDeviceSynchronize());
      CUDA_TRY(cudaFree(c.ptr));
      c.ptr = nullptr;
      c.bytes = 0;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(B200SORT_ENOMEM, "cudaMalloc of %zu workspace bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    c.ptr = p;
    c.bytes = bytes;
    c.gen++;
  }
  *out = c.ptr;
  return 0;
}

static size_t cached_workspace_bytes(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_cache[dev].bytes;
}

static uint64_t cached_workspace_gen(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_cache[dev].gen;
}

// ------------------------------------------------------------------------------------------------
// job description
// ------------------------------------------------------------------------------------------------
struct StreamDesc { void *ptr; uint32_t elem_bytes; };

struct Layout {
  size_t shadow_off[MAX_STREAMS];
  size_t land_off[MAX_STREAMS];
  size_t ctrl_off, ctrl_bytes;       // zeroed at the start of every sort
  size_t ghist_off, ghist2_off, probe_off, tilectr_off, plan_off, binbase_off, lookback_off, hyb_off, jtable_off, flags_off;
  size_t total;
  int64_t n_tiles;
};

static int pick_tile_cfg(int kb, uint32_t stage_bytes, size_t smem_optin, int64_t n) {
  int cfg = (int)opt_tile_cfg.load();
  const bool automatic = cfg < 0 || cfg >= kNumTileCfgs;
  // (wide tiles only where there are plenty of them: 10^6 records are 123 wide tiles for 148 SMs -- 0.18 vs 0.13 ms)
  if (automatic) cfg = (kb == 4 && stage_bytes <= 4 && n >= ((int64_t)1 << 24) && opt_wide_tiles.load() != 0) ? kWideTileCfg : kDefaultTileCfg;
  if (cfg == kWideTileCfg && (kb != 4 || stage_bytes > 4)) cfg = kDefaultTileCfg;  // (only instantiated for that shape)
  // fall back to the smaller tile if the staging buffer would not fit
  if (sweep_smem_bytes(kTileCfgs[cfg], stage_bytes) > smem_optin) cfg = 1;
  return cfg;
}

static void make_layout(const std::vector<StreamDesc> &streams, int64_t n, int tile, Layout *L, bool landing = false) {
  size_t off = 0;
  for (size_t s = 0; s < streams.size(); s++) {
    L->shadow_off[s] = off;
    off = align_up(off + (size_t)n * streams[s].elem_bytes, 256);
  }
  L->n_tiles = (n + tile - 1) / tile;
  // hybrid path tiles may be smaller than the sweep tiles; size the look-back for the smallest tile used
  L->ctrl_off = off;
  L->ghist_off = off;               off += (size_t)8 * RADIX * 8;   // sampled histograms (probe)
  L->ghist2_off = off;              off += (size_t)8 * RADIX * 8;   // exact histograms, one digit position at a time
  L->probe_off = off;               off += 64;
  L->tilectr_off = off;             off += MAX_PASSES * 4;
  L->plan_off = off;                off = align_up(off + sizeof(Plan), 256);
  L->hyb_off = off;                 off = align_up(off + sizeof(HybridCtrl), 256);
  L->lookback_off = off;            off += (size_t)L->n_tiles * RADIX * 8;
  L->ctrl_bytes = off - L->ctrl_off;
  L->binbase_off = off;             off += (size_t)MAX_PASSES * RADIX * 8;
  L->jtable_off = off;              off += (size_t)L->n_tiles * RADIX * 8;   // FIX pass -> junction_fix_kernel
  off = align_up(off, 256);
  L->flags_off = off;               off += 4096;  // multi-GPU: arrival flags written by the peers (never zeroed: epochs)
  // multi-GPU: a third copy of every stream, where the peers deliver this rank's records (see mgpu.cuh)
  for (size_t s = 0; s < streams.size(); s++) {
    L->land_off[s] = off;
    if (landing) off = align_up(off + (size_t)n * streams[s].elem_bytes, 256);
  }
  L->total = align_up(off, 256);
}

static uint32_t chunk_for(const void *p, uint32_t elem) {
  // largest power of two <= 16 dividing both the address and the element size
  uint32_t c = 16;
  while (c > 1 && ((elem % c) != 0 || ((uintptr_t)p % c) != 0)) c >>= 1;
  return c;
}

// Sort arrays that are all in device memory.  streams[0] carries the key at offset 0.
struct DevSortOpts {
  int start_sel = 0;         // 1: the input lies in the workspace's shadow arrays, the result still goes to the caller's
  int64_t layout_n = 0;      // lay the workspace out for this many records (>= n)
  int hint_lead_bits = 0;    // leading bits the caller expects all keys to agree on (steers the probe's guess)
  bool layout_landing = false;  // the layout has landing arrays (make_layout(..., landing))
  bool landing_input = false;   // the input lies in the landing arrays: the first executed pass reads it from
                                // there (landing -> shadow -> caller -> ...), so that an even number of passes
                                // ends in the caller's arrays without a copy
  int force_algo = 0;           // 1 / 2: use this algorithm whatever option "algo" says (the hybrid path's fall-back)
  // Multi-GPU sort with the exchange overlapped (mgpu.cuh): the plan is fixed by the caller (hybrid: digit
  // positions forced_cut .. 7 of the key shifted left by forced_lshift are swept, no range reduction), the
  // control block has been cleared by the caller, and the FIRST executed pass has already been run by the
  // caller (landing arrays -> shadow arrays, its successor's digit counted into the exact histograms).
  bool forced = false;
  uint32_t forced_lshift = 0, forced_cut = 0;
  // Partial-sort mode (B200SORT_CMP_NONE with a threshold >= CMP_NONE_MIN_THRESH): the segments below the plan's
  // cut need not be ordered as long as none has more than this many keys (0 = full sort)
  int64_t cmp_none_thresh = 0;
};
constexpr int64_t CMP_NONE_MIN_THRESH = 8;  // below it nearly every sort would need the finish anyway: full sort

// one event per (host thread, device) for the plan read-back of large sorts: an event belongs to the device it was
// created on, and two host threads sorting on one device (with their own workspaces) must not share one
static cudaEvent_t plan_event_of(int dev) {
  static thread_local std::map<int, cudaEvent_t> t_plan_events;
  cudaEvent_t &e = t_plan_events[dev];
  if (e == nullptr && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) e = nullptr;
  return e;
}
static int sort_device(int key_type, bool ascending, int64_t n, const std::vector<StreamDesc> &streams,
                       cudaStream_t stream, void *workspace, size_t workspace_bytes, const DevSortOpts &xo = DevSortOpts()) {
  const int start_sel = xo.start_sel;
  const int64_t layout_n = xo.layout_n;
  const int hint_lead_bits = xo.hint_lead_bits;
  const int kb = key_bytes_of(key_type);
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  DevInfo di;
  if (int rc = dev_info(dev, &di)) return rc;

  uint32_t stage_bytes = (uint32_t)kb, rec_bytes = 0;
  StreamSet ss{};
  ss.n_streams = (int)streams.size();
  for (size_t s = 0; s < streams.size(); s++) {
    const uint32_t c = std::min<uint32_t>(chunk_for(streams[s].ptr, streams[s].elem_bytes), (uint32_t)std::max<int64_t>(opt_max_chunk.load(), 1));
    ss.streams[s].chunk_bytes = c;
    ss.streams[s].chunks_per_elem = streams[s].elem_bytes / c;
    ss.streams[s].buf[0] = (unsigned char *)streams[s].ptr;
    stage_bytes = std::max(stage_bytes, c);
    rec_bytes += streams[s].elem_bytes;
  }
  if (streams[0].elem_bytes != (uint32_t)kb && ((uintptr_t)streams[0].ptr % kb) != 0)
    return fail(B200SORT_EINVAL, "record array is not aligned to its key type");
  if (((uintptr_t)streams[0].ptr % kb) != 0) return fail(B200SORT_EINVAL, "key array is not aligned to its key type");

  const int cfg = pick_tile_cfg(kb, stage_bytes, di.smem_optin, n);
  const TileCfg tc = kTileCfgs[cfg];
  const int tile = tc.threads * tc.ipt;
  const size_t smem = sweep_smem_bytes(tc, stage_bytes);
  if (smem > di.smem_optin) return fail(B200SORT_ECUDA, "device offers %zu B of shared memory, %zu needed", di.smem_optin, smem);

  Layout L;
  make_layout(streams, std::max(n, layout_n), std::min(tile, HYB_MIN_TILE), &L, xo.layout_landing);
  void *const caller_workspace = workspace;
  CacheGuard cache_guard;  // released (and its event recorded) when this sort has been launched
  if (workspace == nullptr) {
    cache_guard.acquire(dev, stream);
    if (int rc = cached_workspace(dev, L.total, &workspace)) return rc;
  } else if (workspace_bytes < L.total) {
    return fail(B200SORT_ENOMEM, "workspace of %zu bytes given, %zu needed", workspace_bytes, L.total);
  } else if (((uintptr_t)workspace % 256) != 0) {
    return fail(B200SORT_EINVAL, "workspace must be 256-byte aligned");
  }
  unsigned char *ws = (unsigned char *)workspace;
  for (size_t s = 0; s < streams.size(); s++) ss.streams[s].buf[1] = ws + L.shadow_off[s];
  uint64_t *ghist = (uint64_t *)(ws + L.ghist_off);
  uint32_t *tile_counter = (uint32_t *)(ws + L.tilectr_off);
  Plan *plan = (Plan *)(ws + L.plan_off);
  uint64_t *lookback = (uint64_t *)(ws + L.lookback_off);

  const uint64_t launches_before = g_launches.load();
  if (!xo.forced) prof_reset();
  if (!xo.forced) CUDA_TRY(cudaMemsetAsync(ws + L.ctrl_off, 0, L.ctrl_bytes, stream));

  const KeyOrder ko = make_key_order(key_type, ascending);

  int algo = xo.force_algo != 0 ? xo.force_algo : (int)opt_algo.load();
  if (xo.forced) algo = 2;
  if (algo == 0) algo = (kb == 8 && n >= HYB_MIN_N) ? 2 : 1;
  if (algo == 2 && kb != 8) algo = 1;

  b200sort_stats stt{};
  stt.num = n;
  stt.record_bytes = rec_bytes;
  stt.key_bytes = (uint32_t)kb;
  stt.algo = (uint32_t)algo;

  const bool hybrid = algo == 2;
  {
    // ---- one histogram sweep over all digit positions, the bucket-offset scan (+ pass plan), one
    //      scatter pass per digit position (skipped ones return at once), segment finish, copy-back ----
    uint64_t *ghist_exact = (uint64_t *)(ws + L.ghist2_off);
    ProbeOut *probe = (ProbeOut *)(ws + L.probe_off);
    // large sorts (any algorithm): the host reads the plan back and launches exactly what executes, with the
    // plan's values as kernel arguments and the lean kernel instantiations they allow
    const bool big = xo.forced || n >= (int64_t)1 << std::min<int64_t>(std::max<int64_t>(opt_host_plan_min_log2.load(), 0), 62);
    // Input in the landing arrays (multi-GPU): only the large hybrid flow knows on the host which pass runs
    // first; everything else simply starts with a copy into the caller's arrays.
    bool landing = xo.landing_input;
    auto copy_landing_to_caller = [&]() -> int {
      for (size_t s2 = 0; s2 < streams.size(); s2++)
        CUDA_TRY(cudaMemcpyAsync(streams[s2].ptr, ws + L.land_off[s2], (size_t)n * streams[s2].elem_bytes, cudaMemcpyDeviceToDevice, stream));
      return 0;
    };
    if (landing && !big) {
      if (int rc = copy_landing_to_caller()) return rc;
      landing = false;
    }
    StreamSet ss_in = ss;  // what the key sweeps and the first executed pass read
    if (landing)
      for (size_t s2 = 0; s2 < streams.size(); s2++) ss_in.streams[s2].buf[0] = ws + L.land_off[s2];
    HistArgs ha{};
    ha.keys = ss_in.streams[0].buf[start_sel];
    ha.stride = streams[0].elem_bytes;
    ha.n = n; ha.ko = ko; ha.digit_mask = (1u << kb) - 1; ha.ghist = ghist; ha.probe = probe;
    {
      const int64_t tile_keys = hist_tile_keys(kb, ha.keys, ha.stride);
      // entropies from one key per thread of every 16th tile (option probe_sample) once there are plenty of tiles
      const bool sparse = (n / tile_keys) >= 8192;
      ha.sample = sparse ? (uint32_t)std::max<int64_t>(1, opt_probe_sample.load()) : 1;
      ha.sample_one = sparse ? 1 : 0;
    }
    // Large sorts leave the exact key range to a sweep of its own that only runs when the sampled range says
    // range reduction may pay; the others get it from the probe.
    ha.with_minmax = big ? 0 : 1;
    // The digit position the first pass will sweep if the keys are (close to) uniformly distributed: the
    // probe counts it exactly, and if the plan comes out that way hist_kernel is not needed.
    {
      uint32_t guess = 0, guess_l = 0;
      if (hybrid) {
        // hint_lead_bits: leading bits the caller expects all keys to agree on (a shard of the multi-GPU sort)
        const double need = std::log2((double)n) + (double)opt_margin_bits.load();
        const int swept = (int)std::ceil(need / 8.0);
        const int shift_l = opt_allow_lshift.load() != 0 ? (hint_lead_bits & 7) : 0;
        const int cut = 8 - hint_lead_bits / 8 - swept;
        if (cut >= 2) { guess = (uint32_t)cut; guess_l = (uint32_t)shift_l; }
      }
      // (digit-by-digit path: the least significant digit, guess = 0, is the first pass unless it is constant)
      ha.guess_lshift = guess_l;
      ha.guess_p1 = opt_probe_guess.load() != 0 ? guess + 1 : 0;
      ha.ghist_exact = ghist_exact;
    }
    if (!xo.forced) CUDA_TRY(launch_hist(kb, ha, di.sm_count, /*probe=*/1, stream));

    ScanArgs sa{};
    sa.ghist = ghist; sa.probe = probe; sa.plan = plan; sa.n = n; sa.n_passes = kb;
    sa.allow_skip = (int)opt_allow_skip.load();
    sa.hybrid = hybrid ? 1 : 0;
    sa.allow_reduce = (int)opt_allow_reduce.load();
    sa.margin_bits = (float)opt_margin_bits.load();
    sa.have_minmax = (int)ha.with_minmax;
    sa.allow_lshift = (int)opt_allow_lshift.load();
    sa.start_sel = (uint32_t)start_sel;
    sa.guess_p1 = ha.guess_p1; sa.guess_lshift = ha.guess_lshift; sa.ghist_exact = ghist_exact;
    auto launch_scan = [&]() -> int {
      {
        ProfScope ps(PK_SCAN, stream);
        scan_kernel<<<1, RADIX, 0, stream>>>(sa);
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      return 0;
    };
    if (!xo.forced)
      if (int rc = launch_scan()) return rc;

    // Large hybrid sorts read the plan back (small D2H, one host wait) and launch only what executes: the
    // exact min/max sweep and a second planning step if asked for, the histogram kernel unless the probe's
    // guess was right, the passes that are not skipped.  Smaller sorts launch everything; what is not
    // needed returns at once.
    Plan hplan{};
    bool have_plan = false;
    cudaEvent_t plan_event = nullptr;
    HistArgs hb = ha;
    hb.ghist = ghist_exact; hb.plan = plan; hb.probe = nullptr;
    uint32_t hist_sweeps = 1;
    if (xo.forced) {
      // the caller's plan: sweep digit positions forced_cut .. kb-1 in LSD order; the first of them is done
      hist_sweeps = 0;
      uint32_t sel = 0, prev = 0;
      bool any_prev = false;
      for (int p = 0; p < kb; p++) {
        hplan.skip[p] = (uint32_t)p < xo.forced_cut ? 1u : 0u;
        hplan.src_sel[p] = sel;
        if (!hplan.skip[p]) {
          sel ^= 1u;
          hplan.n_exec++;
          if (!any_prev) hplan.first_exec_p1 = (uint32_t)p + 1; else hplan.next_exec_p1[prev] = (uint32_t)p + 1;
          prev = (uint32_t)p;
          any_prev = true;
        }
      }
      hplan.final_sel = sel;
      hplan.cut_digit = xo.forced_cut;
      hplan.lshift = xo.forced_lshift;
      hplan.hist_done = 1;
      CUDA_TRY(cudaMemcpyAsync(plan, &hplan, sizeof hplan, cudaMemcpyHostToDevice, stream));
      have_plan = true;
    } else if (big) {
      plan_event = plan_event_of(dev);
      if (plan_event == nullptr) return fail(B200SORT_ECUDA, "cudaEventCreate failed on device %d", dev);
      CUDA_TRY(cudaMemcpyAsync(&hplan, plan, sizeof hplan, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaEventRecord(plan_event, stream));
      CUDA_TRY(cudaEventSynchronize(plan_event));
      if (hplan.want_minmax) {
        CUDA_TRY(launch_hist(kb, ha, di.sm_count, /*minmax=*/2, stream));
        hist_sweeps++;
        sa.have_minmax = 1;
        if (int rc = launch_scan()) return rc;
        CUDA_TRY(cudaMemcpyAsync(&hplan, plan, sizeof hplan, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaEventRecord(plan_event, stream));
        CUDA_TRY(cudaEventSynchronize(plan_event));
      }
      have_plan = true;
      if (!hplan.hist_done) {
        CUDA_TRY(launch_hist(kb, hb, di.sm_count, /*probe=*/0, stream));
        hist_sweeps++;
      }
    } else {
      // exact histogram of the first executed pass only; every pass counts its successor's digit
      CUDA_TRY(launch_hist(kb, hb, di.sm_count, /*probe=*/0, stream));
      hist_sweeps++;
    }

    const int64_t n_tiles = (n + tile - 1) / tile;
    HybridCtrl *ctrl = (HybridCtrl *)(ws + L.hyb_off);
    // With the plan on the host, an 8-byte-key SoA sort lets the LAST pass order the final segments each tile
    // holds and repairs the tile-straddling ones with junction_fix_kernel; the full segment finish then only
    // runs if one of them reports a run that is too long.
    const bool soa = streams[0].elem_bytes == (uint32_t)kb;
    // partial-sort mode: no ordering of the final segments at all, only a check that none exceeds the threshold
    const bool partial = have_plan && hplan.cut_digit != 0 && xo.cmp_none_thresh >= CMP_NONE_MIN_THRESH;
    const bool use_fix = !partial && have_plan && hplan.cut_digit != 0 && (soa || ss.streams[0].chunk_bytes >= (uint32_t)kb) && cfg == kDefaultTileCfg && opt_fix_in_pass.load() != 0;
    int last_pass = -1;
    bool first_exec = true;
    if (landing && hplan.n_exec == 0) {  // nothing will move the records: deliver them
      if (int rc = copy_landing_to_caller()) return rc;
    }
    for (int p = 0; p < kb; p++) {
      if (have_plan && hplan.skip[p]) continue;
      SweepArgs wa{};
      wa.ss = (landing && first_exec) ? ss_in : ss; wa.n = n;
      const bool was_first = first_exec;
      first_exec = false;
      if (xo.forced && was_first) continue;  // run by the caller, overlapped with the exchange
      wa.ko = ko; wa.pass = p; wa.shift = p * RADIX_BITS;
      wa.bin_base = nullptr; wa.ghist = ghist_exact;
      wa.lookback = lookback; wa.tile_counter = tile_counter; wa.plan = plan;
      wa.tag = (uint32_t)(p + 1); wa.stage_bytes = stage_bytes; wa.spin_ns = (uint32_t)opt_spin_ns.load();
      if (have_plan) {
        wa.plan_in_args = 1; wa.arg_sel = hplan.src_sel[p]; wa.arg_next_p1 = hplan.next_exec_p1[p];
        wa.arg_next_skewed = hplan.next_exec_p1[p] ? hplan.skewed[hplan.next_exec_p1[p] - 1] : 0; wa.arg_sub = hplan.sub; wa.arg_lshift = hplan.lshift;
        if (use_fix && hplan.next_exec_p1[p] == 0) { wa.fix_cut = hplan.cut_digit; wa.fix_flag = &ctrl->flags[1]; wa.jtable = (uint64_t *)(ws + L.jtable_off); last_pass = p; }
      }
      // the first executed pass may rank its keys in any order (nothing has been established yet)
      const bool unordered = have_plan && was_first && hplan.skewed[p] == 0;
      CUDA_TRY(launch_sweep(kb, cfg, wa, n_tiles, di.smem_optin, di.sm_count, stream, unordered));
    }
    auto launch_segfix = [&](const uint32_t *gate) -> int {
      SegfixArgs fa{};
      fa.ss = ss; fa.n = n; fa.ko = ko; fa.plan = plan; fa.ctrl = ctrl; fa.gate = gate;
      if (have_plan) { fa.plan_in_args = 1; fa.arg_cut = hplan.cut_digit; fa.arg_sel = hplan.final_sel; fa.arg_sub = hplan.sub; fa.arg_lshift = hplan.lshift; }
      bool any = false;
      for (int s = 0; s < ss.n_streams; s++) any = any || ss.streams[s].chunk_bytes < 4;
      const unsigned grid = (unsigned)std::min<int64_t>((n + SF_FT - 1) / SF_FT, (int64_t)di.sm_count * 16);
      {
        ProfScope ps(PK_SEGFIX, stream);
        if (any) segfix_kernel<8, true><<<grid, SF_THREADS, 0, stream>>>(fa);
        else segfix_kernel<8, false><<<grid, SF_THREADS, 0, stream>>>(fa);
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      return 0;
    };
    auto launch_copyback = [&](bool force, const uint32_t *skip_if) -> int {
      CopyBackArgs ca{};
      ca.ss = ss; ca.n = n; ca.plan = plan; ca.force = force ? 1 : 0; ca.skip_if = skip_if;
      {
        ProfScope ps(PK_COPYBACK, stream);
        copyback_kernel<<<di.sm_count * 8, 256, 0, stream>>>(ca);
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      return 0;
    };
    bool finish_ran = false, fix_flow = false;
    if (last_pass >= 0) {
      JunctionArgs ja{};
      ja.ss = ss; ja.n = n; ja.ko = ko; ja.ko.sub = hplan.sub; ja.ko.lshift = hplan.lshift; ja.lookback = lookback; ja.n_tiles = n_tiles;
      ja.tag = (uint32_t)(last_pass + 1); ja.cut = hplan.cut_digit; ja.sel = hplan.final_sel; ja.flag = &ctrl->flags[1];
      ja.jtable = opt_junction_table.load() != 0 ? (const uint64_t *)(ws + L.jtable_off) : nullptr;
      const int64_t threads = n_tiles * RADIX;
      {
        ProfScope ps(PK_SEGFIX, stream);
        junction_fix_kernel<8><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(ja);
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      // Both possible continuations are launched; which one does anything is decided on the device by the flag
      // the last pass / the junction kernel raise when they meet a run that is too long for them: the full
      // segment finish repairs everything (and delivers the result into the caller's arrays), else the result
      // is copied out of the shadow if it ended there.  No host round trip in between.
      if (int rc = launch_segfix(&ctrl->flags[1])) return rc;
      if (hplan.final_sel == 1)
        if (int rc = launch_copyback(true, &ctrl->flags[1])) return rc;
      fix_flow = true;
    } else if (partial) {
      PartialCheckArgs pa{};
      pa.ss = ss; pa.n = n; pa.ko = ko; pa.ko.sub = hplan.sub; pa.ko.lshift = hplan.lshift; pa.cut = hplan.cut_digit; pa.sel = hplan.final_sel;
      pa.thresh = xo.cmp_none_thresh; pa.flag = &ctrl->flags[1];
      {
        ProfScope ps(PK_SEGFIX, stream);
        partial_check_kernel<8><<<di.sm_count * 8, 256, 0, stream>>>(pa);
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      if (int rc = launch_segfix(&ctrl->flags[1])) return rc;  // runs only if a segment longer than the threshold is out of order
      if (hplan.final_sel == 1)
        if (int rc = launch_copyback(true, &ctrl->flags[1])) return rc;
      fix_flow = true;
    } else {
      if (hybrid && !(have_plan && hplan.cut_digit == 0)) {
        if (int rc = launch_segfix(nullptr)) return rc;
        finish_ran = true;
      }
      if (int rc = launch_copyback(false, nullptr)) return rc;
    }
    stt.passes_planned = (uint32_t)kb;
    stt.hist_sweeps = hist_sweeps;  // probe (+ exact min/max) (+ exact histogram of the first pass)
    stt.algorithmic_bytes = (uint64_t)hist_sweeps * (uint64_t)n * kb + (uint64_t)kb * 2ull * (uint64_t)n * rec_bytes;
    if (have_plan) {  // the executed passes are known
      stt.passes_planned = hplan.n_exec;
      stt.algorithmic_bytes = (uint64_t)hist_sweeps * (uint64_t)n * kb + (uint64_t)hplan.n_exec * 2ull * (uint64_t)n * rec_bytes;
    }
    if (hybrid) {
      // The plan was made on the device; read it (and the fall-back flag) back.  This is the one host
      // synchronisation of the hybrid path.
      HybridCtrl hctrl{};
      if (!have_plan) CUDA_TRY(cudaMemcpyAsync(&hplan, plan, sizeof hplan, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaMemcpyAsync(&hctrl, ctrl, sizeof hctrl, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      if (fix_flow) finish_ran = hctrl.flags[1] != 0;
      stt.passes_planned = hplan.n_exec;
      stt.segfix_passes = finish_ran ? 1 : 0;
      stt.cut_digit = hplan.cut_digit;
      memcpy(&stt.segfix_moved, &hctrl.flags[2], 8);
      stt.algorithmic_bytes = (uint64_t)hist_sweeps * (uint64_t)n * kb + (uint64_t)(hplan.n_exec + stt.segfix_passes) * 2ull * (uint64_t)n * rec_bytes;
      if (hctrl.flags[0] != 0) {
        // a long bucket with distinct keys: finish with the plain digit-by-digit path (the array is a
        // permutation of the input, already ordered by its top digits)
        DevSortOpts fo;
        fo.layout_n = layout_n; fo.layout_landing = xo.layout_landing; fo.force_algo = 1;
        const int rc = sort_device(key_type, ascending, n, streams, stream, caller_workspace, workspace_bytes, fo);
        if (rc != 0) return rc;
        b200sort_stats s2 = g_last_stats;
        stt.fell_back = 1;
        stt.passes_planned += s2.passes_planned;
        stt.hist_sweeps += s2.hist_sweeps;
        stt.algorithmic_bytes += s2.algorithmic_bytes;
      }
    }
  }
  stt.kernel_launches = (uint32_t)(g_launches.load() - launches_before);
  g_last_stats = stt;
  g_have_stats = true;
  return B200SORT_OK;
}

// ------------------------------------------------------------------------------------------------
// host/device dispatch
// ------------------------------------------------------------------------------------------------
enum Side { SIDE_HOST = 0, SIDE_DEVICE = 1 };

// makes `dev` the current device for the duration of a call (device arrays are sorted on the device that owns
// them, whatever the caller's current device is)
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  int enter(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (prev != dev) {
      if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return -1; }
      switched = true;
    }
    return 0;
  }
  ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};

static int side_of(const void *p, Side *out, int *dev_out = nullptr) {
  cudaPointerAttributes at{};
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(B200SORT_ECUDA, "cudaPointerGetAttributes failed: %s (no usable CUDA device? there is no CPU fallback)",
                cudaGetErrorString(e));
  }
  *out = (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? SIDE_DEVICE : SIDE_HOST;
  if (dev_out) *dev_out = at.device;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Host arrays, SoA with payloads: pipelined staging.  PCIe is the bottleneck of a sort
// called with host arrays (16 GB each way for 1e9 u64+u64 records, 0.3 s per direction, against a 44 ms
// sort), and the two directions are independent.  So only the KEYS are waited for: they are sorted together
// with a 32-bit index while the payload arrays are still on their way in, the sorted keys start their way
// out as soon as the sort is done (full duplex with the payload upload), and every payload array is
// permuted by the index when it has arrived and follows the keys out.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iota_kernel(uint32_t *idx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) idx[i] = (uint32_t)i;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t *__restrict__ idx, const T *__restrict__ in, T *__restrict__ out,
                                                     int64_t n, uint32_t cpe) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const size_t src = (size_t)idx[i] * cpe, dst = (size_t)i * cpe;
    for (uint32_t c = 0; c < cpe; c++) out[dst + c] = in[src + c];
  }
}

static CacheEntry &stage_cache_entry(int dev) {
  static std::map<int, CacheEntry> g_stage;
  return g_stage[dev];
}
static std::vector<CacheEntry *> g_stage_entries;  // for b200sort_release_cache

static int cached_stage(int dev, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(g_mu);
  CacheEntry &c = stage_cache_entry(dev);
  if (c.bytes < bytes) {
    if (c.ptr) {
      CUDA_TRY(cudaDeviceSynchronize());
      CUDA_TRY(cudaFree(c.ptr));
      c.ptr = nullptr;
      c.bytes = 0;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(B200SORT_ENOMEM, "cudaMalloc of %zu staging bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    c.ptr = p;
    c.bytes = bytes;
    bool known = false;
    for (auto *q : g_stage_entries) known = known || q == &c;
    if (!known) g_stage_entries.push_back(&c);
  }
  *out = c.ptr;
  return 0;
}

static bool host_is_pinned(const void *p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

static int sort_host_pipelined(int key_type, bool ascending, int64_t n, const std::vector<StreamDesc> &streams, cudaStream_t caller) {
  const int kb = key_bytes_of(key_type);
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  DevInfo di;
  if (int rc = dev_info(dev, &di)) return rc;
  static thread_local cudaStream_t st_in = nullptr, st_comp = nullptr, st_out = nullptr;
  static thread_local int st_dev = -1;
  if (st_in == nullptr || st_dev != dev) {
    CUDA_TRY(cudaStreamCreateWithFlags(&st_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&st_comp, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&st_out, cudaStreamNonBlocking));
    st_dev = dev;
  }
  const size_t np = streams.size() - 1;
  // device staging: keys | index | payload s in | payload s out
  std::vector<size_t> off_in(np), off_out(np);
  size_t total = 0;
  const size_t off_keys = total; total = align_up(total + (size_t)n * kb, 256);
  const size_t off_idx = total;  total = align_up(total + (size_t)n * 4, 256);
  for (size_t s = 0; s < np; s++) {
    off_in[s] = total;  total = align_up(total + (size_t)n * streams[s + 1].elem_bytes, 256);
    off_out[s] = total; total = align_up(total + (size_t)n * streams[s + 1].elem_bytes, 256);
  }
  CacheGuard cache_guard;  // the staging buffer is shared per device as well; this call is synchronous anyway
  cache_guard.acquire(dev, caller);
  void *dbuf_v = nullptr;
  if (int rc = cached_stage(dev, total, &dbuf_v)) return rc;
  unsigned char *dbuf = (unsigned char *)dbuf_v;

  std::vector<cudaEvent_t> ev;
  auto new_event = [&]() -> cudaEvent_t {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ev.push_back(e);
    return e;
  };
  int rc = B200SORT_OK;
  auto run = [&]() -> int {
    // everything is ordered after what the caller already queued on its stream
    cudaEvent_t e_begin = new_event();
    CUDA_TRY(cudaEventRecord(e_begin, caller));
    CUDA_TRY(cudaStreamWaitEvent(st_in, e_begin, 0));
    CUDA_TRY(cudaStreamWaitEvent(st_comp, e_begin, 0));
    CUDA_TRY(cudaStreamWaitEvent(st_out, e_begin, 0));

    cudaEvent_t e_keys = new_event();
    CUDA_TRY(cudaMemcpyAsync(dbuf + off_keys, streams[0].ptr, (size_t)n * kb, cudaMemcpyHostToDevice, st_in));
    CUDA_TRY(cudaEventRecord(e_keys, st_in));
    std::vector<cudaEvent_t> e_pay(np);
    auto upload_payloads = [&]() -> int {
      for (size_t s = 0; s < np; s++) {
        CUDA_TRY(cudaMemcpyAsync(dbuf + off_in[s], streams[s + 1].ptr, (size_t)n * streams[s + 1].elem_bytes, cudaMemcpyHostToDevice, st_in));
        e_pay[s] = new_event();
        CUDA_TRY(cudaEventRecord(e_pay[s], st_in));
      }
      return 0;
    };
    // pinned host arrays: the uploads are asynchronous, queue them all before the sort (whose plan read-back
    // blocks this thread); pageable arrays: an upload blocks this thread, so launch the sort first
    bool pinned = true;
    for (size_t s = 0; s < streams.size(); s++) pinned = pinned && host_is_pinned(streams[s].ptr);
    if (pinned)
      if (int r = upload_payloads()) return r;

    CUDA_TRY(cudaStreamWaitEvent(st_comp, e_keys, 0));
    iota_kernel<<<di.sm_count * 8, 256, 0, st_comp>>>((uint32_t *)(dbuf + off_idx), n);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<StreamDesc> ds = {{dbuf + off_keys, (uint32_t)kb}, {dbuf + off_idx, 4u}};
    if (int r = sort_device(key_type, ascending, n, ds, st_comp, nullptr, 0)) return r;
    cudaEvent_t e_sorted = new_event();
    CUDA_TRY(cudaEventRecord(e_sorted, st_comp));
    if (!pinned)
      if (int r = upload_payloads()) return r;

    CUDA_TRY(cudaStreamWaitEvent(st_out, e_sorted, 0));
    CUDA_TRY(cudaMemcpyAsync(streams[0].ptr, dbuf + off_keys, (size_t)n * kb, cudaMemcpyDeviceToHost, st_out));
    for (size_t s = 0; s < np; s++) {
      const uint32_t eb = streams[s + 1].elem_bytes;
      uint32_t ck = 16;
      while (ck > 1 && (eb % ck) != 0) ck >>= 1;
      const uint32_t cpe = eb / ck;
      CUDA_TRY(cudaStreamWaitEvent(st_comp, e_pay[s], 0));
      const uint32_t *idx = (const uint32_t *)(dbuf + off_idx);
      const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)di.sm_count * 16);
      switch (ck) {
        case 16: gather_kernel<uint4><<<grid, 256, 0, st_comp>>>(idx, (const uint4 *)(dbuf + off_in[s]), (uint4 *)(dbuf + off_out[s]), n, cpe); break;
        case 8: gather_kernel<uint64_t><<<grid, 256, 0, st_comp>>>(idx, (const uint64_t *)(dbuf + off_in[s]), (uint64_t *)(dbuf + off_out[s]), n, cpe); break;
        case 4: gather_kernel<uint32_t><<<grid, 256, 0, st_comp>>>(idx, (const uint32_t *)(dbuf + off_in[s]), (uint32_t *)(dbuf + off_out[s]), n, cpe); break;
        case 2: gather_kernel<uint16_t><<<grid, 256, 0, st_comp>>>(idx, (const uint16_t *)(dbuf + off_in[s]), (uint16_t *)(dbuf + off_out[s]), n, cpe); break;
        default: gather_kernel<uint8_t><<<grid, 256, 0, st_comp>>>(idx, (const uint8_t *)(dbuf + off_in[s]), (uint8_t *)(dbuf + off_out[s]), n, cpe); break;
      }
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      cudaEvent_t e_g = new_event();
      CUDA_TRY(cudaEventRecord(e_g, st_comp));
      CUDA_TRY(cudaStreamWaitEvent(st_out, e_g, 0));
      CUDA_TRY(cudaMemcpyAsync(streams[s + 1].ptr, dbuf + off_out[s], (size_t)n * eb, cudaMemcpyDeviceToHost, st_out));
    }
    return 0;
  };
  rc = run();
  // host arrays: the call returns when the data is back
  cudaError_t e1 = cudaStreamSynchronize(st_in), e2 = cudaStreamSynchronize(st_comp), e3 = cudaStreamSynchronize(st_out);
  for (cudaEvent_t e : ev) cudaEventDestroy(e);
  if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
    rc = fail(B200SORT_ECUDA, "pipelined host sort failed on the device: %s",
              cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
  return rc;
}

static int sort_any(int key_type, bool ascending, int64_t n, const std::vector<StreamDesc> &streams, void *stream_v,
                    void *workspace, size_t workspace_bytes, const DevSortOpts &base = DevSortOpts()) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  if (n <= 1) return B200SORT_OK;  // src/radix_sort.hpp:276: nothing to do for 0 or 1 element
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B200SORT_ECUDA, "no CUDA device available (%s); libb200sort has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  Side side0;
  int dev0 = 0;
  if (int rc = side_of(streams[0].ptr, &side0, &dev0)) return rc;
  for (size_t s = 1; s < streams.size(); s++) {
    Side sd;
    int dv = 0;
    if (int rc = side_of(streams[s].ptr, &sd, &dv)) return rc;
    if (sd != side0) return fail(B200SORT_EINVAL, "arrays must be all in host memory or all in device memory");
    if (sd == SIDE_DEVICE && dv != dev0) return fail(B200SORT_EINVAL, "device arrays live on different devices (%d and %d)", dev0, dv);
  }
  if (side0 == SIDE_DEVICE) {
    DeviceScope scope;  // sort on the device that owns the arrays (include/b200sort.h)
    if (scope.enter(dev0) != 0) return fail(B200SORT_ECUDA, "cannot make device %d current", dev0);
    return sort_device(key_type, ascending, n, streams, stream, workspace, workspace_bytes, base);
  }

  // ---- host arrays, SoA with payloads, big enough to care: pipelined staging --------------------------
  if (opt_host_pipeline.load() != 0 && base.cmp_none_thresh == 0 && workspace == nullptr && streams.size() >= 2 &&
      streams[0].elem_bytes == (uint32_t)key_bytes_of(key_type) && n >= ((int64_t)1 << 20) && n < ((int64_t)1 << 32)) {
    size_t need = 0;
    for (size_t s = 0; s < streams.size(); s++) need += (size_t)n * streams[s].elem_bytes * (s ? 2 : 1);
    need += (size_t)n * 4 + 2 * ((size_t)n * (key_bytes_of(key_type) + 4));  // index + the sort's own workspace
    size_t free_b = 0, total_b = 0;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    size_t staged = 0;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      staged = stage_cache_entry(cur_dev).bytes;
    }
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need < free_b / 10 * 9 + staged)
      return sort_host_pipelined(key_type, ascending, n, streams, stream);
    cudaGetLastError();
  }
  // ---- host arrays: stage through device memory ---------------------------------------------------
  std::vector<StreamDesc> dstreams(streams.size());
  size_t total = 0;
  std::vector<size_t> offs(streams.size());
  for (size_t s = 0; s < streams.size(); s++) {
    offs[s] = total;
    total = align_up(total + (size_t)n * streams[s].elem_bytes, 256);
  }
  unsigned char *dbuf = nullptr;
  e = cudaMalloc((void **)&dbuf, total);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(B200SORT_ENOMEM, "cudaMalloc of %zu staging bytes failed: %s", total, cudaGetErrorString(e));
  }
  int rc = B200SORT_OK;
  for (size_t s = 0; s < streams.size() && rc == 0; s++) {
    dstreams[s] = {dbuf + offs[s], streams[s].elem_bytes};
    e = cudaMemcpyAsync(dstreams[s].ptr, streams[s].ptr, (size_t)n * streams[s].elem_bytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) rc = fail(B200SORT_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  }
  if (rc == 0) rc = sort_device(key_type, ascending, n, dstreams, stream, workspace, workspace_bytes, base);
  for (size_t s = 0; s < streams.size() && rc == 0; s++) {
    e = cudaMemcpyAsync(streams[s].ptr, dstreams[s].ptr, (size_t)n * streams[s].elem_bytes, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) rc = fail(B200SORT_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  }
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && rc == 0) rc = fail(B200SORT_ECUDA, "sort failed on the device: %s", cudaGetErrorString(e));
  cudaFree(dbuf);
  return rc;
}

static int check_soa(void *keys, int key_type, int64_t num, int n_payloads, void *const *payloads,
                     const uint32_t *payload_elem_bytes, std::vector<StreamDesc> *out) {
  const int kb = key_bytes_of(key_type);
  if (kb == 0) return fail(B200SORT_EINVAL, "unknown key type %d", key_type);
  if (num < 0) return fail(B200SORT_EINVAL, "negative element count");
  if (n_payloads < 0 || n_payloads > MAX_STREAMS - 1)
    return fail(B200SORT_ESHAPE, "%d payload streams (0..%d supported)", n_payloads, MAX_STREAMS - 1);
  if (num > 1 && keys == nullptr) return fail(B200SORT_EINVAL, "keys is NULL");
  out->push_back({keys, (uint32_t)kb});
  for (int p = 0; p < n_payloads; p++) {
    const uint32_t eb = payload_elem_bytes ? payload_elem_bytes[p] : 0;
    if (eb < 1 || eb > 64) return fail(B200SORT_ESHAPE, "payload %d has element size %u (1..64 supported)", p, eb);
    if (num > 1 && (payloads == nullptr || payloads[p] == nullptr)) return fail(B200SORT_EINVAL, "payload %d is NULL", p);
    out->push_back({payloads ? payloads[p] : nullptr, eb});
  }
  return 0;
}

static int check_aos(void *records, int key_type, uint32_t record_bytes, int64_t num, std::vector<StreamDesc> *out) {
  const int kb = key_bytes_of(key_type);
  if (kb == 0) return fail(B200SORT_EINVAL, "unknown key type %d", key_type);
  if (num < 0) return fail(B200SORT_EINVAL, "negative element count");
  if (record_bytes < (uint32_t)kb || record_bytes > 64 || (record_bytes & (record_bytes - 1)) != 0)
    return fail(B200SORT_ERECORD, "record size %u is not a power of two in [%d, 64]", record_bytes, kb);
  if (num > 1 && records == nullptr) return fail(B200SORT_EINVAL, "records is NULL");
  out->push_back({records, record_bytes});
  return 0;
}

}  // namespace b200sort

using namespace b200sort;

extern "C" {

int b200sort_sort_soa_ex(void *keys, int key_type, int64_t num, int ascending, int n_payloads, void *const *payloads,
                         const uint32_t *payload_elem_bytes, int64_t cmp_sort_threshold, int cmp_sorter, void *stream,
                         void *workspace, size_t workspace_bytes) {
  if (cmp_sorter != B200SORT_CMP_INSERTION && cmp_sorter != B200SORT_CMP_NONE)
    return fail(B200SORT_EINVAL, "unknown cmp_sorter %d", cmp_sorter);
  // CmpSorterInsertionSort: the full sort whatever the threshold.  CmpSorterNoSort: buckets of at most
  // cmp_sort_threshold elements may stay unordered (src/radix_sort.hpp:279, src/cmp_sorters.hpp:66-78).
  DevSortOpts base;
  if (cmp_sorter == B200SORT_CMP_NONE && cmp_sort_threshold >= CMP_NONE_MIN_THRESH) base.cmp_none_thresh = cmp_sort_threshold;
  std::vector<StreamDesc> streams;
  if (int rc = check_soa(keys, key_type, num, n_payloads, payloads, payload_elem_bytes, &streams)) return rc;
  return sort_any(key_type, ascending != 0, num, streams, stream, workspace, workspace_bytes, base);
}

int b200sort_sort_aos_ex(void *records, int key_type, uint32_t record_bytes, int64_t num, int ascending,
                         int64_t cmp_sort_threshold, int cmp_sorter, void *stream, void *workspace,
                         size_t workspace_bytes) {
  if (cmp_sorter != B200SORT_CMP_INSERTION && cmp_sorter != B200SORT_CMP_NONE)
    return fail(B200SORT_EINVAL, "unknown cmp_sorter %d", cmp_sorter);
  DevSortOpts base;
  if (cmp_sorter == B200SORT_CMP_NONE && cmp_sort_threshold >= CMP_NONE_MIN_THRESH) base.cmp_none_thresh = cmp_sort_threshold;
  std::vector<StreamDesc> streams;
  if (int rc = check_aos(records, key_type, record_bytes, num, &streams)) return rc;
  return sort_any(key_type, ascending != 0, num, streams, stream, workspace, workspace_bytes, base);
}

int b200sort_sort_soa(void *keys, int key_type, int64_t num, int ascending, int n_payloads, void *const *payloads,
                      const uint32_t *payload_elem_bytes, void *stream, void *workspace, size_t workspace_bytes) {
  return b200sort_sort_soa_ex(keys, key_type, num, ascending, n_payloads, payloads, payload_elem_bytes, 16,
                              B200SORT_CMP_INSERTION, stream, workspace, workspace_bytes);
}

int b200sort_sort_aos(void *records, int key_type, uint32_t record_bytes, int64_t num, int ascending, void *stream,
                      void *workspace, size_t workspace_bytes) {
  return b200sort_sort_aos_ex(records, key_type, record_bytes, num, ascending, 16, B200SORT_CMP_INSERTION, stream,
                              workspace, workspace_bytes);
}

size_t b200sort_workspace_bytes(int key_type, int64_t num, int n_payloads, const uint32_t *payload_elem_bytes,
                                uint32_t record_bytes) {
  const int kb = key_bytes_of(key_type);
  if (kb == 0 || num < 0) return 0;
  std::vector<StreamDesc> streams;
  if (record_bytes) {
    streams.push_back({nullptr, record_bytes});
  } else {
    streams.push_back({nullptr, (uint32_t)kb});
    for (int p = 0; p < n_payloads; p++) streams.push_back({nullptr, payload_elem_bytes[p]});
  }
  Layout L;
  // the smallest tile any geometry uses bounds the look-back size from above
  make_layout(streams, std::max<int64_t>(num, 1), HYB_MIN_TILE, &L);
  return L.total;
}

const char *b200sort_last_error(void) { return g_last_error.c_str(); }
int b200sort_version(void) { return B200SORT_VERSION; }
uint64_t b200sort_launch_count(void) { return g_launches.load(); }

static std::atomic<int64_t> *find_opt(const char *name) {
  if (!name) return nullptr;
  if (!strcmp(name, "algo")) return &opt_algo;
  if (!strcmp(name, "tile_cfg")) return &opt_tile_cfg;
  if (!strcmp(name, "first_atomic")) return &opt_first_atomic;
  if (!strcmp(name, "tma_keys")) return &opt_tma_keys;
  if (!strcmp(name, "prefetch_cols")) return &opt_prefetch_cols;
  if (!strcmp(name, "wide_tiles")) return &opt_wide_tiles;
  if (!strcmp(name, "max_chunk")) return &opt_max_chunk;
  if (!strcmp(name, "bytewise")) return &opt_bytewise;
  if (!strcmp(name, "allow_skip")) return &opt_allow_skip;
  if (!strcmp(name, "allow_reduce")) return &opt_allow_reduce;
  if (!strcmp(name, "spin_ns")) return &opt_spin_ns;
  if (!strcmp(name, "fix_in_pass")) return &opt_fix_in_pass;
  if (!strcmp(name, "hist_match")) return &opt_hist_match;
  if (!strcmp(name, "profile")) return &opt_profile;
  if (!strcmp(name, "margin_bits")) return &opt_margin_bits;
  if (!strcmp(name, "probe_guess")) return &opt_probe_guess;
  if (!strcmp(name, "probe_sample")) return &opt_probe_sample;
  if (!strcmp(name, "allow_lshift")) return &opt_allow_lshift;
  if (!strcmp(name, "mgpu_p2p")) return &opt_mgpu_p2p;
  if (!strcmp(name, "mgpu_landing")) return &opt_mgpu_landing;
  if (!strcmp(name, "mgpu_refine")) return &opt_mgpu_refine;
  if (!strcmp(name, "mgpu_overlap")) return &opt_mgpu_overlap;
  if (!strcmp(name, "mgpu_chunks")) return &opt_mgpu_chunks;
  if (!strcmp(name, "mgpu_wide")) return &opt_mgpu_wide;
  if (!strcmp(name, "mgpu_cons_smem_kb")) return &opt_mgpu_cons_smem_kb;
  if (!strcmp(name, "mgpu_persist_x2")) return &opt_mgpu_persist_x2;
  if (!strcmp(name, "mgpu_chunk_min_log2")) return &opt_mgpu_chunk_min_log2;
  if (!strcmp(name, "host_pipeline")) return &opt_host_pipeline;
  if (!strcmp(name, "junction_table")) return &opt_junction_table;
  if (!strcmp(name, "host_plan_min_log2")) return &opt_host_plan_min_log2;
  if (!strcmp(name, "nstage")) return &opt_nstage;
  return nullptr;
}
int b200sort_set_option(const char *name, int64_t value) {
  auto *o = find_opt(name);
  if (!o) return fail(B200SORT_EINVAL, "unknown option '%s'", name ? name : "(null)");
  o->store(value);
  return 0;
}
int64_t b200sort_get_option(const char *name) {
  auto *o = find_opt(name);
  return o ? o->load() : INT64_MIN;
}

int b200sort_last_stats(b200sort_stats *out) {
  if (!out || !g_have_stats) return B200SORT_EINVAL;
  *out = g_last_stats;
  return 0;
}

int b200sort_last_profile(int *kinds, float *ms, int capacity) {
  int n = 0;
  for (auto &p : g_prof) {
    if (n >= capacity) break;
    if (cudaEventSynchronize(p.e1) != cudaSuccess) return fail(B200SORT_ECUDA, "profile event failed");
    float t = 0;
    cudaEventElapsedTime(&t, p.e0, p.e1);
    kinds[n] = p.kind;
    ms[n] = t;
    n++;
  }
  return n;
}

void b200sort_release_cache(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_cache) {
    if (kv.second.ptr) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(kv.first);
      cudaDeviceSynchronize();
      cudaFree(kv.second.ptr);
      cudaSetDevice(cur);
    }
  }
  g_cache.clear();
  for (auto *c : g_stage_entries) {
    if (c->ptr) { cudaDeviceSynchronize(); cudaFree(c->ptr); c->ptr = nullptr; c->bytes = 0; }
  }
}

}  // extern "C"

#include "mgpu.cuh"
