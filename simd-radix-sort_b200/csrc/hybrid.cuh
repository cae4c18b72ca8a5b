// hybrid.cuh -- the small-bucket finish of the MSB hybrid path.
//
// What it replaces in the reference (paths under /root/reference): the bottom of radixRecursion,
// src/radix_sort.hpp:276-281 -- buckets below the threshold are finished by a comparison sort
// (CmpSorterInsertionSort, src/cmp_sorters.hpp:18-37).  Here the digit sweeps stop at plan->cut_digit:
// the array is then ordered by all key bits >= 8*cut_digit, i.e. partitioned into MSB buckets
// ("segments": maximal runs of keys that agree on those bits), and this kernel orders every segment by
// the full key inside shared memory.  Segments are short by construction (scan_kernel picks the cut from
// the digit entropies) or consist of one repeated key (duplicates); a long segment with distinct keys
// raises ctrl->flags[0] and the host falls back to the full digit-by-digit path.
#pragma once
#include "kernels.cuh"
#include "../../include/b200sort.h"

namespace b200sort {

constexpr int SF_THREADS = 256;                  // 8 warps, each finishing its own chunk: no block barriers
constexpr int SF_WARPS = SF_THREADS / 32;
constexpr int SF_CHUNK = 256;                    // positions a warp is responsible for
constexpr int SF_HALO = 32;                      // keys read on either side to see whole segments
constexpr int SF_MAXSEG = SF_HALO;               // longest segment ordered in shared memory
constexpr int SF_W = SF_CHUNK + 2 * SF_HALO;     // window of a warp
constexpr int SF_PPT = SF_W / 32;                // window positions per lane
constexpr int SF_FT = SF_CHUNK * SF_WARPS;       // positions per CTA
constexpr int SF_BIG = 1 << 20;
static_assert(SF_W % 32 == 0, "window geometry");

constexpr int HYB_MIN_TILE = 2048;               // smallest tile any kernel uses (sizes the look-back array)
constexpr int64_t HYB_MIN_N = 1 << 22;           // below this the plain digit-by-digit path is used

struct HybridCtrl {
  uint32_t flags[64];   // [0] != 0: a long segment with distinct keys was met -> host falls back;
                        // [2..3]: 64-bit count of records the finish moved (statistics)
};

struct SegfixArgs {
  StreamSet ss;
  int64_t n;
  KeyOrder ko;
  const Plan *plan;
  HybridCtrl *ctrl;
  // plan values as arguments when the host has read the plan back (no dependent load at CTA start)
  uint32_t plan_in_args, arg_cut, arg_sel;
  unsigned long long arg_sub;
  uint32_t arg_lshift;
  const uint32_t *gate;   // when set: run only if *gate != 0 (the last pass / junction kernel met a run too long for them)
};

// Moves the queued elements of one chunk column (one warp): entry i goes from window position q_p[i] to
// q_d[i].  Source and destination may be the same array, so the whole column is read into the warp's
// shared-memory scratch (the key window, no longer needed) before anything is written.
template <typename T>
__device__ __forceinline__ void segfix_move(const unsigned char *src, unsigned char *dst, int64_t w0, const uint16_t *q_p,
                                            const uint16_t *q_d, int count, uint32_t cpe, uint32_t c, void *stage_raw, int lane) {
  static_assert(sizeof(T) <= 8, "columns wider than 8 bytes are moved as 8-byte halves");
  const T *s = reinterpret_cast<const T *>(src) + (size_t)w0 * cpe + c;
  T *d = reinterpret_cast<T *>(dst) + (size_t)w0 * cpe + c;
  T *stage = reinterpret_cast<T *>(stage_raw);
  for (int i = lane; i < count; i += 32) stage[i] = s[(size_t)q_p[i] * cpe];
  __syncwarp();
  for (int i = lane; i < count; i += 32) d[(size_t)q_d[i] * cpe] = stage[i];
  __syncwarp();
}

// One warp per chunk of SF_CHUNK positions, everything warp-synchronous:
//  1. the warp loads its window (chunk + SF_HALO keys on either side) as ordered keys;
//  2. a position is a segment head when its prefix (bits >= 8*cut) differs from its left neighbour's;
//     the head flags of the window are ten ballot words that every lane holds in registers;
//  3. one-key segments (the common case) need nothing; a short segment (<= SF_MAXSEG keys, hence fully
//     inside the window of every warp it touches) is ordered by rank counting by the warp that owns its
//     head; a long segment must consist of one repeated key (duplicates) -- a long segment with distinct
//     keys raises ctrl->flags[0];
//  4. the records that have to move (all owned ones when the data still sits in the shadow, only the
//     displaced ones when it already lies in the caller's arrays) are queued and moved column by column.
template <int KB, bool ANYCHUNK>
__global__ void __launch_bounds__(SF_THREADS, 4) segfix_kernel(const __grid_constant__ SegfixArgs a) {
  using O = typename OrdOf<KB>::type;
  const uint32_t cut = a.plan_in_args ? a.arg_cut : a.plan->cut_digit;
  if (cut == 0) return;  // every varying digit was swept: nothing to finish
  if (a.gate != nullptr && *reinterpret_cast<const volatile uint32_t *>(a.gate) == 0) return;  // not needed (decided on the device)
  const uint32_t sel = a.plan_in_args ? a.arg_sel : a.plan->final_sel;

  __shared__ O s_wkey[SF_WARPS][SF_W];
  __shared__ uint16_t s_qp[SF_WARPS][SF_W];
  __shared__ uint16_t s_qd[SF_WARPS][SF_W];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  O *wkey = s_wkey[warp];
  uint16_t *q_p = s_qp[warp], *q_d = s_qd[warp];
  // every warp walks over chunks with a grid-wide stride: a few long-lived CTAs instead of one CTA per
  // 2048 keys (488 K CTA launches at 1e9 records cost more than the work itself)
  const int64_t n_chunks = (a.n + SF_CHUNK - 1) / SF_CHUNK;
  for (int64_t chunk = (int64_t)blockIdx.x * SF_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * SF_WARPS) {
  const int64_t s0 = chunk * SF_CHUNK;
  const int64_t e0 = s0 + SF_CHUNK < a.n ? s0 + SF_CHUNK : a.n;
  const int64_t w0 = s0 - SF_HALO > 0 ? s0 - SF_HALO : 0;
  const int64_t w1 = e0 + SF_HALO < a.n ? e0 + SF_HALO : a.n;
  const int wn = (int)(w1 - w0);
  const int off_s = (int)(s0 - w0), off_e = (int)(e0 - w0);  // the chunk's own positions inside the window
  const O pmask = (O)(~(O)0) << (8 * cut);  // cut < KB

  const Stream &ks = a.ss.streams[0];
  const uint32_t key_stride = ks.chunk_bytes * ks.chunks_per_elem;
  const unsigned char *ksrc = ks.buf[sel];
  KeyOrder ko = a.ko;
  ko.sub = a.plan_in_args ? a.arg_sub : a.plan->sub;
  ko.lshift = a.plan_in_args ? a.arg_lshift : a.plan->lshift;

  O key[SF_PPT];
#pragma unroll
  for (int j = 0; j < SF_PPT; j++) {
    const int p = j * 32 + lane;
    key[j] = p < wn ? to_ordered<KB>(load_key<KB>(ksrc, w0 + p, key_stride), ko) : (O)0;
  }
#pragma unroll
  for (int j = 0; j < SF_PPT; j++) wkey[j * 32 + lane] = key[j];
  __syncwarp();

  // head flags: hb[j] bit l <-> window position 32 j + l; the array end closes the last segment
  uint32_t hb[SF_PPT + 2];
#pragma unroll
  for (int j = 0; j < SF_PPT; j++) {
    const int p = j * 32 + lane;
    bool h = false;
    if (p < wn) h = p == 0 ? (w0 == 0) : (((key[j] ^ wkey[p - 1]) & pmask) != 0);  // p == 0: unknown unless the array starts here
    else if (p == wn) h = w1 == a.n;
    hb[j] = __ballot_sync(0xffffffffu, h);
  }
  hb[SF_PPT] = (wn == SF_W && w1 == a.n) ? 1u : 0u;
  hb[SF_PPT + 1] = 0u;

  int q_count = 0;  // uniform across the warp
  bool fail = false;
#pragma unroll
  for (int j = 0; j < SF_PPT; j++) {
    const int p = j * 32 + lane;
    int dest = p;
    bool act = false;
    const uint32_t cur = hb[j], nxt = hb[j + 1], prv = j > 0 ? hb[j - 1] : 0u;
    const uint32_t two = __funnelshift_r(cur, nxt, lane) & 3u;  // head bits of p and p + 1
    if (p < wn && two == 3u) {
      // a head directly followed by a head: one-key segment, nothing to order
      act = sel != 0 && p >= off_s && p < off_e;
    } else if (p < wn) {
      // start: nearest head at or before p (this word, else the previous one = up to 63 positions back)
      int st = -1;
      const uint32_t below = cur & (0xffffffffu >> (31 - lane));
      if (below) st = j * 32 + 31 - __clz(below);
      else if (prv) st = (j - 1) * 32 + 31 - __clz(prv);
      // end: nearest head after p (this word, else the next one)
      int en = SF_BIG;
      const uint32_t above = cur & ~(0xffffffffu >> (31 - lane));
      if (above) en = j * 32 + __ffs(above) - 1;
      else if (nxt) en = (j + 1) * 32 + __ffs(nxt) - 1;
      const bool is_long = st < 0 || en == SF_BIG || (en - st) > SF_MAXSEG;
      if (is_long) {
        // long segments stay as they are and must consist of one repeated key
        if (p >= off_s && p < off_e) {
          act = sel != 0;
          if (p > 0 && ((cur >> lane) & 1u) == 0 && key[j] != wkey[p - 1]) fail = true;
        }
      } else if (st >= off_s && st < off_e) {  // the chunk holding a segment's head orders the whole segment
        if (en - st > 1) {
          const O mine = key[j];
          int cnt = 0;
          for (int q = st; q < en; q++) {
            const O o = wkey[q];
            cnt += (o < mine || (o == mine && q < p)) ? 1 : 0;
          }
          dest = st + cnt;
        }
        // In place only displaced records move.  (Writing whole aligned 4-record groups instead of single
        // records was measured slower, 13.1 vs 11.5 ms at 1e9: profiles/README.md.)
        act = sel != 0 || dest != p;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (act) {
      const int i = q_count + __popc(m & lanemask_lt());
      q_p[i] = (uint16_t)p;
      q_d[i] = (uint16_t)dest;
    }
    q_count += __popc(m);
  }
  if (__any_sync(0xffffffffu, fail) && lane == 0) atomicOr(&a.ctrl->flags[0], 1u);
  __syncwarp();
  if (q_count == 0) continue;
  if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&a.ctrl->flags[2]), (unsigned long long)q_count);  // statistics

  // ---- move every stream: window position q_p -> q_d (side `sel` -> side 0) -------------------------------
  for (int s = 0; s < a.ss.n_streams; s++) {
    const Stream &st = a.ss.streams[s];
    const unsigned char *src = st.buf[sel];
    unsigned char *dst = st.buf[0];
    for (uint32_t c = 0; c < st.chunks_per_elem; c++) {
      const uint32_t cb = st.chunk_bytes, cpe = st.chunks_per_elem;
      if (cb == 8) segfix_move<uint64_t>(src, dst, w0, q_p, q_d, q_count, cpe, c, wkey, lane);
      else if (cb == 4) segfix_move<uint32_t>(src, dst, w0, q_p, q_d, q_count, cpe, c, wkey, lane);
      else if (cb == 16) {
        segfix_move<uint64_t>(src, dst, w0, q_p, q_d, q_count, cpe * 2, c * 2, wkey, lane);
        segfix_move<uint64_t>(src, dst, w0, q_p, q_d, q_count, cpe * 2, c * 2 + 1, wkey, lane);
      } else if constexpr (ANYCHUNK) {
        if (cb == 2) segfix_move<uint16_t>(src, dst, w0, q_p, q_d, q_count, cpe, c, wkey, lane);
        else segfix_move<uint8_t>(src, dst, w0, q_p, q_d, q_count, cpe, c, wkey, lane);
      }
    }
  }
  __syncwarp();  // the warp's scratch is reused by its next chunk
  }  // chunk loop
}

// ------------------------------------------------------------------------------------------------
// Partial-sort mode (CmpSorterNoSort, src/cmp_sorters.hpp:66-78 with the threshold of src/radix_sort.hpp:279):
// the reference stops its recursion at buckets of at most `thresh` elements and leaves them as they are.  Here
// the digit sweeps stop at the plan's cut; this key-only sweep then checks the contract instead of finishing
// the segments: a segment (run of keys that agree on all swept bits) may stay unordered if it has at most
// `thresh` keys.  Only a segment that is LONGER than the threshold and actually out of order raises the flag
// that makes the (gated) full segment finish run.  One coalesced read of the keys; no record moves.
// ------------------------------------------------------------------------------------------------
struct PartialCheckArgs {
  StreamSet ss;
  int64_t n;
  KeyOrder ko;   // with the plan's range reduction / shift applied by the host
  uint32_t cut, sel;
  int64_t thresh;
  uint32_t *flag;
};

template <int KB>
__global__ void __launch_bounds__(256) partial_check_kernel(const __grid_constant__ PartialCheckArgs a) {
  using O = typename OrdOf<KB>::type;
  const Stream &ks = a.ss.streams[0];
  const uint32_t key_stride = ks.chunk_bytes * ks.chunks_per_elem;
  const unsigned char *kbuf = ks.buf[a.sel];
  const O pmask = (O)(~(O)0) << (8 * a.cut);
  auto okey = [&](int64_t i) -> O { return to_ordered<KB>(load_key<KB>(kbuf, i, key_stride), a.ko); };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const O k1 = okey(i), k0 = okey(i - 1);
    if (((k0 ^ k1) & pmask) != 0 || k1 >= k0) continue;  // segment boundary, or the pair is in order
    // an out-of-order pair inside a segment: fine if the segment has at most `thresh` keys
    int64_t len = 2;
    for (int64_t j = i - 2; j >= 0 && len <= a.thresh && ((okey(j) ^ k1) & pmask) == 0; j--) len++;
    for (int64_t j = i + 1; j < a.n && len <= a.thresh && ((okey(j) ^ k1) & pmask) == 0; j++) len++;
    if (len > a.thresh) atomicOr(a.flag, 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// Junction fix: companion of the FIX instantiation of onesweep_kernel.  The last pass has put every
// final segment in order as far as one tile held it; what can still be out of order are segments whose
// keys came from two (or more) consecutive tiles.  In the output such a segment lies across a "junction":
// the place inside a bucket where the contribution of tile t ends and that of tile t+1 begins -- and that
// place is exactly the inclusive prefix that tile t published in the look-back table of the last pass.
// One thread per (tile, digit) looks at the two keys around its junction; if they agree on all swept bits
// it orders the (short) segment there, all streams included.  A segment longer than JF_CAP raises the flag
// that makes the host run the full segment finish.
// ------------------------------------------------------------------------------------------------
constexpr int JF_CAP = 32;

struct JunctionArgs {
  StreamSet ss;
  int64_t n;
  KeyOrder ko;               // with the plan's range reduction already applied by the host
  const uint64_t *lookback;  // [n_tiles][RADIX] status words of the last executed pass
  int64_t n_tiles;
  uint32_t tag;              // generation tag of that pass
  uint32_t cut, sel;         // Plan::cut_digit, side holding the data
  uint32_t *flag;            // raised when a segment is too long for this kernel
  const uint64_t *jtable;    // [n_tiles][RADIX] fingerprints written by the FIX pass (SweepArgs::jtable), or nullptr
};

template <typename T>
__device__ __forceinline__ void junction_permute(unsigned char *buf, int64_t lo, int len, const uint8_t *src_of, uint32_t cpe, uint32_t c) {
  T *d = reinterpret_cast<T *>(buf) + (size_t)lo * cpe + c;
  T tmp[JF_CAP];
  for (int x = 0; x < len; x++) tmp[x] = d[(size_t)x * cpe];
  for (int x = 0; x < len; x++) d[(size_t)x * cpe] = tmp[src_of[x]];
}

template <typename T>
__device__ __forceinline__ void junction_swap(unsigned char *buf, int64_t lo, uint32_t cpe, uint32_t c) {
  T *d = reinterpret_cast<T *>(buf) + (size_t)lo * cpe + c;
  const T x = d[0], y = d[cpe];
  d[0] = y;
  d[cpe] = x;
}

template <int KB>
__global__ void __launch_bounds__(256) junction_fix_kernel(const __grid_constant__ JunctionArgs a) {
  using O = typename OrdOf<KB>::type;
  __shared__ int64_t q_g[256], q_prev[256];
  __shared__ int q_n;
  __shared__ uint32_t s_skip_all;
  if (threadIdx.x == 0) { q_n = 0; s_skip_all = *reinterpret_cast<volatile uint32_t *>(a.flag); }
  __syncthreads();
  if (s_skip_all != 0) return;  // the full finish runs anyway (one decision per CTA: there are barriers below)
  const Stream &ks = a.ss.streams[0];
  const uint32_t key_stride = ks.chunk_bytes * ks.chunks_per_elem;
  unsigned char *kbuf = ks.buf[a.sel];
  const O pmask = (O)(~(O)0) << (8 * a.cut);
  auto okey = [&](int64_t i) -> O { return to_ordered<KB>(load_key<KB>(kbuf, i, key_stride), a.ko); };

  // phase A, one thread per (tile, digit): do the two keys around the junction belong to one segment?
  // (about one junction in ten for a plan that leaves 0.25 keys per segment: collected in shared memory so
  // that the repair below runs with full warps instead of three lanes in 32)
  {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = idx / RADIX;
    const int d = (int)(idx % RADIX);
    if (t < a.n_tiles - 1) {  // the last tile has no successor
      const int64_t g = (int64_t)(a.lookback[(size_t)t * RADIX + d] & LB_VALUE_MASK);  // end of tile t's part
      const int64_t g_end = (int64_t)(a.lookback[(size_t)(a.n_tiles - 1) * RADIX + d] & LB_VALUE_MASK);  // bucket end
      const int64_t g_prev = t > 0 ? (int64_t)(a.lookback[(size_t)(t - 1) * RADIX + d] & LB_VALUE_MASK) : -1;
      // g == g_prev: tile t put nothing here, the junction coincides with an earlier one
      if (g < g_end && g > 0 && g != g_prev) {
        bool cand;
        const int64_t g_next = (int64_t)(a.lookback[(size_t)(t + 1) * RADIX + d] & LB_VALUE_MASK);
        if (a.jtable != nullptr && g_next != g) {
          // both tiles hold keys of this bucket: compare the fingerprints they left (coalesced reads; the keys
          // themselves are only fetched for the candidates, below)
          const uint32_t last_f = (uint32_t)(a.jtable[(size_t)t * RADIX + d] >> 32);
          const uint32_t first_f = (uint32_t)a.jtable[(size_t)(t + 1) * RADIX + d];
          cand = last_f == first_f;
        } else {
          cand = ((okey(g - 1) ^ okey(g)) & pmask) == 0;
        }
        if (cand) {
          const int slot = atomicAdd(&q_n, 1);
          q_g[slot] = g;
          q_prev[slot] = g_prev;
        }
      }
    }
  }
  __syncthreads();

  // phase B: order the segments found
  const int n_q = q_n;
  for (int qi = threadIdx.x; qi < n_q; qi += blockDim.x) {
    const int64_t g = q_g[qi], g_prev = q_prev[qi];
    const O k0 = okey(g - 1);
    if (((k0 ^ okey(g)) & pmask) != 0) continue;  // equal fingerprints, different segments
    int64_t lo = g - 1, hi = g + 1;
    while (lo > 0 && g - lo <= JF_CAP && ((okey(lo - 1) ^ k0) & pmask) == 0) lo--;
    while (hi < a.n && hi - g <= JF_CAP && ((okey(hi) ^ k0) & pmask) == 0) hi++;
    const int len = (int)(hi - lo);
    if (len > JF_CAP) {  // too long to merge here: fine if the two parts happen to be in order (duplicates)
      if (okey(g) < k0) atomicOr(a.flag, 1u);
      continue;
    }
    if (lo < g_prev) continue;  // an earlier junction lies inside this segment: its thread orders it
    if (len == 2) {  // by far the most common case
      if (okey(g) < k0) {
        for (int s = 0; s < a.ss.n_streams; s++) {
          const Stream &st = a.ss.streams[s];
          for (uint32_t c = 0; c < st.chunks_per_elem; c++) {
            switch (st.chunk_bytes) {
              case 1: junction_swap<uint8_t>(st.buf[a.sel], lo, st.chunks_per_elem, c); break;
              case 2: junction_swap<uint16_t>(st.buf[a.sel], lo, st.chunks_per_elem, c); break;
              case 4: junction_swap<uint32_t>(st.buf[a.sel], lo, st.chunks_per_elem, c); break;
              case 8: junction_swap<uint64_t>(st.buf[a.sel], lo, st.chunks_per_elem, c); break;
              default: junction_swap<uint4>(st.buf[a.sel], lo, st.chunks_per_elem, c); break;
            }
          }
        }
      }
      continue;
    }
    // rank counting on the full key -> source of every destination slot, then permute every stream
    O keys[JF_CAP];
    uint8_t src_of[JF_CAP];
    for (int x = 0; x < len; x++) keys[x] = okey(lo + x);
    bool moved = false;
    for (int x = 0; x < len; x++) {
      int r = 0;
      for (int y = 0; y < len; y++) r += (keys[y] < keys[x] || (keys[y] == keys[x] && y < x)) ? 1 : 0;
      src_of[r] = (uint8_t)x;
      moved = moved || r != x;
    }
    if (!moved) continue;
    for (int s = 0; s < a.ss.n_streams; s++) {
      const Stream &st = a.ss.streams[s];
      for (uint32_t c = 0; c < st.chunks_per_elem; c++) {
        switch (st.chunk_bytes) {
          case 1: junction_permute<uint8_t>(st.buf[a.sel], lo, len, src_of, st.chunks_per_elem, c); break;
          case 2: junction_permute<uint16_t>(st.buf[a.sel], lo, len, src_of, st.chunks_per_elem, c); break;
          case 4: junction_permute<uint32_t>(st.buf[a.sel], lo, len, src_of, st.chunks_per_elem, c); break;
          case 8: junction_permute<uint64_t>(st.buf[a.sel], lo, len, src_of, st.chunks_per_elem, c); break;
          default: junction_permute<uint4>(st.buf[a.sel], lo, len, src_of, st.chunks_per_elem, c); break;
        }
      }
    }
  }
}

}  // namespace b200sort
