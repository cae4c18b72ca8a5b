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

constexpr int SF_THREADS = 256;
constexpr int SF_FT = 4096;                      // positions a tile is responsible for
constexpr int SF_HALO = 256;                     // keys read on either side to see whole segments
constexpr int SF_MAXSEG = SF_HALO;               // longest segment ordered in shared memory
constexpr int SF_W = SF_FT + 2 * SF_HALO;        // window
constexpr int SF_PPT = SF_W / SF_THREADS;        // window positions per thread
constexpr int SF_BIG = 32767;
static_assert(SF_W % SF_THREADS == 0 && SF_W < SF_BIG, "window geometry");

constexpr int HYB_MIN_TILE = 2048;               // smallest tile any kernel uses (sizes the look-back array)
constexpr int64_t HYB_MIN_N = 1 << 22;           // below this the plain digit-by-digit path is used

struct HybridCtrl {
  uint32_t flags[64];   // [0] != 0: a long segment with distinct keys was met -> host falls back
};

struct SegfixArgs {
  StreamSet ss;
  int64_t n;
  KeyOrder ko;
  const Plan *plan;
  HybridCtrl *ctrl;
};

// Moves the queued elements of one chunk column: entry i goes from window position q_p[i] to q_d[i].
// Source and destination may be the same array, so everything is read before anything is written.
// Moves the queued elements of one chunk column: entry i goes from window position q_p[i] to q_d[i].
// Source and destination may be the same array, so the whole column is read into shared memory (the
// key window is no longer needed and is reused as the buffer) before anything is written.
template <typename T>
__device__ __forceinline__ void segfix_move(const unsigned char *src, unsigned char *dst, int64_t w0, const uint16_t *q_p,
                                            const uint16_t *q_d, int count, uint32_t cpe, uint32_t c, void *stage_raw) {
  static_assert(sizeof(T) <= 8, "columns wider than 8 bytes are moved as 8-byte halves");
  const T *s = reinterpret_cast<const T *>(src) + (size_t)w0 * cpe + c;
  T *d = reinterpret_cast<T *>(dst) + (size_t)w0 * cpe + c;
  T *stage = reinterpret_cast<T *>(stage_raw);
  for (int i = threadIdx.x; i < count; i += SF_THREADS) stage[i] = s[(size_t)q_p[i] * cpe];
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += SF_THREADS) d[(size_t)q_d[i] * cpe] = stage[i];
  __syncthreads();
}

template <int KB, bool ANYCHUNK>
__global__ void __launch_bounds__(SF_THREADS, 4) segfix_kernel(const __grid_constant__ SegfixArgs a) {
  using O = typename OrdOf<KB>::type;
  const uint32_t cut = a.plan->cut_digit;
  if (cut == 0) return;  // every varying digit was swept: nothing to finish
  const uint32_t sel = a.plan->final_sel;

  __shared__ O wkey[SF_W];
  __shared__ uint32_t hbits[SF_W / 32 + 2];  // one bit per window position: set = first key of a segment
  __shared__ uint16_t q_p[SF_W];              // move queue: window position ...
  __shared__ uint16_t q_d[SF_W];              // ... and where it goes
  __shared__ int q_n;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * SF_FT;
  const int64_t e0 = s0 + SF_FT < a.n ? s0 + SF_FT : a.n;
  const int64_t w0 = s0 - SF_HALO > 0 ? s0 - SF_HALO : 0;
  const int64_t w1 = e0 + SF_HALO < a.n ? e0 + SF_HALO : a.n;
  const int wn = (int)(w1 - w0);
  const O pmask = (O)(~(O)0) << (8 * cut);  // cut < KB

  const Stream &ks = a.ss.streams[0];
  const uint32_t key_stride = ks.chunk_bytes * ks.chunks_per_elem;
  const unsigned char *ksrc = ks.buf[sel];
#pragma unroll
  for (int k = 0; k < SF_PPT; k++) {
    const int p = tid + k * SF_THREADS;
    if (p < wn) wkey[p] = to_ordered<KB>(load_key<KB>(ksrc, w0 + p, key_stride), a.ko);
  }
  if (tid < 2) hbits[SF_W / 32 + tid] = 0;
  if (tid == 0) q_n = 0;
  __syncthreads();

  // a position is a segment head when its prefix differs from its left neighbour's
  auto is_head = [&](int p) -> bool {
    if (p == 0) return w0 == 0;  // unknown when the window does not start at the array start
    return ((wkey[p] ^ wkey[p - 1]) & pmask) != 0;
  };
  // head bits, one 32-bit word per warp and step; the array end closes the last segment
#pragma unroll
  for (int k = 0; k < SF_PPT; k++) {
    const int p = tid + k * SF_THREADS;
    const bool h = (p < wn && is_head(p)) || (p == wn && w1 == a.n);
    const unsigned word = __ballot_sync(0xffffffffu, h);
    if (lane == 0) hbits[warp + k * (SF_THREADS / 32)] = word;
  }
  if (tid == 0 && wn == SF_W && w1 == a.n) hbits[SF_W / 32] = 1u;
  __syncthreads();

  constexpr int MAXWORDS = SF_MAXSEG / 32;
  // start of the segment holding p: the nearest head at or before p (-1: none within SF_MAXSEG)
  auto seg_start = [&](int p) -> int {
    int w = p >> 5;
    uint32_t m = hbits[w] & (0xffffffffu >> (31 - (p & 31)));
    if (m == 0)
      for (int s = 0; m == 0 && w > 0 && s < MAXWORDS; s++) m = hbits[--w];
    return m ? (w << 5) + 31 - __clz(m) : -1;
  };
  // end of the segment holding p: the nearest head after p (SF_BIG: none within SF_MAXSEG)
  auto seg_end = [&](int p) -> int {
    int w = p >> 5;
    uint32_t m = hbits[w] & ~(0xffffffffu >> (31 - (p & 31)));  // bits above p
    if (m == 0)
      for (int s = 0; m == 0 && w < SF_W / 32 && s < MAXWORDS; s++) m = hbits[++w];
    return m ? (w << 5) + __ffs(m) - 1 : SF_BIG;
  };

  // ---- classify every window position; rank inside short segments; queue what has to move --------------
  const int off_s = (int)(s0 - w0), off_e = (int)(e0 - w0);  // this tile's own positions inside the window
  bool fail = false;
#pragma unroll 2
  for (int k = 0; k < SF_PPT; k++) {
    const int p = tid + k * SF_THREADS;
    int dest = p;
    bool act = false;
    // a head directly followed by another head is a one-key bucket: nothing to order (the common case);
    // in place it does not even have to move
    const uint32_t two = (uint32_t)((((uint64_t)hbits[(p >> 5) + 1] << 32) | hbits[p >> 5]) >> (p & 31)) & 3u;
    if (p < wn && two == 3u) {
      act = sel != 0 && p >= off_s && p < off_e;
    } else if (p < wn) {
      const int st = seg_start(p), en = seg_end(p);
      const bool is_long = st < 0 || en == SF_BIG || (en - st) > SF_MAXSEG;
      if (is_long) {
        // long segments stay as they are (and must consist of one repeated key); the tile that owns the
        // position copies it when the data still sits in the shadow
        if (p >= off_s && p < off_e) {
          act = sel != 0;
          if (p > 0 && !is_head(p) && wkey[p] != wkey[p - 1]) fail = true;
        }
      } else if (st >= off_s && st < off_e) {  // the tile holding a segment's head orders the whole segment
        if (en - st > 1) {
          const O mine = wkey[p];
          int cnt = 0;
          for (int q = st; q < en; q++) {
            const O o = wkey[q];
            cnt += (o < mine || (o == mine && q < p)) ? 1 : 0;
          }
          dest = st + cnt;
        }
        act = sel != 0 || dest != p;  // in place only displaced elements move
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&q_n, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (act) {
        const int i = base + __popc(m & lanemask_lt());
        q_p[i] = (uint16_t)p;
        q_d[i] = (uint16_t)dest;
      }
    }
  }
  if (__any_sync(0xffffffffu, fail) && lane == 0) atomicOr(&a.ctrl->flags[0], 1u);
  __syncthreads();
  const int count = q_n;
  if (count == 0) return;

  // ---- move every stream: window position q_p -> q_d (side `sel` -> side 0) -------------------------------
  for (int s = 0; s < a.ss.n_streams; s++) {
    const Stream &st = a.ss.streams[s];
    const unsigned char *src = st.buf[sel];
    unsigned char *dst = st.buf[0];
    for (uint32_t c = 0; c < st.chunks_per_elem; c++) {
      const uint32_t cb = st.chunk_bytes, cpe = st.chunks_per_elem;
      if (cb == 8) segfix_move<uint64_t>(src, dst, w0, q_p, q_d, count, cpe, c, wkey);
      else if (cb == 4) segfix_move<uint32_t>(src, dst, w0, q_p, q_d, count, cpe, c, wkey);
      else if (cb == 16) {
        segfix_move<uint64_t>(src, dst, w0, q_p, q_d, count, cpe * 2, c * 2, wkey);
        segfix_move<uint64_t>(src, dst, w0, q_p, q_d, count, cpe * 2, c * 2 + 1, wkey);
      } else if constexpr (ANYCHUNK) {
        if (cb == 2) segfix_move<uint16_t>(src, dst, w0, q_p, q_d, count, cpe, c, wkey);
        else segfix_move<uint8_t>(src, dst, w0, q_p, q_d, count, cpe, c, wkey);
      }
    }
  }
}

}  // namespace b200sort
