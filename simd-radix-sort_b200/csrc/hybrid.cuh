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
constexpr int SF_FT = 2048;                      // positions a tile is responsible for
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

template <typename T>
__device__ __forceinline__ void segfix_move(const unsigned char *src, unsigned char *dst, int64_t w0, const int (&dest)[SF_PPT],
                                            uint32_t active, uint32_t cpe, uint32_t c) {
  const T *s = reinterpret_cast<const T *>(src);
  T *d = reinterpret_cast<T *>(dst);
  T v[SF_PPT];
#pragma unroll
  for (int k = 0; k < SF_PPT; k++)
    if ((active >> k) & 1) v[k] = s[(size_t)(w0 + threadIdx.x + k * SF_THREADS) * cpe + c];
  __syncthreads();  // source and destination may be the same array: read everything before writing
#pragma unroll
  for (int k = 0; k < SF_PPT; k++)
    if ((active >> k) & 1) d[(size_t)(w0 + dest[k]) * cpe + c] = v[k];
}

template <int KB, bool ANYCHUNK>
__global__ void __launch_bounds__(SF_THREADS, 2) segfix_kernel(const __grid_constant__ SegfixArgs a) {
  using O = typename OrdOf<KB>::type;
  const uint32_t cut = a.plan->cut_digit;
  if (cut == 0) return;  // every varying digit was swept: nothing to finish
  const uint32_t sel = a.plan->final_sel;

  __shared__ O wkey[SF_W];
  __shared__ int16_t sstart[SF_W];
  __shared__ int16_t send[SF_W];
  __shared__ int s_tot[SF_THREADS];

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
  for (int p = tid; p < wn; p += SF_THREADS) wkey[p] = to_ordered<KB>(load_key<KB>(ksrc, w0 + p, key_stride), a.ko);
  __syncthreads();

  // a position is a segment head when its prefix differs from its left neighbour's
  auto is_head = [&](int p) -> bool {
    if (p == 0) return w0 == 0;  // unknown when the window does not start at the array start
    return ((wkey[p] ^ wkey[p - 1]) & pmask) != 0;
  };

  // ---- segment start of every window position: running max of head positions (blocked scan) ------
  const int pb = tid * SF_PPT;
  int loc[SF_PPT];
  int run = -1;
#pragma unroll
  for (int j = 0; j < SF_PPT; j++) {
    const int p = pb + j;
    if (p < wn && is_head(p)) run = p;
    loc[j] = run;
  }
  s_tot[tid] = run;
  __syncthreads();
  {
    // exclusive prefix max over the threads before this one
    // = max over the warps before this one (warp-wide reductions) and over the lower lanes of this warp
    int ex = -1;
    for (int t = 0; t < (warp << 5); t += 32) {
      int m = s_tot[t + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      ex = max(ex, m);
    }
    int own = s_tot[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, own, o);
      if (lane >= o) own = max(own, u);
    }
    const int prev = __shfl_up_sync(0xffffffffu, own, 1);
    if (lane > 0) ex = max(ex, prev);
#pragma unroll
    for (int j = 0; j < SF_PPT; j++) {
      const int p = pb + j;
      if (p < wn) sstart[p] = (int16_t)(loc[j] >= 0 ? loc[j] : ex);
    }
  }
  __syncthreads();

  // ---- segment end (position of the next head to the right): running min from the right ------------
  run = SF_BIG;
#pragma unroll
  for (int j = SF_PPT - 1; j >= 0; j--) {
    const int p = pb + j;
    loc[j] = run;
    if (p < wn && is_head(p)) run = p;
  }
  s_tot[tid] = run;
  __syncthreads();
  {
    int ex = SF_BIG;
    for (int t = ((warp + 1) << 5); t < SF_THREADS; t += 32) {
      int m = s_tot[t + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
      ex = min(ex, m);
    }
    int own = s_tot[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_down_sync(0xffffffffu, own, o);
      if (lane + o < 32) own = min(own, u);
    }
    const int next = __shfl_down_sync(0xffffffffu, own, 1);
    if (lane < 31) ex = min(ex, next);
    const int end_of_window = (w1 == a.n) ? wn : SF_BIG;  // the array end closes the last segment
#pragma unroll
    for (int j = 0; j < SF_PPT; j++) {
      const int p = pb + j;
      if (p < wn) {
        int e = loc[j] != SF_BIG ? loc[j] : ex;
        if (e == SF_BIG) e = end_of_window;
        send[p] = (int16_t)e;
      }
    }
  }
  __syncthreads();

  // ---- classify every position this thread moves; rank inside short segments ---------------------------
  int dest[SF_PPT];
  uint32_t active = 0;
  bool fail = false;
#pragma unroll
  for (int k = 0; k < SF_PPT; k++) {
    const int p = tid + k * SF_THREADS;
    dest[k] = p;
    if (p < wn) {
      const int st = sstart[p], en = send[p];
      const int64_t gp = w0 + p;
      const bool is_long = st < 0 || en == SF_BIG || (en - st) > SF_MAXSEG;
      if (is_long) {
        // long segments stay as they are (and must consist of one repeated key); the tile that owns the
        // position copies it
        if (gp >= s0 && gp < e0) {
          active |= 1u << k;
          if (p > 0 && !is_head(p) && wkey[p] != wkey[p - 1]) fail = true;
        }
      } else {
        const int64_t gh = w0 + st;
        if (gh >= s0 && gh < e0) {  // the tile holding a segment's head orders the whole segment
          active |= 1u << k;
          if (en - st > 1) {
            const O mine = wkey[p];
            int cnt = 0;
            for (int q = st; q < en; q++) {
              const O o = wkey[q];
              cnt += (o < mine || (o == mine && q < p)) ? 1 : 0;
            }
            dest[k] = st + cnt;
          }
        }
      }
    }
  }
  if (__any_sync(0xffffffffu, fail) && lane == 0) atomicOr(&a.ctrl->flags[0], 1u);

  // ---- move every stream: window position p -> dest (side `sel` -> side 0) ----------------------------------
  for (int s = 0; s < a.ss.n_streams; s++) {
    const Stream &st = a.ss.streams[s];
    const unsigned char *src = st.buf[sel];
    unsigned char *dst = st.buf[0];
    for (uint32_t c = 0; c < st.chunks_per_elem; c++) {
      const uint32_t cb = st.chunk_bytes;
      if (cb == 8) segfix_move<uint64_t>(src, dst, w0, dest, active, st.chunks_per_elem, c);
      else if (cb == 4) segfix_move<uint32_t>(src, dst, w0, dest, active, st.chunks_per_elem, c);
      else if (cb == 16) segfix_move<uint4>(src, dst, w0, dest, active, st.chunks_per_elem, c);
      else if constexpr (ANYCHUNK) {
        if (cb == 2) segfix_move<uint16_t>(src, dst, w0, dest, active, st.chunks_per_elem, c);
        else segfix_move<uint8_t>(src, dst, w0, dest, active, st.chunks_per_elem, c);
      }
    }
  }
}

}  // namespace b200sort
