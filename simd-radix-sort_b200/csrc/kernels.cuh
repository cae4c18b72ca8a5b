// kernels.cuh -- sm_100a kernels of the digit pass (histogram, bucket-offset scan, one-sweep scatter).
//
// What they replace in the reference (paths under /root/reference):
//   hist_kernel      <- getSortMasks + kpopcnt          src/radix_sort.hpp:225-240, :152  (per-vector 1-bit count)
//   scan_kernel      <- writePosLeft/Right cursors       src/radix_sort.hpp:134-135,174-175 (2-bucket running prefix)
//   onesweep_kernel  <- compress_store_left_right        src/radix_sort.hpp:242-267 (the permutation of keys and
//                       + the look-back chain              every payload stream / the AoS record)
// The reference partitions by ONE key bit per pass; here a pass partitions by an 8-bit digit of the
// order-mapped key (to_ordered(), the closed form of bitDirUp, src/radix_sort.hpp:51-64).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace b200sort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_STREAMS = 64;   // key stream + 63 payload streams (src/test.cpp:124-137 uses 63)
constexpr int MAX_PASSES = 16;

// ------------------------------------------------------------------------------------------------
// key order
// ------------------------------------------------------------------------------------------------
template <int KB> struct UIntOf;
template <> struct UIntOf<1> { using type = uint8_t; };
template <> struct UIntOf<2> { using type = uint16_t; };
template <> struct UIntOf<4> { using type = uint32_t; };
template <> struct UIntOf<8> { using type = uint64_t; };
template <int KB> struct OrdOf { using type = uint32_t; };
template <> struct OrdOf<8> { using type = uint64_t; };

// Maps raw key bits to an unsigned integer whose ascending order is the order the reference produces
// (bitDirUp, src/radix_sort.hpp:51-64; table in bachelors-thesis.tex:1100-1113):
//   unsigned: u = k; signed: u = k ^ SIGN; IEEE: u = (k & SIGN) ? ~k : k ^ SIGN; descending: ~u.
// Encoded as two xor constants so that the kernels stay type-agnostic:
//   u = k ^ xor_const ^ (sign(k) ? neg_xor : 0)
struct KeyOrder {
  uint64_t xor_const;  // SIGN for signed/float, ^ MASK when descending
  uint64_t neg_xor;    // MASK ^ SIGN for float keys, else 0
  uint64_t sub;        // range reduction: the smallest mapped key of this sort (or 0), subtracted last.
                       // Keys of a narrow range that straddles the sign (e.g. -8..7, N(0,100)) differ in every
                       // digit position; after the subtraction only the low digit positions vary.
  uint32_t lshift;     // left shift applied last (0..7): when the keys agree on their leading l bits, the digit
  uint32_t pad_;       // positions of the shifted key are windows of the key that start right below them
};

template <int KB, bool SHIFT = true>
__device__ __forceinline__ typename OrdOf<KB>::type to_ordered(typename UIntOf<KB>::type raw, const KeyOrder &ko) {
  using O = typename OrdOf<KB>::type;
  O u = (O)raw;
  const O neg = (O)0 - ((u >> (8 * KB - 1)) & 1);  // all ones when the sign bit is set
  const O v = (u ^ (O)ko.xor_const ^ (neg & (O)ko.neg_xor)) - (O)ko.sub;
  return SHIFT ? (O)(v << ko.lshift) : v;  // SHIFT = false: callers that know lshift == 0 (the probe)
}

// multi-GPU: histogram bin of an ordered key inside the (sampled) global key range [lo, lo + (nb << shift)):
// keys outside the sampled range fall into the first / last bin, which keeps the mapping monotonic
__device__ __forceinline__ uint32_t range_bin(unsigned long long u, unsigned long long lo, int shift, uint32_t nb) {
  if (u <= lo) return 0u;
  const unsigned long long b = (u - lo) >> shift;
  return b >= nb ? nb - 1u : (uint32_t)b;
}

// multi-GPU partition: destination rank of a record.  The W-1 splitters are pairs (key, block): the record with
// ordered key u at local index i belongs to a rank >= r iff (u, i >> blk_shift) >= (key[r-1], blk[r-1])
// lexicographically.  blk is 0 for an ordinary splitter (every key >= key[r-1] goes right); a splitter that
// sits ON a heavy key value (more equal keys than a rank may hold: SURVEY 8e(3)) splits the equal keys by where
// they lie -- by source rank and, inside the one source the boundary falls into, by position block -- which is
// allowed because the order among equal keys is free.  The pairs are sorted, so the destination is an upper bound.
struct PartArgs {
  const unsigned long long *key;  // [n] splitter keys (ordered-key space), ascending; nullptr = not a partition pass
  const uint32_t *blk;            // [n] position-block thresholds of THIS rank (0: all equal keys go right,
                                  //     0xffffffff: none)
  int n;                          // world - 1
  int blk_shift;                  // block of local index i = i >> blk_shift
  int has_tie;                    // some blk[] is nonzero (else the index is not needed)
  // fast form (one box, 8-byte keys, no ties, splitters with zero low words = bin boundaries of full-width keys):
  // destination = number of r < 7 with (high word of u) > hi_m1[r]   (0xffffffff: never)
  uint32_t hi_only;
  uint32_t hi_m1[7];
};

__device__ __forceinline__ uint32_t part_dest(unsigned long long u, int64_t idx, const PartArgs &pa) {
  if (pa.hi_only) {
    const uint32_t h = (uint32_t)(u >> 32);
    uint32_t d = 0;
#pragma unroll
    for (int r = 0; r < 7; r++) d += h > pa.hi_m1[r] ? 1u : 0u;
    return d;
  }
  const uint32_t b = pa.has_tie ? (uint32_t)(idx >> pa.blk_shift) : 0u;
  int lo = 0, hi = pa.n;  // first splitter that is > (u, b)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long k = pa.key[mid];
    const bool ge = u > k || (u == k && b >= pa.blk[mid]);  // (u, b) >= splitter mid
    if (ge) lo = mid + 1; else hi = mid;
  }
  return (uint32_t)lo;
}

template <int KB>
__device__ __forceinline__ typename UIntOf<KB>::type load_key(const unsigned char *base, int64_t i, uint32_t stride) {
  return *reinterpret_cast<const typename UIntOf<KB>::type *>(base + (size_t)i * stride);
}

// ------------------------------------------------------------------------------------------------
// shared descriptors
// ------------------------------------------------------------------------------------------------
struct Stream {
  unsigned char *buf[2];    // [0] caller's array, [1] shadow copy in the workspace
  uint32_t chunk_bytes;     // 1,2,4,8,16: granularity the stream is moved in
  uint32_t chunks_per_elem; // element bytes / chunk_bytes
};

struct StreamSet {
  int n_streams;            // streams[0] carries the key at byte offset 0 of each element
  int pad;
  Stream streams[MAX_STREAMS];
};

// written by scan_kernel, read by every later kernel of the same sort: which passes run, and from
// which side of the ping-pong they read.
struct Plan {
  uint32_t skip[MAX_PASSES];
  uint32_t src_sel[MAX_PASSES];
  uint32_t final_sel;   // side that holds the result after the last executed pass
  uint32_t n_exec;
  // hybrid MSB path: digit positions below cut_digit are not swept; the segment-finish kernel orders
  // each run of keys that agree on all bits >= 8*cut_digit (0 = every varying digit is swept)
  uint32_t cut_digit;
  uint32_t n_const;     // digit positions on which all keys agree
  // exact digit histograms are produced just in time: first_exec_p1 - 1 is the digit position of the
  // first executed pass (counted by hist_kernel), next_exec_p1[p] - 1 the digit position of the pass
  // executed after pass p (counted by pass p itself while it has the keys in registers); 0 = none
  uint32_t first_exec_p1;
  uint32_t next_exec_p1[MAX_PASSES];
  uint32_t skewed[MAX_PASSES];  // digit position has a low-entropy histogram: aggregate equal digits per warp
  unsigned long long sub;       // range reduction (KeyOrder::sub) chosen by the plan, 0 = none
  uint32_t lshift;              // KeyOrder::lshift chosen by the plan, 0 = none
  uint32_t hist_done;           // probe_kernel's exact histogram (its guessed digit) is the first pass's: no hist_kernel
  uint32_t want_minmax;         // a low-entropy digit: the smallest/largest key would tell whether range reduction pays
};

// per-sort scalars produced by probe_kernel (zero-initialised by the host)
struct ProbeOut {
  unsigned long long or_bits;    // OR of all ordered keys
  unsigned long long nand_bits;  // OR of the complements: a bit varies iff it is set in both words
  unsigned long long max_key;    // largest ordered key
  unsigned long long nmin_key;   // complement of the smallest ordered key (so that zero-initialisation works)
  unsigned long long smax_key, snmin_key;  // the same over the sampled keys only (probe_kernel without with_minmax)
};

// look-back status word: [63:62] flag, [61:57] generation tag, [56:0] value
constexpr uint64_t LB_FLAG_AGG = 1ull << 62;
constexpr uint64_t LB_FLAG_PREFIX = 2ull << 62;
constexpr uint64_t LB_FLAG_MASK = 3ull << 62;
constexpr int LB_TAG_SHIFT = 57;
constexpr uint64_t LB_TAG_MASK = 31ull << LB_TAG_SHIFT;
constexpr uint64_t LB_VALUE_MASK = (1ull << LB_TAG_SHIFT) - 1;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// PRMT: result byte i = byte (sel nibble i & 7) of {a (0..3), b (4..7)}; a nibble with bit 3 set replicates the
// selected byte's sign bit over the result byte instead
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// peers of this lane = lanes of the (full) warp holding the same digit
template <bool USE_MATCH>
__device__ __forceinline__ unsigned digit_peers(uint32_t d) {
  if constexpr (USE_MATCH) {
    return __match_any_sync(0xffffffffu, d);
  } else {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; b++) {
      const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
      peers &= ((d >> b) & 1) ? bal : ~bal;
    }
    return peers;
  }
}

// The same for the digit in byte BYTE of a register that packs four digits.  Per bit: one LOP3 that tests the
// bit straight into a predicate, the vote, one select and one 3-input LOP3 -- written out in PTX because the
// compiler otherwise extracts every bit twice (shift, mask, compare for the vote; test, select for the mask:
// seven instructions per bit, the largest single item of the scatter pass's instruction count in round 1).
template <uint32_t MASK>
__device__ __forceinline__ void peer_step(unsigned &peers, uint32_t word) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 t, bal, m;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
      "selp.b32 m, 0xffffffff, 0, p;\n\t"
      "lop3.b32 %0, %0, bal, m, 0x90;\n\t"  // peers & ~(bal ^ m): bal where my bit is set, ~bal where it is clear
      "}"
      : "+r"(peers)
      : "r"(word), "n"(MASK));
}
// (partition pass of the multi-GPU sort on one box: at most 8 destinations, three bits tell them apart)
template <int BYTE>
__device__ __forceinline__ unsigned packed_digit_peers3(uint32_t word) {
  unsigned peers = 0xffffffffu;
  peer_step<1u << (8 * BYTE + 0)>(peers, word);
  peer_step<1u << (8 * BYTE + 1)>(peers, word);
  peer_step<1u << (8 * BYTE + 2)>(peers, word);
  return peers;
}
template <int BYTE>
__device__ __forceinline__ unsigned packed_digit_peers(uint32_t word) {
  unsigned peers = 0xffffffffu;
  peer_step<1u << (8 * BYTE + 0)>(peers, word);
  peer_step<1u << (8 * BYTE + 1)>(peers, word);
  peer_step<1u << (8 * BYTE + 2)>(peers, word);
  peer_step<1u << (8 * BYTE + 3)>(peers, word);
  peer_step<1u << (8 * BYTE + 4)>(peers, word);
  peer_step<1u << (8 * BYTE + 5)>(peers, word);
  peer_step<1u << (8 * BYTE + 6)>(peers, word);
  peer_step<1u << (8 * BYTE + 7)>(peers, word);
  return peers;
}

// ------------------------------------------------------------------------------------------------
// K1: digit histograms of every digit position in one sweep over the keys.
// Warp-aggregated shared-memory counters: lanes holding the same digit elect one lane that adds the
// group's population (__match_any_sync), so a skewed digit costs one atomic per warp, not 32.
// ------------------------------------------------------------------------------------------------
struct HistArgs {
  const unsigned char *keys;
  uint32_t stride;      // bytes between consecutive keys (element size of stream 0)
  int64_t n;
  KeyOrder ko;
  uint32_t digit_mask;  // bit p set: digit position p is counted (probe_kernel: on the sampled tiles)
  uint64_t *ghist;      // [KB][RADIX] counters, zeroed by the host
  const Plan *plan;     // hist_kernel: when set, count exactly the digit position plan->first_exec_p1 - 1
  ProbeOut *probe;      // probe_kernel
  uint32_t sample;      // probe_kernel: every sample-th tile contributes to the sampled histograms ...
  uint32_t sample_one;  // ... with one key per thread (1) or all of its keys (0)
  uint32_t with_minmax; // probe_kernel: also the smallest / largest key (else left to minmax_kernel, on demand)
  uint32_t guess_p1;    // probe_kernel: digit position + 1 whose EXACT histogram is counted on the way (0 = none)
  uint32_t guess_lshift;  // ... of the key shifted left by this many bits (the plan's expected Plan::lshift)
  uint64_t *ghist_exact;  // [KB][RADIX] exact histograms (probe_kernel: row guess_p1 - 1)
};

// One digit of one key into the block's shared-memory counters.  Lanes of a warp that all hold the same
// digit (a constant or near-constant digit position: the common skew) are aggregated into one atomic
// by a vote; USE_MATCH additionally aggregates arbitrary groups with __match_any_sync.
template <bool USE_MATCH>
__device__ __forceinline__ void hist_add(uint32_t *bins, uint32_t d, unsigned vmask) {
  if constexpr (USE_MATCH) {
    const unsigned peers = __match_any_sync(vmask, d);
    if ((peers & lanemask_lt()) == 0) atomicAdd(&bins[d], (uint32_t)__popc(peers));
  } else {
    const int leader = __ffs(vmask) - 1;
    const uint32_t d0 = __shfl_sync(vmask, d, leader);
    if (__all_sync(vmask, d == d0)) {
      if ((int)(threadIdx.x & 31) == leader) atomicAdd(&bins[d], (uint32_t)__popc(vmask));
    } else {
      atomicAdd(&bins[d], 1u);
    }
  }
}

// Key loading shared by the two key-only sweeps.  Order does not matter for a histogram, so when the keys
// are a dense, 16-byte aligned array every thread pulls NLD 16-byte vectors per tile (many bytes in
// flight per thread; with several resident CTAs per SM this is what keeps HBM busy); otherwise (AoS
// records, odd alignment) one key per load.
template <int KB, int THREADS, int NLD>
struct KeyTile {
  using O = typename OrdOf<KB>::type;
  static constexpr int VEC = 16 / KB;                 // keys per 16-byte vector
  static constexpr int PER_THREAD = NLD * VEC;
  static constexpr int TILE = THREADS * PER_THREAD;   // keys per tile
  O u[PER_THREAD];
  uint32_t valid;  // bit i: u[i] holds a key

  template <bool SHIFT = true>
  __device__ __forceinline__ void load(const unsigned char *keys, uint32_t stride, int64_t n, int64_t tile, const KeyOrder &ko) {
    using KeyT = typename UIntOf<KB>::type;
    const int64_t base = tile * TILE;
    valid = 0;
    const bool dense = stride == KB && (((uintptr_t)keys) & 15) == 0;
    if (dense && base + TILE <= n) {
      const uint4 *v = reinterpret_cast<const uint4 *>(keys + (size_t)base * KB) + threadIdx.x;
      uint4 q[NLD];
#pragma unroll
      for (int j = 0; j < NLD; j++) q[j] = v[j * THREADS];
#pragma unroll
      for (int j = 0; j < NLD; j++) {
        KeyT k[VEC];
        memcpy(k, &q[j], 16);
#pragma unroll
        for (int e = 0; e < VEC; e++) u[j * VEC + e] = to_ordered<KB, SHIFT>(k[e], ko);
      }
      valid = PER_THREAD >= 32 ? 0xffffffffu : ((1u << PER_THREAD) - 1);
    } else {
#pragma unroll
      for (int i = 0; i < PER_THREAD; i++) {
        const int64_t idx = base + (int64_t)i * THREADS + threadIdx.x;
        if (idx < n) {
          u[i] = to_ordered<KB, SHIFT>(load_key<KB>(keys, idx, stride), ko);
          valid |= 1u << i;
        } else {
          u[i] = 0;
        }
      }
    }
  }
};

// K1a: probe.  One sweep over the keys that is cheap in atomics: the exact OR / AND of all ordered keys
// (which digit positions are constant), and the histograms of every digit position on a 1/sample
// subset of the tiles (their entropies steer the pass plan; they are never used as offsets).
template <int KB, int THREADS, int NLD, bool USE_MATCH>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) probe_kernel(HistArgs a) {
  using KT = KeyTile<KB, THREADS, NLD>;
  using O = typename OrdOf<KB>::type;
  __shared__ uint32_t sh[KB * RADIX];
  __shared__ uint32_t shx[RADIX];
  for (int i = threadIdx.x; i < KB * RADIX; i += THREADS) sh[i] = 0;
  for (int i = threadIdx.x; i < RADIX; i += THREADS) shx[i] = 0;
  __syncthreads();
  const int64_t n_tiles = (a.n + KT::TILE - 1) / KT::TILE;
  const bool exact = a.guess_p1 != 0;
  const int gshift = exact ? (int)(a.guess_p1 - 1) * RADIX_BITS : 0;
  O acc_or = 0, acc_nand = 0, acc_max = 0, acc_nmin = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    KT kt;
    kt.template load<false>(a.keys, a.stride, a.n, tile, a.ko);
    if (a.with_minmax) {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++)
        if ((kt.valid >> i) & 1) {
          acc_or |= kt.u[i];
          acc_nand |= ~kt.u[i];
          acc_max = kt.u[i] > acc_max ? kt.u[i] : acc_max;
          acc_nmin = (O)~kt.u[i] > acc_nmin ? (O)~kt.u[i] : acc_nmin;
        }
    } else {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++)
        if ((kt.valid >> i) & 1) {
          acc_or |= kt.u[i];
          acc_nand |= ~kt.u[i];
        }
    }
    // the exact histogram of the digit position the first pass will most likely sweep (if the plan agrees,
    // hist_kernel does not have to read the keys a second time)
    if (exact) {
      // plain shared-memory atomics unless this warp's first row shows a crowded digit (skewed keys would
      // serialise on one counter): then equal digits are aggregated first
      const uint32_t d0 = (uint32_t)((O)(kt.u[0] << a.guess_lshift) >> gshift) & (RADIX - 1);
      const bool full = kt.valid == (KT::PER_THREAD >= 32 ? 0xffffffffu : ((1u << KT::PER_THREAD) - 1));
      const bool crowded = !__all_sync(0xffffffffu, full) ||
                           __popc(__ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0))) >= 4;
      if (crowded) {
#pragma unroll
        for (int i = 0; i < KT::PER_THREAD; i++) {
          const bool v = (kt.valid >> i) & 1;
          const unsigned vmask = __ballot_sync(0xffffffffu, v);
          if (v) hist_add<true>(shx, (uint32_t)((O)(kt.u[i] << a.guess_lshift) >> gshift) & (RADIX - 1), vmask);
        }
      } else {
#pragma unroll
        for (int i = 0; i < KT::PER_THREAD; i++) atomicAdd(&shx[(uint32_t)((O)(kt.u[i] << a.guess_lshift) >> gshift) & (RADIX - 1)], 1u);
      }
    }
    // every CTA samples every sample-th of ITS OWN tiles (tile % sample would pile all sampled tiles on
    // 1/sample of the CTAs when the grid size is a multiple of sample)
    if (((tile / gridDim.x) + blockIdx.x) % a.sample == 0) {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++) {
        if (i > 0 && a.sample_one) break;
        const bool v = (kt.valid >> i) & 1;
        const unsigned vmask = __ballot_sync(0xffffffffu, v);
        if (v) {
          if (!a.with_minmax) {  // range of the sample: tells the plan whether the exact range is worth a sweep
            acc_max = kt.u[i] > acc_max ? kt.u[i] : acc_max;
            acc_nmin = (O)~kt.u[i] > acc_nmin ? (O)~kt.u[i] : acc_nmin;
          }
#pragma unroll
          for (int p = 0; p < KB; p++)
            hist_add<USE_MATCH>(&sh[p * RADIX], (uint32_t)(kt.u[i] >> (p * RADIX_BITS)) & (RADIX - 1), vmask);
        }
      }
    }
  }
  constexpr unsigned long long KEYMASK = KB == 8 ? ~0ull : ((1ull << (8 * (KB & 7))) - 1);
  unsigned long long o = acc_or, na = acc_nand & KEYMASK, mx = acc_max, nm = acc_nmin & KEYMASK;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, s);
    na |= __shfl_xor_sync(0xffffffffu, na, s);
    const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, s), n2 = __shfl_xor_sync(0xffffffffu, nm, s);
    mx = m2 > mx ? m2 : mx;
    nm = n2 > nm ? n2 : nm;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(&a.probe->or_bits, o);
    atomicOr(&a.probe->nand_bits, na);
    if (a.with_minmax) {
      atomicMax(&a.probe->max_key, mx);
      atomicMax(&a.probe->nmin_key, nm);
    } else {
      atomicMax(&a.probe->smax_key, mx);
      atomicMax(&a.probe->snmin_key, nm);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KB * RADIX; i += THREADS) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(reinterpret_cast<unsigned long long *>(&a.ghist[i]), (unsigned long long)c);
  }
  if (exact) {
    uint64_t *out = a.ghist_exact + (size_t)(a.guess_p1 - 1) * RADIX;
    for (int i = threadIdx.x; i < RADIX; i += THREADS) {
      const uint32_t c = shx[i];
      if (c) atomicAdd(reinterpret_cast<unsigned long long *>(&out[i]), (unsigned long long)c);
    }
  }
}

// K1a': smallest and largest ordered key, only launched when the plan asks for them (Plan::want_minmax).
template <int KB, int THREADS, int NLD>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) minmax_kernel(HistArgs a) {
  using KT = KeyTile<KB, THREADS, NLD>;
  using O = typename OrdOf<KB>::type;
  const int64_t n_tiles = (a.n + KT::TILE - 1) / KT::TILE;
  O acc_max = 0, acc_nmin = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    KT kt;
    kt.template load<false>(a.keys, a.stride, a.n, tile, a.ko);
#pragma unroll
    for (int i = 0; i < KT::PER_THREAD; i++)
      if ((kt.valid >> i) & 1) {
        acc_max = kt.u[i] > acc_max ? kt.u[i] : acc_max;
        acc_nmin = (O)~kt.u[i] > acc_nmin ? (O)~kt.u[i] : acc_nmin;
      }
  }
  constexpr unsigned long long KEYMASK = KB == 8 ? ~0ull : ((1ull << (8 * (KB & 7))) - 1);
  unsigned long long mx = acc_max, nm = acc_nmin & KEYMASK;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, s), n2 = __shfl_xor_sync(0xffffffffu, nm, s);
    mx = m2 > mx ? m2 : mx;
    nm = n2 > nm ? n2 : nm;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&a.probe->max_key, mx);
    atomicMax(&a.probe->nmin_key, nm);
  }
}

// K1b: exact histogram of ONE digit position -- the first executed pass (a.plan) -- every later pass's
// histogram being counted by the pass before it.
template <int KB, int THREADS, int NLD, bool USE_MATCH>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) hist_kernel(HistArgs a) {
  using KT = KeyTile<KB, THREADS, NLD>;
  const uint32_t f = a.plan->first_exec_p1;
  if (f == 0 || a.plan->hist_done) return;  // nothing will be swept / probe_kernel counted this digit already
  const int shift = (int)(f - 1) * RADIX_BITS;
  const bool skewed = a.plan->skewed[f - 1] != 0;
  KeyOrder ko = a.ko;
  ko.sub = a.plan->sub;
  ko.lshift = a.plan->lshift;
  __shared__ uint32_t sh[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += THREADS) sh[i] = 0;
  __syncthreads();
  const int64_t n_tiles = (a.n + KT::TILE - 1) / KT::TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    KT kt;
    kt.load(a.keys, a.stride, a.n, tile, ko);
    if (skewed) {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++) {
        const bool v = (kt.valid >> i) & 1;
        const unsigned vmask = __ballot_sync(0xffffffffu, v);
        if (v) hist_add<USE_MATCH>(sh, (uint32_t)(kt.u[i] >> shift) & (RADIX - 1), vmask);
      }
    } else {
#pragma unroll
      for (int i = 0; i < KT::PER_THREAD; i++)
        if ((kt.valid >> i) & 1) atomicAdd(&sh[(uint32_t)(kt.u[i] >> shift) & (RADIX - 1)], 1u);
    }
  }
  __syncthreads();
  uint64_t *out = a.ghist + (size_t)(f - 1) * RADIX;
  for (int i = threadIdx.x; i < RADIX; i += THREADS) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(reinterpret_cast<unsigned long long *>(&out[i]), (unsigned long long)c);
  }
}

// ------------------------------------------------------------------------------------------------
// K2a: per-pass exclusive scan of the 256 bucket counts -> bucket offsets, plus the pass plan
// (a digit position on which every key agrees is skipped: the early-out the reference lacks,
// bachelors-thesis.tex:4156-4176).  One block of RADIX threads.
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
  const uint64_t *ghist;   // [n_passes][RADIX] SAMPLED histograms from probe_kernel, indexed by digit position
  const ProbeOut *probe;
  Plan *plan;
  int64_t n;
  int n_passes;            // = key bytes; pass p sweeps digit position p
  int allow_skip;
  int hybrid;              // 1: choose cut_digit (MSB hybrid)
  int allow_reduce;        // 1: range reduction allowed
  float margin_bits;       // hybrid: sweep top digits until their entropies sum to log2(n) + margin_bits
  int have_minmax;         // 1: probe->max_key / nmin_key are valid
  int allow_lshift;        // 1: the plan may shift the keys left by their common leading bits
  uint32_t start_sel;      // side of the ping-pong that holds the input (0 = the caller's arrays)
  uint32_t guess_p1;       // digit position + 1 probe_kernel counted exactly into ghist_exact (0 = none)
  uint32_t guess_lshift;   // ... of the key shifted left by this many bits
  uint64_t *ghist_exact;
};

static __global__ void __launch_bounds__(RADIX) scan_kernel(ScanArgs a) {
  __shared__ float warp_ent[RADIX / 32];
  __shared__ unsigned long long warp_tot[RADIX / 32];
  __shared__ uint32_t s_skip[MAX_PASSES];
  __shared__ float s_entropy[MAX_PASSES];
  __shared__ uint32_t s_hist_done;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned long long varying = a.probe->or_bits & a.probe->nand_bits;  // bits on which keys differ
  if (t < MAX_PASSES) s_skip[t] = (a.allow_skip && t < a.n_passes && ((varying >> (t * RADIX_BITS)) & (RADIX - 1)) == 0) ? 1 : 0;
  __syncthreads();
  for (int p = 0; p < a.n_passes; p++) {
    // entropy of digit position p on the sample: log2 m - sum(c log2 c)/m
    const unsigned long long c = a.ghist[p * RADIX + t];
    float e = c ? (float)c * log2f((float)c) : 0.f;
    unsigned long long m = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (lane == 0) { warp_ent[w] = e; warp_tot[w] = m; }
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      unsigned long long mm = 0;
      for (int i = 0; i < RADIX / 32; i++) { tot += warp_ent[i]; mm += warp_tot[i]; }
      s_entropy[p] = mm ? fmaxf(log2f((float)mm) - tot / (float)mm, 0.f) : 0.f;
    }
    __syncthreads();
  }
  if (t == 0) {
    uint32_t cut = 0, n_const = 0;
    for (int p = 0; p < a.n_passes; p++) n_const += s_skip[p];
    uint32_t lshift = 0;
    if (a.hybrid) {
      // MSB hybrid: sweep only the top digits whose (marginal) entropies add up to log2(n) + margin; the
      // runs that still agree on those digits are then short (expected length 2^-margin) or constant, and
      // the finish orders them.  A cut that saves fewer than two sweeps is not worth the finishing pass.
      const float need = log2f((float)a.n) + a.margin_bits;
      auto choose_cut = [&](const float *ent, uint32_t *skip, uint32_t *n_exec) -> uint32_t {
        uint32_t c = 0;
        float acc = 0.f;
        for (int p = a.n_passes - 1; p >= 0; p--) {
          if (!skip[p]) acc += ent[p];
          if (acc >= need) { c = (uint32_t)p; break; }
        }
        uint32_t saved = 0;
        for (uint32_t p = 0; p < c; p++) saved += skip[p] ? 0 : 1;
        if (saved < 2) c = 0;
        for (uint32_t p = 0; p < c; p++) skip[p] = 1;
        *n_exec = 0;
        for (int p = 0; p < a.n_passes; p++) *n_exec += skip[p] ? 0 : 1;
        return c;
      };
      // Variant with the keys shifted left by the l (< 8) leading bits on which they all agree: the digit
      // positions then start right below those bits, and the top position is a full 8-bit digit instead of
      // 8-l varying bits (a shard of a multi-GPU sort has its leading bits fixed by the partition).  The
      // entropies of the shifted positions are estimated from those measured on the unshifted ones, a
      // position's entropy taken as evenly spread over its varying bits.
      uint32_t skip2[MAX_PASSES];
      float ent2[MAX_PASSES];
      const int l = (a.allow_lshift && a.n_passes == 8 && varying != 0) ? (__clzll((long long)varying) & 7) : 0;
      if (l != 0) {
        const unsigned long long varying2 = varying << l;
        for (int p = 0; p < a.n_passes; p++) {
          const uint32_t vb = (uint32_t)(varying >> (p * RADIX_BITS)) & (RADIX - 1);
          const uint32_t vb_lo = p > 0 ? (uint32_t)(varying >> ((p - 1) * RADIX_BITS)) & (RADIX - 1) : 0u;
          const int n_all = __popc(vb), n_take = __popc(vb & (0xffu >> l));            // this position's low 8-l bits
          const int m_all = __popc(vb_lo), m_take = __popc(vb_lo & ~(0xffu >> l) & 0xffu);  // top l bits of the one below
          float e = 0.f;
          if (n_all) e += s_entropy[p] * (float)n_take / (float)n_all;
          if (m_all) e += s_entropy[p - 1] * (float)m_take / (float)m_all;
          ent2[p] = e;
          skip2[p] = (a.allow_skip && ((varying2 >> (p * RADIX_BITS)) & (RADIX - 1)) == 0) ? 1u : 0u;
        }
      }
      uint32_t n1 = 0, n2 = 0;
      cut = choose_cut(s_entropy, s_skip, &n1);
      if (l != 0) {
        const uint32_t cut2 = choose_cut(ent2, skip2, &n2);
        if (cut2 != 0 && n2 < n1) {
          lshift = (uint32_t)l;
          cut = cut2;
          for (int p = 0; p < a.n_passes; p++) { s_skip[p] = skip2[p]; s_entropy[p] = ent2[p]; }
        }
      }
    }
    // Range reduction: if the keys span fewer digit positions than the plan above would sweep, subtract
    // the smallest key and sweep just those low digit positions (every digit above them is then zero).
    unsigned long long sub = 0;
    uint32_t want_minmax = 0;
    {
      const unsigned long long kmask = a.n_passes >= 8 ? ~0ull : ((1ull << (8 * a.n_passes)) - 1);
      uint32_t planned = cut ? 1 : 0;  // the finish costs about a pass
      for (int p = 0; p < a.n_passes; p++) planned += s_skip[p] ? 0 : 1;
      auto digits_of_range = [&](unsigned long long mn, unsigned long long mx) -> int {
        const unsigned long long range = mx - mn;
        return range == 0 ? 0 : (63 - __clzll((long long)range)) / RADIX_BITS + 1;
      };
      if (!a.have_minmax) {
        // only the range of the sampled keys is known: if even that spans as many digit positions as the plan
        // sweeps, the true range cannot do better; else it is worth one more read of the keys (minmax_kernel)
        const unsigned long long mn = (~a.probe->snmin_key) & kmask, mx = a.probe->smax_key;
        if (a.allow_reduce && mn != 0 && mx >= mn && (uint32_t)digits_of_range(mn, mx) < planned) want_minmax = 1;
      } else {
        const unsigned long long mn = (~a.probe->nmin_key) & kmask, mx = a.probe->max_key;
        const int rb = digits_of_range(mn, mx);
        if (a.allow_reduce && mn != 0 && (uint32_t)rb < planned) {
          sub = mn;
          cut = 0;
          lshift = 0;
          for (int p = 0; p < a.n_passes; p++) {
            s_skip[p] = p >= rb ? 1 : 0;
            s_entropy[p] = 0.f;  // unknown for the shifted keys: treat every digit as skewed (safe)
          }
        }
      }
    }
    a.plan->sub = sub;
    a.plan->lshift = lshift;
    a.plan->want_minmax = want_minmax;
    uint32_t sel = a.start_sel, n_exec = 0;
    int prev = -1;
    a.plan->first_exec_p1 = 0;
    for (int p = 0; p < a.n_passes; p++) {
      a.plan->skip[p] = s_skip[p];
      a.plan->src_sel[p] = sel;
      a.plan->next_exec_p1[p] = 0;
      a.plan->skewed[p] = s_entropy[p] < 5.0f ? 1u : 0u;
      if (!s_skip[p]) {
        sel ^= 1;
        n_exec++;
        if (prev < 0) a.plan->first_exec_p1 = (uint32_t)p + 1; else a.plan->next_exec_p1[prev] = (uint32_t)p + 1;
        prev = p;
      }
    }
    a.plan->final_sel = sel;
    a.plan->n_exec = n_exec;
    a.plan->cut_digit = cut;
    a.plan->n_const = n_const;
    s_hist_done = (a.guess_p1 != 0 && a.plan->first_exec_p1 == a.guess_p1 && sub == 0 && lshift == a.guess_lshift) ? 1u : 0u;
    a.plan->hist_done = s_hist_done;
  }
  __syncthreads();
  // a wrong guess: that row of the exact histograms has to be empty again before a pass counts into it
  if (a.guess_p1 != 0 && !s_hist_done) a.ghist_exact[(size_t)(a.guess_p1 - 1) * RADIX + t] = 0;
}

// ------------------------------------------------------------------------------------------------
// K3: one-sweep scatter of one digit pass.
//  - a CTA takes the next tile (atomic ticket), ranks its keys by digit in shared memory (warp-private
//    counters + peer masks: stable), publishes the tile's per-digit counts, and resolves its global
//    bucket offsets with a decoupled look-back over the preceding tiles' status words (K2b: the
//    exclusive scan across tiles never becomes a separate pass);
//  - keys, then every payload stream (or every 16-byte column of the AoS record), are staged in shared
//    memory in bucket order and written out with consecutive threads writing consecutive addresses of
//    each bucket.
// ------------------------------------------------------------------------------------------------
struct SweepArgs {
  StreamSet ss;
  int64_t n;
  KeyOrder ko;
  int pass;                // index into plan / bin_base / tile_counter
  int shift;               // bit offset of this pass's digit in the ordered key
  const uint64_t *bin_base;  // [RADIX] bucket offsets of this pass, or nullptr: tile 0 scans ghist[pass]
  uint64_t *ghist;           // [KB][RADIX] exact digit histograms (this pass reads its own, counts the next)
  uint64_t *lookback;      // [n_tiles][RADIX]
  uint32_t *tile_counter;  // one per pass
  const Plan *plan;
  uint32_t tag;            // generation tag of this pass's status words (1..31)
  uint32_t stage_bytes;    // bytes per staged chunk (max chunk size over streams)
  // multi-GPU partition pass: when part.key != nullptr the "digit" is the destination rank of the record
  // (part_dest), not a radix digit
  PartArgs part;
  int lut_world;              // number of destinations (buckets in use)
  // partition pass over a sub-range of the tiles (the chunks of the overlapped exchange): the launch covers
  // tiles tile_first, tile_first + 1, ... (its tickets are relative to tile_first) up to record n
  uint32_t tile_first;
  uint32_t peer_wide;         // partition pass: 16-byte stores of element pairs into the destination arrays
  // partition pass with arrival flags (overlapped exchange): tiles [sig_ct[c], sig_ct[c+1]) form chunk c
  uint32_t sig_n;             // number of chunks (0: no signalling)
  uint32_t sig_ct[9];
  uint32_t *sig_done;         // [sig_n] tiles of each chunk delivered so far (zeroed)
  uint32_t *sig_flags[8];     // flag arrays of the destinations: word [source * 8 + chunk]
  uint32_t sig_me, sig_value; // this rank; the flag of chunk c is set to sig_value + c
  // ... and bucket d is written at byte offset peer_delta[d] from this GPU's own destination arrays: the
  // same array in the workspace of GPU d, mapped into this process (nullptr: everything stays local)
  const int64_t *peer_delta;
  uint32_t spin_ns;        // look-back back-off (option "spin_ns"), 0 = busy poll
  // When the host has read the plan back (large sorts) it passes what this pass needs as arguments, so a
  // CTA does not start with a dependent global load (244 K CTAs per pass at 1e9 records).
  uint32_t plan_in_args, arg_sel, arg_next_p1, arg_next_skewed;
  unsigned long long arg_sub;
  uint32_t arg_lshift;
  // FIX instantiation (last pass of the MSB hybrid plan): Plan::cut_digit and where to report a run of
  // more than FIX_CAP keys that agree on all swept bits (the full segment finish then has to run)
  // the key tile of a full tile arrives by one TMA bulk copy (SoA keys, 16-byte aligned arrays), issued by the
  // kernel's first thread right after it has drawn the tile ticket
  uint32_t tma_keys;
  // ... and the tile's part of every other column is prefetched into L2 at the same moment: the column pipeline
  // only gets to load it after the key column has left the staging buffer, and then finds it in L2
  uint32_t prefetch_cols, prefetch_bytes;       // (one column: the first one after the keys; bytes per tile)
  const unsigned char *prefetch_ptr[2];         // its array on either side of the ping-pong
  uint32_t fix_cut;
  uint32_t *fix_flag;
  // ... and [n_tiles][RADIX] words {fingerprint of the swept bits of the first key | of the last key << 32} of
  // this tile's part of every bucket, for junction_fix_kernel
  uint64_t *jtable;
};

constexpr int FIX_CAP = 8;   // longest run of keys agreeing on all swept bits that is ordered on the fly

// Per-pass constants that make the digit a handful of 32-bit operations: because the order mapping is
// an xor, digit(ordered key) = digit(raw key) ^ digit(xor_const) ^ (negative ? digit(neg_xor) : 0).
struct DigitX {
  uint32_t hi_word;  // 1: the digit lies in bits 32..63 of an 8-byte key
  uint32_t bit;      // shift inside that 32-bit word
  uint32_t xc, nx;   // digit(xor_const), digit(neg_xor)
};

// (a left shift of the ordered key by l bits moves the digit at `shift` to bit shift - l of the unshifted key;
//  it may then straddle the two 32-bit words: the funnel shift takes care of that)
__device__ __forceinline__ DigitX make_digitx(const KeyOrder &ko, int shift) {
  DigitX x;
  const int es = shift - (int)ko.lshift;  // callers use this path only when es >= 0
  x.hi_word = es >= 32;
  x.bit = es & 31;
  x.xc = (uint32_t)(ko.xor_const >> es) & (RADIX - 1);
  x.nx = (uint32_t)(ko.neg_xor >> es) & (RADIX - 1);
  return x;
}

template <int KB>
__device__ __forceinline__ uint32_t fast_digit(typename UIntOf<KB>::type raw, const DigitX &x) {
  if constexpr (KB == 8) {
    const uint32_t hi = (uint32_t)(raw >> 32), lo = (uint32_t)raw;
    const uint32_t w = x.hi_word ? (hi >> x.bit) : __funnelshift_r(lo, hi, x.bit);
    const uint32_t neg = (uint32_t)((int32_t)hi >> 31);
    return ((w & (RADIX - 1)) ^ x.xc) ^ (neg & x.nx);
  } else {
    const uint32_t w = (uint32_t)raw;
    const uint32_t neg = (uint32_t)((int32_t)(w << (32 - 8 * KB)) >> 31);
    return (((w >> x.bit) & (RADIX - 1)) ^ x.xc) ^ (neg & x.nx);
  }
}

// ---- staging of one chunk column of one stream ---------------------------------------------------------
// cp.async (LDGSTS) copies every item straight from global memory into its bucket-ordered slot of a
// staging buffer: no register round trip, and the loads stay in flight while the CTA does the look-back
// and writes the previous column out.
template <int CB>
__device__ __forceinline__ void cp_async(uint32_t smem_addr, const void *gptr) {
  if constexpr (CB == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr), "l"(gptr), "n"(CB) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous key tile into shared memory -------------
// One elected thread arms an mbarrier with the tile's byte count and issues ONE bulk copy for the whole tile
// (32 KB for 4096 8-byte keys); every thread then waits on the barrier's phase and picks its keys up from
// shared memory.  Replaces IPT global loads (+ their address arithmetic) per thread.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_tile(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
// L2 prefetch of a contiguous global range (cp.async.bulk.prefetch.L2): no shared memory, no completion to wait for
__device__ __forceinline__ void tma_prefetch_l2(const void *gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b), "r"(parity)
        : "memory");
  }
}

template <int CB, int IPT, bool FULL>
__device__ __forceinline__ void stage_async(const unsigned char *src, unsigned char *buf, const uint16_t *srank, int64_t tile_base,
                                            int idx0, int n_valid, uint32_t cpe, uint32_t c) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(buf);
  const unsigned char *s = src + ((size_t)(tile_base + idx0) * cpe + c) * CB;
  const uint16_t *sr = srank + idx0;
  if (cpe == 1) {
#pragma unroll
    for (int r = 0; r < IPT; r++)
      if (FULL || idx0 + r * 32 < n_valid) cp_async<CB>(base + (uint32_t)sr[r * 32] * CB, s + (size_t)(r * 32) * CB);
  } else {
#pragma unroll
    for (int r = 0; r < IPT; r++)
      if (FULL || idx0 + r * 32 < n_valid) cp_async<CB>(base + (uint32_t)sr[r * 32] * CB, s + (size_t)(r * 32) * cpe * CB);
  }
}

template <typename T, int IPT, bool FULL>
__device__ __forceinline__ void stage_sync(const unsigned char *src, unsigned char *buf, const uint16_t *srank, int64_t tile_base,
                                           int idx0, int n_valid, uint32_t cpe, uint32_t c) {
  T *stage = reinterpret_cast<T *>(buf);
  const T *s = reinterpret_cast<const T *>(src) + ((size_t)(tile_base + idx0) * cpe + c);
  const uint16_t *sr = srank + idx0;
#pragma unroll
  for (int r = 0; r < IPT; r++)
    if (FULL || idx0 + r * 32 < n_valid) stage[sr[r * 32]] = s[(size_t)(r * 32) * cpe];
}

// goff[k] is the destination element index of staged slot tid + k*THREADS (computed once per tile,
// shared by all streams): consecutive threads write consecutive addresses of each bucket.
template <typename T, int THREADS, int IPT, bool FULL, bool PEER = false, typename GOff = int64_t>
__device__ __forceinline__ void write_out(unsigned char *dst, const unsigned char *buf, const GOff (&goff)[IPT], int n_valid,
                                          uint32_t cpe, uint32_t c, const int64_t *pdelta = nullptr, const uint32_t *s_prefix = nullptr,
                                          const int64_t *gbase = nullptr, int n_buckets = 0, bool wide = false) {
  const T *stage = reinterpret_cast<const T *>(buf) + threadIdx.x;
  T *d = reinterpret_cast<T *>(dst) + c;
  if constexpr (PEER) {
    // Partition pass of the multi-GPU sort: a bucket is a destination GPU, its arrays are peer memory.
    // Bucket by bucket, with the threads aligned to the DESTINATION: every warp store then covers one
    // naturally aligned 32-element block of the peer's array (256 B of 8-byte elements), which is what
    // NVLink moves best; only the first and last block of a bucket's run are partial.
    for (int b = 0; b < n_buckets; b++) {
      const int s0 = (int)s_prefix[b];
      const int s1 = b + 1 < RADIX ? (int)s_prefix[b + 1] : n_valid;
      const int cnt = min(s1, n_valid) - s0;
      if (cnt <= 0) continue;
      const int64_t g0 = gbase[b] + s0;  // destination index of the bucket's first staged slot
      T *dp = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(d) + pdelta[b]);
      const T *sp = reinterpret_cast<const T *>(buf);
      if (sizeof(T) == 8 && cpe == 1 && wide && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
        // 8-byte elements: every lane moves the PAIR of elements that forms one aligned 16-byte word of the
        // destination, a warp 512 contiguous bytes of it (half the store instructions, twice the bytes per
        // NVLink write; the two staged elements are loaded separately: their shared-memory address need not be
        // 16-byte aligned).  Only the first / last element of a run can be a single.
        for (int64_t pp = (((g0 >> 1) & ~(int64_t)31) + threadIdx.x); pp * 2 < g0 + cnt; pp += THREADS) {
          const int ea = (int)(pp * 2 - g0), eb = ea + 1;
          const bool va = ea >= 0, vb = eb >= 0 && eb < cnt;  // (ea < cnt by the loop bound)
          if (va && vb) {
            ulonglong2 v;
            memcpy(&v.x, &sp[s0 + ea], 8);
            memcpy(&v.y, &sp[s0 + eb], 8);
            *reinterpret_cast<ulonglong2 *>(dp + pp * 2) = v;
          } else if (va) {
            dp[pp * 2] = sp[s0 + ea];
          } else if (vb) {
            dp[pp * 2 + 1] = sp[s0 + eb];
          }
        }
        continue;
      }
      for (int e = (int)threadIdx.x - (int)(g0 & 31); e < cnt; e += THREADS)
        if (e >= 0) dp[(size_t)(g0 + e) * cpe] = sp[s0 + e];
    }
  } else if (cpe == 1) {  // one chunk per element (the common shapes): no index multiply
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (FULL || (int)threadIdx.x + k * THREADS < n_valid) d[goff[k]] = stage[k * THREADS];
  } else {
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (FULL || (int)threadIdx.x + k * THREADS < n_valid) d[(size_t)goff[k] * cpe] = stage[k * THREADS];
  }
}

// RANK: how a key's position among the tile's keys of the same digit is found
//   RANK_BALLOT  eight ballots per row of 32 keys (peer mask from the digit's bits), warp-private counters:
//                stable
//   RANK_ATOMIC  one shared-memory atomicAdd per key on per-CTA counters, its return value is the rank: NOT
//                stable -- only for the first executed pass of a sort (its input order carries no information)
//                on a digit whose histogram is not skewed, full tiles only (-12 % on that pass)
// (An atomicOr-match ranking -- lanes OR their lane bit into a per-warp, per-digit shared-memory word and read
//  the peer mask back -- was measured 7 % SLOWER than the ballots at 1e9 records, and 32-bit destination
//  offsets 4 % slower than 64-bit ones: profiles/README.md, round 2.  Both are gone.)
constexpr int RANK_BALLOT = 0, RANK_ATOMIC = 2;

template <int KB, int THREADS, int IPT, int NSTAGE, bool ANYCHUNK, bool LUT, bool FIX, int RANK, bool BYTEWISE, bool FULL>
__device__ __forceinline__ void sweep_tile(const SweepArgs &a, unsigned char *smem, const int64_t tile, const int n_valid,
                                           const uint32_t sel, uint64_t *key_bar, const uint32_t key_parity) {
  constexpr int TILE = THREADS * IPT;
  constexpr int NWARPS = THREADS / 32;
  using KeyT = typename UIntOf<KB>::type;
  using GOff = int64_t;
  static_assert(FULL || RANK != RANK_ATOMIC, "the padded last tile is ranked by a stable method");
  static_assert(!(LUT && BYTEWISE), "the partition pass's digit is a table look-up");

  unsigned char *stage = smem;                                                               // NSTAGE * TILE * stage_bytes
  uint32_t *warp_cnt = reinterpret_cast<uint32_t *>(stage + (size_t)NSTAGE * TILE * a.stage_bytes);  // NWARPS*RADIX
  int64_t *gbase64 = reinterpret_cast<int64_t *>(warp_cnt + NWARPS * RADIX);                   // RADIX
  GOff *gbase = gbase64;
  uint32_t *s_prefix = reinterpret_cast<uint32_t *>(gbase64 + RADIX);                         // RADIX
  uint32_t *s_wsum = s_prefix + RADIX;                                                        // 32
  uint16_t *srank = reinterpret_cast<uint16_t *>(s_wsum + 32);                               // TILE
  uint8_t *sdigit = reinterpret_cast<uint8_t *>(srank + TILE);                               // TILE
  uint32_t *nhist = s_prefix;  // RADIX, zeroed by the kernel: the next digit's counts live here until thread d has
                               // read nhist[d], just before it writes s_prefix[d]
  int8_t *sdelta = reinterpret_cast<int8_t *>(sdigit + TILE);                                // TILE (FIX only): slot displacement
  int64_t *pdelta = reinterpret_cast<int64_t *>(sdigit + TILE);                              // RADIX (LUT only): peer byte offsets

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile_base = tile * TILE;
  const int idx0 = warp * (IPT * 32) + lane;  // this thread's items are idx0 + r*32

  const Stream &ks = a.ss.streams[0];
  const uint32_t key_stride = ks.chunk_bytes * ks.chunks_per_elem;
  const unsigned char *kp = ks.buf[sel] + (size_t)(tile_base + idx0) * key_stride;

  // ---- load keys ------------------------------------------------------------------------------------
  KeyT raw[IPT];
  if (FULL && a.tma_keys) {
    // the tile was fetched by the bulk copy issued at kernel entry: wait for its bytes, then read the keys from
    // the staging buffer (consecutive lanes, consecutive keys: conflict-free).  The buffer is free again after
    // the barrier that follows the ranking, long before the bucket-ordered keys are staged into it.
    mbar_wait(key_bar, key_parity);
    const KeyT *kt = reinterpret_cast<const KeyT *>(stage) + idx0;
#pragma unroll
    for (int r = 0; r < IPT; r++) raw[r] = kt[r * 32];
  } else {
#pragma unroll
    for (int r = 0; r < IPT; r++)
      raw[r] = (FULL || idx0 + r * 32 < n_valid) ? *reinterpret_cast<const KeyT *>(kp + (size_t)(r * 32) * key_stride) : (KeyT)0;
  }
  // Digits are computed once and kept packed four to a register (IPT/4 registers).
  KeyOrder ko = a.ko;
  ko.sub = LUT ? 0ull : (a.plan_in_args ? a.arg_sub : a.plan->sub);
  ko.lshift = LUT ? 0u : (a.plan_in_args ? a.arg_lshift : a.plan->lshift);
  // range reduction active (or a shift larger than this digit's offset, which the plan never produces): the
  // xor shortcut of fast_digit does not apply
  // BYTEWISE (the host knows the plan: no range reduction, no left shift): this pass's digit and the next
  // pass's are whole bytes of the raw key xor-ed with constants, and the code for everything else is not
  // even compiled in (fewer registers, no spills)
  const bool has_sub = !BYTEWISE && (ko.sub != 0 || (int)ko.lshift > a.shift);
  const int es = a.shift - (BYTEWISE ? 0 : (int)ko.lshift);  // bit offset of this pass's digit in the (mapped, unshifted) key
  // A slot's digit is needed once more when the staged tile is written out (which bucket a slot belongs to).  In
  // the common case it is re-derived there from the staged key itself (one byte permute); otherwise the
  // ranking leaves it in a byte array.
  // (the partition pass's "digit" is a table look-up, that of range-reduced or left-shifted keys costs the full
  //  order mapping: those are kept in a byte array by the ranking)
  constexpr bool use_sdigit = !BYTEWISE;
  static_assert(IPT % 4 == 0, "digits are packed four per register");
  uint32_t dpack[IPT / 4];
  // 32-bit word of the raw key that holds the (byte-aligned) digit at bit offset es of the key
  auto word_at = [&](int r, int es) -> uint32_t {
    if constexpr (KB == 8) return (uint32_t)(raw[r] >> (es & 32));  // (a shift, not a select: no predicate)
    else return (uint32_t)raw[r];
  };
  auto top_word = [&](int r) -> uint32_t {  // the word whose byte TOPB carries the sign bit
    if constexpr (KB == 8) return (uint32_t)(raw[r] >> 32);
    else return (uint32_t)raw[r];
  };
  constexpr uint32_t TOPB = KB == 8 ? 3u : (uint32_t)(KB - 1);
  {
    const DigitX dx = make_digitx(ko, a.shift);
    // padding of the last tile ranks behind everything (digit 255, last in index order)
    auto pad = [&](int r, uint32_t d) -> uint32_t { return (!FULL && idx0 + r * 32 >= n_valid) ? (uint32_t)(RADIX - 1) : d; };
    if constexpr (LUT) {
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; e++)
          w |= pad(4 * q + e, part_dest((unsigned long long)to_ordered<KB>(raw[4 * q + e], ko), tile_base + idx0 + (4 * q + e) * 32, a.part)) << (8 * e);  // partition pass
        dpack[q] = w;
      }
    } else if (!BYTEWISE && has_sub) {
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; e++)
          w |= pad(4 * q + e, (uint32_t)(to_ordered<KB>(raw[4 * q + e], ko) >> a.shift) & (RADIX - 1)) << (8 * e);
        dpack[q] = w;
      }
    } else if (BYTEWISE || (es & 7) == 0) {
      // The common case (no left shift): the digit is one byte of the raw key xor-ed with a constant, so four
      // digits are gathered into a register by three byte permutes and one xor (+ the sign handling of IEEE keys).
      const uint32_t bsel = (uint32_t)(es >> 3) & 3u;
      const uint32_t s2 = bsel | ((4u + bsel) << 4) | 0x4400u;
      const uint32_t xc4 = dx.xc * 0x01010101u, nx4 = dx.nx * 0x01010101u;
      constexpr uint32_t SGN2 = (8u + TOPB) | ((12u + TOPB) << 4) | 0x4400u;  // sign bytes of two keys
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        const uint32_t t01 = prmt(word_at(4 * q, es), word_at(4 * q + 1, es), s2);
        const uint32_t t23 = prmt(word_at(4 * q + 2, es), word_at(4 * q + 3, es), s2);
        uint32_t w = prmt(t01, t23, 0x5410u) ^ xc4;
        if (nx4 != 0) {  // IEEE keys: negative keys have their other bits flipped as well
          const uint32_t m01 = prmt(top_word(4 * q), top_word(4 * q + 1), SGN2);
          const uint32_t m23 = prmt(top_word(4 * q + 2), top_word(4 * q + 3), SGN2);
          w ^= prmt(m01, m23, 0x5410u) & nx4;
        }
        if constexpr (!FULL) {
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (idx0 + (4 * q + e) * 32 >= n_valid) w |= 0xffu << (8 * e);
        }
        dpack[q] = w;
      }
    } else if constexpr (!BYTEWISE) {
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) w |= pad(4 * q + e, fast_digit<KB>(raw[4 * q + e], dx)) << (8 * e);
        dpack[q] = w;
      }
    }
  }
  auto digit_of = [&](int r) -> uint32_t { return (dpack[r >> 2] >> (8 * (r & 3))) & (RADIX - 1); };

  // ---- rank inside the warp ------------------------------------------------------------------------------
  // ranks (< TILE <= 65536) are kept two to a register: the kernel is at its register limit for 3 CTAs/SM
  uint32_t rankp[IPT / 2];
  auto rank_get = [&](int r) -> uint32_t { return (r & 1) ? (rankp[r >> 1] >> 16) : (rankp[r >> 1] & 0xffffu); };
  auto rank_set = [&](int r, uint32_t v) { rankp[r >> 1] = (r & 1) ? ((rankp[r >> 1] & 0xffffu) | (v << 16)) : ((rankp[r >> 1] & 0xffff0000u) | v); };
#pragma unroll
  for (int q = 0; q < IPT / 2; q++) rankp[q] = 0;
  uint32_t *wc = RANK == RANK_ATOMIC ? warp_cnt : warp_cnt + warp * RADIX;
  // the pass executed after this one gets its exact digit histogram from here (keys are in registers)
  const uint32_t next_p1 = LUT ? 0u : (a.plan_in_args ? a.arg_next_p1 : a.plan->next_exec_p1[a.pass]);
  if (next_p1 != 0) {
    const int nshift = (int)(next_p1 - 1) * RADIX_BITS;
    const DigitX dn = make_digitx(ko, nshift);
    const bool skewed = (a.plan_in_args ? a.arg_next_skewed : a.plan->skewed[next_p1 - 1]) != 0;
    const int esn = nshift - (BYTEWISE ? 0 : (int)ko.lshift);
    if (BYTEWISE && dn.nx == 0 && !skewed) {
      // byte-aligned digit of an integer key: one byte permute (+ the constant) per key
      const uint32_t s1 = ((uint32_t)(esn >> 3) & 3u) | 0x4440u;
#pragma unroll
      for (int r = 0; r < IPT; r++)
        if (FULL || idx0 + r * 32 < n_valid) atomicAdd(&nhist[prmt(word_at(r, esn), 0u, s1) ^ dn.xc], 1u);
    } else if (has_sub || skewed) {
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const bool valid = FULL || idx0 + r * 32 < n_valid;
        const unsigned vmask = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
        const uint32_t dnext = has_sub ? (uint32_t)(to_ordered<KB>(raw[r], ko) >> nshift) & (RADIX - 1) : fast_digit<KB>(raw[r], dn);
        if (valid) hist_add<false>(nhist, dnext, vmask);
      }
    } else {
#pragma unroll
      for (int r = 0; r < IPT; r++)
        if (FULL || idx0 + r * 32 < n_valid) atomicAdd(&nhist[fast_digit<KB>(raw[r], dn)], 1u);
    }
  }
  if constexpr (RANK == RANK_ATOMIC) {
    // unstable: the old value of the CTA's counter is the key's rank among the tile's keys of that digit
#pragma unroll
    for (int r = 0; r < IPT; r++) rank_set(r, atomicAdd(&wc[digit_of(r)], 1u));
  } else {
    // The lowest lane of every group of equal digits adds the group's size to the warp's counter of that digit;
    // the counter's old value is the group's first rank, handed to the other lanes by a shuffle.  (One returning
    // shared-memory atomic by the leaders instead of a load by everyone plus a store by the leaders: the scatter
    // pass is bound by L1TEX/shared-memory wavefronts, and rows are issued in order, so the ranks stay stable.)
    const unsigned lt = lanemask_lt();
    // (Ranking two or four rows at a time -- all ballots, then the leader atomics back to back, then the shuffles --
    //  would overlap the atomic -> shuffle round trips, but exhausts the seven predicate registers: ptxas gives up.)
    auto rank_row = [&](int r, unsigned peers) {
      const uint32_t d = digit_of(r);
      const uint32_t lower = __popc(peers & lt);
      uint32_t first = 0;
      if (lower == 0) first = atomicAdd(&wc[d], (uint32_t)__popc(peers));
      first = __shfl_sync(0xffffffffu, first, __ffs(peers) - 1);
      rank_set(r, first + lower);
    };
    // (LUT, full tile, at most 8 destinations: the digits are < 8; a padded last tile has digit 255 in it)
    if (LUT && FULL && a.lut_world <= 8) {
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        rank_row(4 * q + 0, packed_digit_peers3<0>(dpack[q]));
        rank_row(4 * q + 1, packed_digit_peers3<1>(dpack[q]));
        rank_row(4 * q + 2, packed_digit_peers3<2>(dpack[q]));
        rank_row(4 * q + 3, packed_digit_peers3<3>(dpack[q]));
      }
    } else {
#pragma unroll
      for (int q = 0; q < IPT / 4; q++) {
        rank_row(4 * q + 0, packed_digit_peers<0>(dpack[q]));
        rank_row(4 * q + 1, packed_digit_peers<1>(dpack[q]));
        rank_row(4 * q + 2, packed_digit_peers<2>(dpack[q]));
        rank_row(4 * q + 3, packed_digit_peers<3>(dpack[q]));
      }
    }
  }
  __syncthreads();

  // ---- per-digit totals: exclusive scan across warps, then across digits ----------------------------
  uint32_t my_count = 0;
  if (tid < RADIX) {
    if (next_p1 != 0) {
      const uint32_t c = nhist[tid];
      if (c) atomicAdd(reinterpret_cast<unsigned long long *>(&a.ghist[(size_t)(next_p1 - 1) * RADIX + tid]), (unsigned long long)c);
    }
    uint32_t run = 0;
    if constexpr (RANK == RANK_ATOMIC) {
      run = warp_cnt[tid];
    } else {
#pragma unroll
      for (int w = 0; w < NWARPS; w++) run += warp_cnt[w * RADIX + tid];
    }
    my_count = run;
    // the tile's count of digit tid goes out as early as possible: successors are waiting for it
    // (counts published to other tiles exclude the padding of the last tile; tile 0 publishes its prefix below)
    if (tile != (int64_t)a.tile_first) {
      const uint64_t vc = run - ((!FULL && tid == RADIX - 1) ? (uint32_t)(TILE - n_valid) : 0u);
      st_relaxed_u64(&a.lookback[(size_t)tile * RADIX + tid], LB_FLAG_AGG | ((uint64_t)a.tag << LB_TAG_SHIFT) | vc);
    }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_wsum[warp] = inc;
    s_prefix[tid] = inc - run;  // exclusive within the warp's 32 digits; warp offset added below
  }
  // tile 0 turns the pass's exact histogram into bucket offsets (exclusive scan over the 256 counts)
  uint64_t tile0_base = 0, t0_inc = 0, t0_cnt = 0;
  const bool first_tile = tile == (int64_t)a.tile_first;  // the launch's first tile seeds the chain with the bucket offsets
  if (first_tile && tid < RADIX) {
    if (a.bin_base != nullptr) {
      tile0_base = a.bin_base[tid];
    } else {
      t0_cnt = a.ghist[(size_t)a.pass * RADIX + tid];
      t0_inc = t0_cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, t0_inc, o);
        if (lane >= o) t0_inc += v;
      }
      if (lane == 31) gbase64[warp] = (int64_t)t0_inc;  // gbase is free until the look-back
    }
  }
  __syncthreads();
  if (first_tile && tid < RADIX && a.bin_base == nullptr) {
    uint64_t off = 0;
    for (int w = 0; w < warp; w++) off += (uint64_t)gbase64[w];
    tile0_base = off + t0_inc - t0_cnt;
  }
  const uint64_t tagbits = (uint64_t)a.tag << LB_TAG_SHIFT;
  uint64_t valid_count = 0;
  if (tid < RADIX) {
    uint32_t off = 0;
    for (int w = 0; w < warp; w++) off += s_wsum[w];
    const uint32_t first_slot = s_prefix[tid] + off;  // staging slot of the tile's first key of digit tid
    s_prefix[tid] = first_slot;
    // The counters turn into staging offsets: first slot of (this warp's) keys of the digit, so that a key's
    // final slot is one shared-memory load away from its rank inside the warp (conflict-free here: consecutive
    // threads, consecutive words; one random load less per key there).
    if constexpr (RANK == RANK_ATOMIC) {
      warp_cnt[tid] = first_slot;
    } else {
      uint32_t acc = first_slot;
#pragma unroll
      for (int w = 0; w < NWARPS; w++) {
        const uint32_t t = warp_cnt[w * RADIX + tid];
        warp_cnt[w * RADIX + tid] = acc;
        acc += t;
      }
    }
    // counts published to other tiles exclude the padding of the last tile
    valid_count = my_count - ((!FULL && tid == RADIX - 1) ? (uint32_t)(TILE - n_valid) : 0u);
    if constexpr (LUT) pdelta[tid] = a.peer_delta ? a.peer_delta[tid] : 0;
    if (first_tile) st_relaxed_u64(&a.lookback[(size_t)tile * RADIX + tid], LB_FLAG_PREFIX | tagbits | (tile0_base + valid_count));
  }
  __syncthreads();  // s_prefix complete

  // final rank inside the tile = digit offset + offset of this warp inside the digit + rank inside warp
#pragma unroll
  for (int r = 0; r < IPT; r++) {
    const uint32_t d = digit_of(r);
    const uint32_t rk = rank_get(r) + wc[d];
    rank_set(r, rk);
    if constexpr (use_sdigit) sdigit[rk] = (uint8_t)d;
    srank[idx0 + r * 32] = (uint16_t)rk;  // payload streams pick their slot up from here
  }

  // ---- the column pipeline.  Column 0 is the key array itself (SoA: staged from registers) or the first
  //      column of the record (AoS); columns are staged NSTAGE deep so that the loads of the next column
  //      overlap the look-back and the write-out of the current one. ------------------------------------------
  const bool soa_keys = key_stride == KB;
  unsigned char *bufs[2] = {stage, stage + (NSTAGE == 2 ? (size_t)TILE * a.stage_bytes : 0)};
  int is_s = soa_keys ? 1 : 0;  // issue cursor: next column to stage = chunk is_c of stream is_s
  uint32_t is_c = 0;
  auto issue_next = [&](unsigned char *buf) -> bool {
    if (is_s >= a.ss.n_streams) return false;
    const Stream &st = a.ss.streams[is_s];
    const unsigned char *src = st.buf[sel];
    const uint32_t cb = st.chunk_bytes, cpe = st.chunks_per_elem;
    if (cb == 8) stage_async<8, IPT, FULL>(src, buf, srank, tile_base, idx0, n_valid, cpe, is_c);
    else if (cb == 4) stage_async<4, IPT, FULL>(src, buf, srank, tile_base, idx0, n_valid, cpe, is_c);
    else if (cb == 16) stage_async<16, IPT, FULL>(src, buf, srank, tile_base, idx0, n_valid, cpe, is_c);
    else if constexpr (ANYCHUNK) {
      if (cb == 2) stage_sync<uint16_t, IPT, FULL>(src, buf, srank, tile_base, idx0, n_valid, cpe, is_c);
      else stage_sync<uint8_t, IPT, FULL>(src, buf, srank, tile_base, idx0, n_valid, cpe, is_c);
    }
    cp_async_commit();
    if (++is_c == cpe) { is_c = 0; is_s++; }
    return true;
  };
  if (soa_keys) {
    KeyT *kst = reinterpret_cast<KeyT *>(bufs[0]);
#pragma unroll
    for (int r = 0; r < IPT; r++)
      if (FULL || idx0 + r * 32 < n_valid) {
        kst[rank_get(r)] = raw[r];
      }
  } else {
    issue_next(bufs[0]);
  }
  bool have_next = false;               // column j+1 already issued into the other buffer
  if (NSTAGE == 2) have_next = issue_next(bufs[1]);

  // FIX: the tile-local ordering only needs the staged keys, so it runs BEFORE the look-back -- the
  // predecessors get that much more time to publish their prefixes and the look-back finds them ready.
  __shared__ uint32_t s_given_up;
  if constexpr (FIX) {
    cp_async_wait<0>();  // column 0 has landed (mine)
    if (tid == 0) s_given_up = ld_relaxed_u32(a.fix_flag);  // one decision for the whole tile (see below)
    __syncthreads();     // column 0 visible to everyone
  }
  if constexpr (FIX) {
    // Last pass of the MSB hybrid plan.  In bucket order, the tile's keys that agree on ALL swept bits
    // (>= 8*fix_cut) are adjacent: within a bucket they are still ordered by the lower swept digits.  Such a
    // run is one final segment as far as this tile holds it, so it is put in full-key order here, while the
    // keys sit in shared memory.  Nothing is moved in shared memory: every member of a run computes where it
    // belongs inside the run and leaves the displacement in sdelta[slot]; the destination offsets of ALL
    // columns (goff, below) pick it up.  Only segments that straddle two tiles are left for
    // junction_fix_kernel.
    {
      using O = typename OrdOf<KB>::type;
      // column 0 holds the keys: the key array itself, or the leading chunk of every record (key at offset 0)
      const KeyT *kcol = reinterpret_cast<const KeyT *>(bufs[0]);
      const int kstep = soa_keys ? 1 : (int)(ks.chunk_bytes / KB);
      auto key_at = [&](int i) -> KeyT { return kcol[i * kstep]; };
      // swept bits of the ordered (and possibly left-shifted) key, as a mask on the raw key
      const int cut_bit = 8 * (int)a.fix_cut - (int)ko.lshift;
      const O pmask = (O)(~(O)0) << cut_bit;
      const int base = tid * IPT;
      const int lim = FULL ? TILE : n_valid;
#pragma unroll
      for (int q = 0; q < IPT / 16; q++) reinterpret_cast<uint4 *>(sdelta + base)[q] = make_uint4(0, 0, 0, 0);
      // The thread owns slots base .. base+IPT-1 and looks at the window base-1 .. base+IPT (W = IPT+2 slots),
      // lane l starting at window position l mod W: a warp's loads then spread over the banks although the
      // threads' windows are IPT keys apart.  (Equal swept bits of the ordered keys <=> equal swept bits of
      // the raw keys: the order mapping is an xor whose only key-dependent part is the sign bit, itself a
      // swept bit, and no range reduction is active when a cut exists.  Slots outside the tile are masked below.)
      // once any tile has met a run that is too long, the full segment finish is going to run anyway: the
      // tiles after it do not bother (keys with many duplicates would otherwise pay for nothing).  The whole
      // tile takes the same decision: the members of a run must all apply their displacement, or none.
      if (s_given_up == 0) {
        constexpr int W = IPT + 2;
        const int rot = lane % W;
        uint32_t eq = 0;  // bit k: window positions (k+rot)%W and (k+1+rot)%W hold equal swept bits
        auto scan_window = [&](auto tag) {
          using T = decltype(tag);
          T pref[W];
  #pragma unroll
          for (int k = 0; k < W; k++) {
            int w = k + rot;
            if (w >= W) w -= W;
            const int i = base - 1 + w;
            T v = (T)1;
            if (i >= 0 && i < lim) {
              if constexpr (sizeof(T) == sizeof(O)) v = (T)((O)key_at(i) & pmask);
              else v = reinterpret_cast<const uint32_t *>(kcol)[2 * i * kstep + 1] & (uint32_t)((uint64_t)pmask >> 32);
            }
            pref[k] = v;
          }
  #pragma unroll
          for (int k = 0; k < W; k++) eq |= (uint32_t)(pref[k] == pref[(k + 1) % W]) << k;
        };
        if (KB == 8 && cut_bit >= 32) scan_window((uint32_t)0); else scan_window((O)0);
        // re-index by window position: bit j = slots (base-1+j, base+j) agree, j = 0 .. IPT; bit W-1 is the wrap
        eq = ((eq << rot) | (eq >> (W - rot))) & ((1u << (W - 1)) - 1u);
        // Window positions outside the tile ([0, lim)) hold no key: a pair that involves one never agrees,
        // whatever the placeholder compared equal to (a real key's swept bits may well be 0x...01, e.g. a raw
        // high word of 1 under a 32-bit cut).  Bit j pairs slots (base-1+j, base+j).
        if (base == 0) eq &= ~1u;
        if (lim - base <= IPT) eq &= (1u << max(lim - base, 0)) - 1u;
        uint32_t members = (eq | (eq >> 1)) & ((1u << IPT) - 1u);  // own slots with an equal neighbour
        if (!FULL) members &= lim - base >= IPT ? ~0u : (1u << max(lim - base, 0)) - 1u;  // sentinels are no members
        // Exact pairs that lie inside the own slots and whose surroundings are known from the window (four runs
        // out of five): one short iteration for both members instead of two general ones.
        {
          uint32_t ph = (eq >> 1) & ~eq & ~(eq >> 2) & (((1u << IPT) - 1u) >> 1);  // bit s: slots (s, s+1) are a run of two
          if (!FULL) ph &= members;
          members &= ~(ph | (ph << 1));
          while (ph) {
            const int sl = __ffs(ph) - 1;
            ph &= ph - 1;
            const int i = base + sl;
            const KeyT k0 = key_at(i), k1 = key_at(i + 1);
            // (order inside a run = order of the raw keys, reversed if the order mapping's low bits are ones: below)
            const O sgn = (O)0 - (((O)k0 >> (8 * KB - 1)) & 1);
            const O flip = (O)0 - (((O)ko.xor_const ^ (sgn & (O)ko.neg_xor)) & 1);
            if (((O)k1 ^ flip) < ((O)k0 ^ flip)) {
              sdelta[i] = (int8_t)1;
              sdelta[i + 1] = (int8_t)-1;
            }
          }
        }
        while (members) {
          const int sl = __ffs(members) - 1;
          members &= members - 1;
          const int i = base + sl;
          const KeyT kraw = key_at(i);
          // All members of a run share the swept bits, the sign among them: the order mapping xors them with
          // one and the same constant, whose bits below the cut are all zeros or all ones (no range reduction
          // here, and a left shift keeps the order).  So the full-key order inside the run is the order of the
          // raw keys, reversed if that constant's low bits are ones: compare raw keys xor-ed with 0 or ~0.
          const O sgn = (O)0 - (((O)kraw >> (8 * KB - 1)) & 1);
          const O flip = (O)0 - (((O)ko.xor_const ^ (sgn & (O)ko.neg_xor)) & 1);
          const O ok = (O)kraw ^ flip;
          // neighbours known to be in the run from the window bits, then (rarely) beyond the window
          const uint32_t below = ~eq & ((2u << sl) - 1u);  // zero bits at or below sl stop the run on the left
          int nl = below ? sl - (31 - __clz(below)) : sl + 1;
          const uint32_t above = ~(eq >> (sl + 1));
          int nr = __ffs(above) - 1;  // IPT - sl when the run reaches the end of the window
          // extent of the run: beyond the window only if it reaches the window's ends, and never further than
          // it takes to know that it is too long (nl + nr = run length - 1 for every member alike)
          if (nl == sl + 1 && nl + nr <= FIX_CAP) {  // includes slot base-1
            int j = i - nl - 1;
            while (j >= 0 && nl + nr <= FIX_CAP && (((O)key_at(j) ^ (O)kraw) & pmask) == 0) { nl++; j--; }
          }
          if (nr == IPT - sl && nl + nr <= FIX_CAP) {  // includes slot base+IPT
            int j = i + nr + 1;
            while (j < lim && nl + nr <= FIX_CAP && (((O)key_at(j) ^ (O)kraw) & pmask) == 0) { nr++; j++; }
          }
          int cl = 0, cr = 0;
          if (nl + nr > FIX_CAP) {
            // Too long to be ordered here: every member keeps its slot and checks its own (slot, slot+1) pair;
            // if none is out of order the run is in order as it stands (duplicates of one key, typically)
            // and nobody has to be told.
            if (((eq >> (sl + 1)) & 1u) && ((O)key_at(i + 1) ^ flip) < ok && ld_relaxed_u32(a.fix_flag) == 0) atomicOr(a.fix_flag, 1u);
          } else {
            // (runs are pairs four times out of five: the first neighbour on either side without a loop)
            if (nl) cl += ((O)key_at(i - 1) ^ flip) > ok;
            if (nr) cr += ((O)key_at(i + 1) ^ flip) < ok;
            for (int q = 2; q <= nl; q++) cl += ((O)key_at(i - q) ^ flip) > ok;
            for (int q = 2; q <= nr; q++) cr += ((O)key_at(i + q) ^ flip) < ok;
          }
          if (cr != cl) sdelta[i] = (int8_t)(cr - cl);
        }
      }
      // what junction_fix_kernel needs to know about this tile: the swept bits (as a 32-bit fingerprint: equal
      // bits give equal fingerprints) of the first and the last key of its part of every bucket
      if (a.jtable != nullptr && tid < RADIX) {
        const int s0 = (int)s_prefix[tid];
        const int s1 = min(tid + 1 < RADIX ? (int)s_prefix[tid + 1] : lim, lim);
        uint64_t w = 0;
        if (s1 > s0) {
          const O p0 = (O)key_at(s0) & pmask, p1 = (O)key_at(s1 - 1) & pmask;
          w = (uint64_t)((uint32_t)((uint64_t)p0 >> 32) ^ (uint32_t)p0) | ((uint64_t)((uint32_t)((uint64_t)p1 >> 32) ^ (uint32_t)p1) << 32);
        }
        a.jtable[(size_t)tile * RADIX + tid] = w;
      }
      __syncthreads();
    }
  }

  // ---- decoupled look-back (one thread per digit), after the staging stores so that the predecessors
  //      have had time to publish.  LB_BATCH predecessors are polled per round trip: the chain of
  //      dependent L2 loads is what this phase costs. ------------------------------------------------------
  if (tid < RADIX) {
    uint64_t excl;
    if (first_tile) {
      excl = tile0_base;
    } else {
      constexpr int LB_BATCH = 4;
      excl = 0;
      int64_t t = tile - 1;
      bool done = false;
      while (!done) {
        uint64_t w[LB_BATCH];
#pragma unroll
        for (int j = 0; j < LB_BATCH; j++) {
          const int64_t tt = t - j;
          w[j] = tt >= (int64_t)a.tile_first ? ld_relaxed_u64(&a.lookback[(size_t)tt * RADIX + tid]) : (LB_FLAG_PREFIX | tagbits);
        }
        int consumed = 0;
#pragma unroll
        for (int j = 0; j < LB_BATCH; j++) {
          const bool ready = (w[j] & LB_TAG_MASK) == tagbits && (w[j] & LB_FLAG_MASK) != 0;
          if (!done && consumed == j && ready) {  // consume in order, stop at the first unpublished word
            excl += w[j] & LB_VALUE_MASK;
            consumed = j + 1;
            if ((w[j] & LB_FLAG_MASK) == LB_FLAG_PREFIX) done = true;
          }
        }
        t -= consumed;
        if (consumed == 0 && a.spin_ns != 0) __nanosleep(a.spin_ns);  // optional back-off while nothing is published
      }
      st_relaxed_u64(&a.lookback[(size_t)tile * RADIX + tid], LB_FLAG_PREFIX | tagbits | (excl + valid_count));
    }
    gbase[tid] = (GOff)excl - (GOff)s_prefix[tid];  // (32-bit: wraps, the slot index added later brings it back)
  }
  if constexpr (!FIX) {
    if (NSTAGE == 2 && have_next) cp_async_wait<1>(); else cp_async_wait<0>();  // column 0 has landed (mine)
  }
  __syncthreads();  // column 0 and gbase visible to everyone

  GOff goff[IPT];
  int wr_s = 0;  // write cursor
  uint32_t wr_c = 0;
  int j0 = 0;    // first column the loop below writes
  // after column j has been written out: wait for / stage the next one
  auto next_column = [&](int j) {
    if (NSTAGE == 2) {
      cp_async_wait<0>();   // my part of column j+1 has landed
      __syncthreads();      // everyone is done writing column j out, and column j+1 is complete
      issue_next(bufs[j & 1]);  // column j+2 into the buffer column j just vacated
    } else {
      __syncthreads();      // everyone is done with the staging buffer
      issue_next(bufs[0]);
      cp_async_wait<0>();
      __syncthreads();
    }
  };
  if constexpr (use_sdigit) {
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const int i = tid + k * THREADS;
      goff[k] = (FULL || i < n_valid) ? (GOff)(gbase[sdigit[i]] + (GOff)(i + (FIX ? (int)sdelta[i] : 0))) : (GOff)0;
    }
  } else {
    // bucket of a slot = digit of the key staged there: byte (es/8) of the raw key, mapped like in the ranking
    const KeyT *kcol0 = reinterpret_cast<const KeyT *>(bufs[0]);
    const int kstep0 = soa_keys ? 1 : (int)(ks.chunk_bytes / KB);
    const DigitX dx = make_digitx(ko, a.shift);
    const uint32_t s1 = ((uint32_t)(es >> 3) & 3u) | 0x4440u;
    const bool direct = soa_keys && ks.chunk_bytes == KB;  // the key array is column 0: write it out on the way
    KeyT *kdst = reinterpret_cast<KeyT *>(ks.buf[sel ^ 1]);
    auto slot_loop = [&](auto direct_c) {
#pragma unroll
      for (int k = 0; k < IPT; k++) {
        // (groups of four: without the fence the scheduler hoists all IPT staged keys into registers at once
        //  and spills)
        if (k % 4 == 0 && k != 0) asm volatile("" ::: "memory");
        const int i = tid + k * THREADS;
        goff[k] = (GOff)0;
        if (FULL || i < n_valid) {
          const KeyT kv = kcol0[i * kstep0];
          uint32_t w, top;
          if constexpr (KB == 8) { top = (uint32_t)(kv >> 32); w = (uint32_t)(kv >> (es & 32)); }
          else { top = (uint32_t)kv << (32 - 8 * KB); w = (uint32_t)kv; }
          const uint32_t d = (prmt(w, 0u, s1) ^ dx.xc) ^ ((uint32_t)((int32_t)top >> 31) & dx.nx);
          goff[k] = (GOff)(gbase[d] + (GOff)(i + (FIX ? (int)sdelta[i] : 0)));
          if constexpr (decltype(direct_c)::value) kdst[goff[k]] = kv;
        }
      }
    };
    if (direct) {
      slot_loop(std::true_type{});
      wr_s = 1;
      j0 = 1;
      if (wr_s >= a.ss.n_streams) return;
      next_column(0);
    } else {
      slot_loop(std::false_type{});
    }
  }
  for (int j = j0;; j++) {
    {
      const Stream &st = a.ss.streams[wr_s];
      unsigned char *dst = st.buf[sel ^ 1];
      const unsigned char *buf = bufs[NSTAGE == 2 ? (j & 1) : 0];
      const uint32_t cb = st.chunk_bytes, cpe = st.chunks_per_elem;
      if (cb == 8) write_out<uint64_t, THREADS, IPT, FULL, LUT>(dst, buf, goff, n_valid, cpe, wr_c, pdelta, s_prefix, gbase64, a.lut_world, a.peer_wide != 0);
      else if (cb == 4) write_out<uint32_t, THREADS, IPT, FULL, LUT>(dst, buf, goff, n_valid, cpe, wr_c, pdelta, s_prefix, gbase64, a.lut_world, a.peer_wide != 0);
      else if (cb == 16) write_out<uint4, THREADS, IPT, FULL, LUT>(dst, buf, goff, n_valid, cpe, wr_c, pdelta, s_prefix, gbase64, a.lut_world, a.peer_wide != 0);
      else if (cb == 2) write_out<uint16_t, THREADS, IPT, FULL, LUT>(dst, buf, goff, n_valid, cpe, wr_c, pdelta, s_prefix, gbase64, a.lut_world, a.peer_wide != 0);
      else write_out<uint8_t, THREADS, IPT, FULL, LUT>(dst, buf, goff, n_valid, cpe, wr_c, pdelta, s_prefix, gbase64, a.lut_world, a.peer_wide != 0);
      if (++wr_c == cpe) { wr_c = 0; wr_s++; }
    }
    if (wr_s >= a.ss.n_streams) break;
    next_column(j);
  }
  if constexpr (LUT) {
    // Overlapped exchange: the destinations are told when a whole CHUNK of this rank's tiles has been delivered.
    // Every thread makes its peer stores visible system-wide, then the tile is counted; whoever completes the
    // chunk's count writes the chunk's arrival flag into every destination's flag array.
    if (a.sig_n != 0) {
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        uint32_t ch = 0;
        while (ch + 1 < a.sig_n && (uint32_t)tile >= a.sig_ct[ch + 1]) ch++;
        const uint32_t done = atomicAdd(&a.sig_done[ch], 1u) + 1u;
        if (done == a.sig_ct[ch + 1] - a.sig_ct[ch]) {
          __threadfence_system();
          for (int d = 0; d < a.lut_world; d++)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.sig_flags[d] + a.sig_me * 8u + ch), "r"(a.sig_value + ch) : "memory");
        }
      }
    }
  }
}

// ANYCHUNK = false: every stream moved by the loop has 4-, 8- or 16-byte chunks (the common shapes).
// ANYCHUNK = true additionally handles 1- and 2-byte chunks; with all five widths inlined ptxas needs
// far more registers per thread, so the narrow widths get their own instantiation.
template <int KB, int THREADS, int IPT, int MINB, int NSTAGE, bool ANYCHUNK, bool LUT, bool FIX = false, int RANK = RANK_BALLOT, bool BYTEWISE = false>
__global__ void __launch_bounds__(THREADS, MINB) onesweep_kernel(const __grid_constant__ SweepArgs a) {
  static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is assumed");
  constexpr int TILE = THREADS * IPT;
  constexpr int NWARPS = THREADS / 32;
  if (!a.plan_in_args && a.plan->skip[a.pass]) return;
  const uint32_t sel = a.plan_in_args ? a.arg_sel : a.plan->src_sel[a.pass];

  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t *warp_cnt = reinterpret_cast<uint32_t *>(smem + (size_t)NSTAGE * TILE * a.stage_bytes);
  constexpr int RANK_WORDS = NWARPS * RADIX;
  uint32_t *nhist = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(warp_cnt + RANK_WORDS) + RADIX * 8);  // = s_prefix
  // One tile per CTA.  (A persistent grid that loops over tickets was measured 17 % slower: CTAs that
  // start together stay in phase, so loads, look-backs and stores of the whole machine bunch up; fresh
  // CTAs per tile stagger naturally.  profiles/README.md)
  // The partition pass of the multi-GPU sort (LUT) is the exception: it is bound by the NVLink, not by the SMs,
  // and when its exchange is overlapped with the receivers' first pass it is launched with FEWER CTAs than
  // tiles -- about one and a half per SM -- which loop over the tickets, so that the other kernel's CTAs find
  // room on every SM the whole time (mgpu.cuh).
  __shared__ uint32_t s_tile;
  __shared__ __align__(8) uint64_t s_key_bar;
  uint32_t key_parity = 0;
  if (LUT && a.tma_keys && threadIdx.x == 0) mbar_init(&s_key_bar, 1);
  for (;;) {
    if (threadIdx.x == 0) {
      const uint32_t t = atomicAdd(&a.tile_counter[a.pass], 1u) + a.tile_first;
      s_tile = t;
      if (a.tma_keys && a.n - (int64_t)t * TILE >= TILE) {
        if (!LUT) mbar_init(&s_key_bar, 1);
        tma_load_tile(smem, a.ss.streams[0].buf[sel] + (size_t)t * TILE * KB, (uint32_t)(TILE * KB), &s_key_bar);
      }
      // (the host has checked alignment and sizes of streams prefetch_first .. prefetch_first + prefetch_cols - 1)
      if (a.prefetch_cols != 0 && a.n - (int64_t)t * TILE >= TILE)
        tma_prefetch_l2(a.prefetch_ptr[sel] + (size_t)t * a.prefetch_bytes, a.prefetch_bytes);
      // (Prefetching the tile ~444 tickets further on as well -- what the CTAs in flight will draw next -- was
      //  measured SLOWER: 37.5 vs 36.2 ms at 1e9 records, 38.4 with 1000 tickets: profiles/README.md.)
    }
    for (int i = threadIdx.x; i < RANK_WORDS / 4; i += THREADS) reinterpret_cast<uint4 *>(warp_cnt)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < RADIX; i += THREADS) nhist[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t remaining = a.n - tile * TILE;
    if (LUT && remaining <= 0) break;  // (looping CTAs: the tickets are used up)
    if (remaining >= TILE) {
      sweep_tile<KB, THREADS, IPT, NSTAGE, ANYCHUNK, LUT, FIX, RANK, BYTEWISE, true>(a, smem, tile, TILE, sel, &s_key_bar, key_parity);
      if (a.tma_keys) key_parity ^= 1u;
    } else {
      sweep_tile<KB, THREADS, IPT, NSTAGE, ANYCHUNK, LUT, FIX, RANK_BALLOT, BYTEWISE, false>(a, smem, tile, (int)remaining, sel, &s_key_bar, key_parity);
    }
    if (!LUT) break;
    __syncthreads();  // everyone is done with the tile's shared memory before the next ticket is drawn
  }
}

// ------------------------------------------------------------------------------------------------
// copy the result back into the caller's arrays when the last executed pass left it in the shadow
// ------------------------------------------------------------------------------------------------
struct CopyBackArgs {
  StreamSet ss;
  int64_t n;
  const Plan *plan;
  int force;  // 1: the host knows the result is in the shadow (no finish kernel delivered it)
  const uint32_t *skip_if;  // when set: do nothing if *skip_if != 0 (the segment finish ran and delivered the result)
};

static __global__ void __launch_bounds__(256) copyback_kernel(const __grid_constant__ CopyBackArgs a) {
  if (!a.force && (a.plan->final_sel == 0 || a.plan->cut_digit != 0)) return;  // (the segment finish already wrote side 0)
  if (a.skip_if != nullptr && *reinterpret_cast<const volatile uint32_t *>(a.skip_if) != 0) return;
  for (int s = 0; s < a.ss.n_streams; s++) {
    const Stream &st = a.ss.streams[s];
    const size_t bytes = (size_t)a.n * st.chunk_bytes * st.chunks_per_elem;
    const unsigned char *src = st.buf[1];
    unsigned char *dst = st.buf[0];
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t gsz = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const size_t nv = bytes / 16;
      const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
      uint4 *d4 = reinterpret_cast<uint4 *>(dst);
      for (size_t i = gtid; i < nv; i += gsz) d4[i] = s4[i];
      for (size_t i = nv * 16 + gtid; i < bytes; i += gsz) dst[i] = src[i];
    } else {
      // rare: caller array not 16-byte aligned; move in the stream's own chunk size
      const size_t nc = bytes / st.chunk_bytes;
      switch (st.chunk_bytes) {
        case 1: for (size_t i = gtid; i < nc; i += gsz) dst[i] = src[i]; break;
        case 2: for (size_t i = gtid; i < nc; i += gsz) ((uint16_t *)dst)[i] = ((const uint16_t *)src)[i]; break;
        case 4: for (size_t i = gtid; i < nc; i += gsz) ((uint32_t *)dst)[i] = ((const uint32_t *)src)[i]; break;
        case 8: for (size_t i = gtid; i < nc; i += gsz) ((uint64_t *)dst)[i] = ((const uint64_t *)src)[i]; break;
        default: for (size_t i = gtid; i < nc; i += gsz) ((uint4 *)dst)[i] = ((const uint4 *)src)[i]; break;
      }
    }
  }
}

}  // namespace b200sort
