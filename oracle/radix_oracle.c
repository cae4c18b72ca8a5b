/*
 * oracle/radix_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, scalar, type-erased restatement of the reference's bit-wise MSB
 * radix sort (jonicho/simd-radix-sort).  It exists so that the CUDA path can be
 * checked against the reference's algorithm on machines where the reference's
 * AVX-512 build (oracle/_ref) cannot run.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load this file's
 * library; the product (simd-radix-sort_b200/, include/) never does.
 *
 * Parity pin: tests/test_oracle.py checks this port (a) against the reference
 * itself compiled from /root/reference (oracle/_ref/libref_sort.so, built by
 * oracle/Makefile), (b) against the golden vectors in tests/golden/ that were
 * produced by that compiled reference, and (c) against the known-answer values
 * recorded in SURVEY.md section 4 (Data<uint32_t,uint32_t>(n, Uniform, 42)).
 *
 * What follows what (all paths relative to /root/reference):
 *   key_bit()            src/radix_sort.hpp:26-29   isBitSet (bit_cast to UInt)
 *   bit_dir_up()         src/radix_sort.hpp:51-64   bitDirUp<T,Up,IsHighestBit,IsRightSide>
 *   sort_bit_seq()       src/radix_sort.hpp:66-92   BitSorterSequential::sortBit
 *   insertion_sort()     src/cmp_sorters.hpp:18-37  CmpSorterInsertionSort::sort
 *   radix_recursion()    src/radix_sort.hpp:270-295 radixRecursion
 *   oracle_sort_soa()    src/radix_sort.hpp:297-312,334-337  sort(thresh,num,keys,payloads...)
 *   oracle_sort_aos()    src/radix_sort.hpp:314-332 sort(thresh,num,DataElement*)
 *   oracle_gen_*()       src/data.hpp:105-170,364-406 Data<> generators (Uniform, payloads)
 *
 * The reference's default bit sorter is the AVX-512 BitSorterSIMD; this port
 * follows the scalar BitSorterSequential that the reference ships next to it.
 * Both realise the same partition predicate, so the sorted KEY sequence is the
 * same; the order of payloads among equal keys may differ (the reference is
 * not stable either way), which is why payload parity is defined per
 * equal-key run as a multiset.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
  KT_U8 = 0, KT_I8 = 1, KT_U16 = 2, KT_I16 = 3, KT_U32 = 4,
  KT_I32 = 5, KT_U64 = 6, KT_I64 = 7, KT_F32 = 8, KT_F64 = 9
};

#define MAX_STREAMS 72
#define MAX_ELEM 64

typedef struct {
  int key_type;
  int key_bytes;
  int up;
  int cmp_sorter;          /* 0 = insertion sort, 1 = none (CmpSorterNoSort) */
  int64_t thresh;
  int n_streams;           /* stream 0 carries the key at byte offset 0 */
  unsigned char *base[MAX_STREAMS];
  uint32_t elem[MAX_STREAMS];
} ctx_t;

static int key_bytes_of(int kt) {
  switch (kt) {
    case KT_U8: case KT_I8: return 1;
    case KT_U16: case KT_I16: return 2;
    case KT_U32: case KT_I32: case KT_F32: return 4;
    case KT_U64: case KT_I64: case KT_F64: return 8;
    default: return 0;
  }
}
static int is_signed_int(int kt) { return kt == KT_I8 || kt == KT_I16 || kt == KT_I32 || kt == KT_I64; }
static int is_float(int kt) { return kt == KT_F32 || kt == KT_F64; }

static inline const unsigned char *key_ptr(const ctx_t *c, int64_t i) {
  return c->base[0] + (size_t)i * c->elem[0];
}

/* src/radix_sort.hpp:26-29 */
static inline int key_bit(const ctx_t *c, int64_t i, int bit_no) {
  uint64_t v = 0;
  memcpy(&v, key_ptr(c, i), (size_t)c->key_bytes); /* little endian host */
  return (int)((v >> bit_no) & 1u);
}

/* src/radix_sort.hpp:51-64 */
static inline int bit_dir_up(const ctx_t *c, int is_highest_bit, int is_right_side) {
  if (is_float(c->key_type)) return is_highest_bit ? !c->up : is_right_side;
  if (is_signed_int(c->key_type)) return is_highest_bit ? !c->up : c->up;
  return c->up;
}

static inline void swap_elems(const ctx_t *c, int64_t a, int64_t b) {
  unsigned char tmp[MAX_ELEM];
  for (int s = 0; s < c->n_streams; s++) {
    size_t e = c->elem[s];
    unsigned char *pa = c->base[s] + (size_t)a * e;
    unsigned char *pb = c->base[s] + (size_t)b * e;
    memcpy(tmp, pa, e);
    memcpy(pa, pb, e);
    memcpy(pb, tmp, e);
  }
}

/* src/radix_sort.hpp:66-92 */
static int64_t sort_bit_seq(const ctx_t *c, int bit_no, int64_t left, int64_t right,
                            int is_highest_bit, int is_right_side) {
  const int dir = bit_dir_up(c, is_highest_bit, is_right_side);
  int64_t l = left, r = right;
  while (l <= r) {
    while (l <= r && (dir != key_bit(c, l, bit_no))) l++;
    while (l <= r && ((!dir) != key_bit(c, r, bit_no))) r--;
    if (l < r) swap_elems(c, l, r);
  }
  return l;
}

/* language-level  a < b  on the key type (IEEE compare for floats),
 * src/cmp_sorters.hpp:26 and src/data.hpp:29-30 */
static inline int key_less(int kt, const void *a, const void *b) {
  switch (kt) {
#define CASE(T, CT) case T: { CT x, y; memcpy(&x, a, sizeof x); memcpy(&y, b, sizeof y); return x < y; }
    CASE(KT_U8, uint8_t) CASE(KT_I8, int8_t) CASE(KT_U16, uint16_t) CASE(KT_I16, int16_t)
    CASE(KT_U32, uint32_t) CASE(KT_I32, int32_t) CASE(KT_U64, uint64_t) CASE(KT_I64, int64_t)
    CASE(KT_F32, float) CASE(KT_F64, double)
#undef CASE
  }
  return 0;
}

/* src/cmp_sorters.hpp:18-37 */
static void insertion_sort(const ctx_t *c, int64_t left, int64_t right) {
  unsigned char saved[MAX_STREAMS][MAX_ELEM];
  for (int64_t i = left + 1; i <= right; i++) {
    for (int s = 0; s < c->n_streams; s++)
      memcpy(saved[s], c->base[s] + (size_t)i * c->elem[s], c->elem[s]);
    int64_t j = i;
    while (j > left && (c->up ? key_less(c->key_type, saved[0], key_ptr(c, j - 1))
                              : key_less(c->key_type, key_ptr(c, j - 1), saved[0]))) {
      for (int s = 0; s < c->n_streams; s++)
        memcpy(c->base[s] + (size_t)j * c->elem[s], c->base[s] + (size_t)(j - 1) * c->elem[s], c->elem[s]);
      j--;
    }
    for (int s = 0; s < c->n_streams; s++)
      memcpy(c->base[s] + (size_t)j * c->elem[s], saved[s], c->elem[s]);
  }
}

/* src/radix_sort.hpp:270-295 */
static void radix_recursion(const ctx_t *c, int bit_no, int64_t left, int64_t right,
                            int is_right_side, int is_highest_bit) {
  if (right - left <= 0) return;
  if (right - left < c->thresh) {
    if (c->cmp_sorter == 0) insertion_sort(c, left, right);
    return;
  }
  const int64_t split = sort_bit_seq(c, bit_no, left, right, is_highest_bit, is_right_side);
  if (bit_no > 0) {
    radix_recursion(c, bit_no - 1, left, split - 1, is_highest_bit ? 0 : is_right_side, 0);
    radix_recursion(c, bit_no - 1, split, right, is_highest_bit ? 1 : is_right_side, 0);
  }
}

/* src/radix_sort.hpp:297-312 (thresh form) and :334-337 (thresh = 16).
 * cmp_sorter: 0 insertion sort (default), 1 CmpSorterNoSort (src/cmp_sorters.hpp:66-78). */
int oracle_sort_soa(void *keys, int key_type, int64_t num, int up, int n_payloads,
                    void *const *payloads, const uint32_t *payload_elem_bytes,
                    int64_t thresh, int cmp_sorter) {
  ctx_t c;
  memset(&c, 0, sizeof c);
  c.key_type = key_type;
  c.key_bytes = key_bytes_of(key_type);
  if (c.key_bytes == 0 || n_payloads < 0 || n_payloads + 1 > MAX_STREAMS) return -1;
  c.up = up ? 1 : 0;
  c.thresh = thresh;
  c.cmp_sorter = cmp_sorter;
  c.n_streams = 1 + n_payloads;
  c.base[0] = (unsigned char *)keys;
  c.elem[0] = (uint32_t)c.key_bytes;
  for (int p = 0; p < n_payloads; p++) {
    if (payload_elem_bytes[p] == 0 || payload_elem_bytes[p] > MAX_ELEM) return -2;
    c.base[1 + p] = (unsigned char *)payloads[p];
    c.elem[1 + p] = payload_elem_bytes[p];
  }
  radix_recursion(&c, 8 * c.key_bytes - 1, 0, num - 1, 0, 1);
  return 0;
}

/* src/radix_sort.hpp:314-332: AoS records, power-of-two size, key at offset 0 */
int oracle_sort_aos(void *records, int key_type, uint32_t record_bytes, int64_t num, int up,
                    int64_t thresh, int cmp_sorter) {
  ctx_t c;
  memset(&c, 0, sizeof c);
  c.key_type = key_type;
  c.key_bytes = key_bytes_of(key_type);
  if (c.key_bytes == 0) return -1;
  if (record_bytes == 0 || record_bytes > MAX_ELEM || (record_bytes & (record_bytes - 1)) != 0 ||
      record_bytes < (uint32_t)c.key_bytes)
    return -3; /* the reference static_asserts the power-of-two rule, :318-319 */
  c.up = up ? 1 : 0;
  c.thresh = thresh;
  c.cmp_sorter = cmp_sorter;
  c.n_streams = 1;
  c.base[0] = (unsigned char *)records;
  c.elem[0] = record_bytes;
  radix_recursion(&c, 8 * c.key_bytes - 1, 0, num - 1, 0, 1);
  return 0;
}

/* ------------------------------------------------------------------------ *
 * Generators restating src/data.hpp so that config 1's input can be rebuilt
 * on a box that has no /root/reference.
 * ------------------------------------------------------------------------ */

/* std::mt19937 (32-bit Mersenne Twister, Matsumoto & Nishimura 1998), the
 * engine src/data.hpp:108 seeds with `seed`. */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* src/data.hpp:364-371 with K = uint32_t: uniform_int_distribution over the
 * generator's full range returns the raw 32-bit draw (libstdc++). */
void oracle_gen_uniform_u32(uint32_t *keys, int64_t num, uint32_t seed) {
  mt19937_t g;
  mt_seed(&g, seed);
  for (int64_t i = 0; i < num; i++) keys[i] = mt_next(&g);
}

/* src/data.hpp:393-406 + :57-63: payload bytes are glibc rand() after
 * srand(first min(sizeof K, 4) bytes of the key); one rand() per payload byte,
 * payload streams in declaration order. */
void oracle_make_payloads(const void *keys, int key_bytes, int64_t num, int n_payloads,
                          void *const *payloads, const uint32_t *payload_elem_bytes) {
  for (int64_t i = 0; i < num; i++) {
    unsigned int seed = 0;
    memcpy(&seed, (const unsigned char *)keys + (size_t)i * key_bytes,
           key_bytes < 4 ? (size_t)key_bytes : 4u);
    srand(seed);
    for (int p = 0; p < n_payloads; p++) {
      unsigned char *dst = (unsigned char *)payloads[p] + (size_t)i * payload_elem_bytes[p];
      for (uint32_t b = 0; b < payload_elem_bytes[p]; b++) dst[b] = (unsigned char)rand();
    }
  }
}

/* src/data.hpp:249-270 checkPayloads: 1 if every payload matches its key */
int oracle_check_payloads(const void *keys, int key_bytes, int64_t num, int n_payloads,
                          void *const *payloads, const uint32_t *payload_elem_bytes) {
  for (int64_t i = 0; i < num; i++) {
    unsigned int seed = 0;
    memcpy(&seed, (const unsigned char *)keys + (size_t)i * key_bytes,
           key_bytes < 4 ? (size_t)key_bytes : 4u);
    srand(seed);
    for (int p = 0; p < n_payloads; p++) {
      const unsigned char *src = (const unsigned char *)payloads[p] + (size_t)i * payload_elem_bytes[p];
      for (uint32_t b = 0; b < payload_elem_bytes[p]; b++)
        if (src[b] != (unsigned char)rand()) return 0;
    }
  }
  return 1;
}

/* src/data.hpp:195-220 isSorted with the language-level compare */
int oracle_is_sorted(const void *keys, int key_type, uint32_t stride, int64_t num, int up) {
  const unsigned char *k = (const unsigned char *)keys;
  for (int64_t i = 1; i < num; i++) {
    const void *a = k + (size_t)(i - 1) * stride, *b = k + (size_t)i * stride;
    if (up ? key_less(key_type, b, a) : key_less(key_type, a, b)) return 0;
  }
  return 1;
}
