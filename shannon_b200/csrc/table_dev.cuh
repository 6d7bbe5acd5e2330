// Device-side primitives of the K1-mer table (shared by table.cu, l3.cu, synth.cu).
//
// Bucket = SHN_BSLOTS (4) slots of 16 bytes = 64 bytes, 64-byte aligned: one DRAM burst / two L2
// sectors per probe.  A key lives in the first bucket of its probe sequence (linear over buckets)
// that had a free slot when it was inserted.  An insert that walks past a full bucket sets that
// bucket's OVERFLOW flag (bit 30 of slot 0's weight word); a lookup stops at the first bucket that
// holds the key, has a free slot, or has no overflow flag -- so ~95 % of the probes for ABSENT keys
// (three of the four successor probes of every walk step) finish after one bucket at load 0.5.
#pragma once
#include "common.cuh"

// 32-byte (two-slot) global load: one L2 sector per request instead of two 16-byte requests that
// hit the same sector twice.  sm_100 has 256-bit LDG (SASS: LDG.E.ENL2.256).
__device__ __forceinline__ void shn_ld256_cg(const void* p, uint64_t* w) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
               : "l"(p));
}
__device__ __forceinline__ void shn_ld256_nc(const void* p, uint64_t* w) {
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
               : "l"(p));
}

struct ShnBucket {
  uint64_t w[2 * SHN_BSLOTS];  // slot j = {w[2j] = key, w[2j+1] = weight | idx << 32}
  __device__ __forceinline__ uint64_t key(int j) const { return w[2 * j]; }
  __device__ __forceinline__ uint32_t weight(int j) const { return (uint32_t)w[2 * j + 1]; }
};

__device__ __forceinline__ void table_load_bucket(const ShnTableView& t, uint64_t b, ShnBucket* out) {
  const ShnSlot* p = t.slots + SHN_BSLOTS * b;  // 64-byte aligned
  shn_ld256_cg(p, out->w);                      // .cg: data changes under atomics
  shn_ld256_cg(p + 2, out->w + 4);
}

// Looks `key` up in an already loaded bucket.  Returns 1 = found (slot index in *j_out, raw weight
// word in *w_out), 0 = definitely absent, -1 = undecided: continue with the next bucket.
__device__ __forceinline__ int table_match_bucket(const ShnBucket& bk, uint64_t key, int* j_out,
                                                  uint32_t* w_out) {
  bool has_empty = false;
  int found = -1;
  uint32_t w = 0;
#pragma unroll
  for (int j = 0; j < SHN_BSLOTS; ++j) {  // fully unrolled: no dynamic register indexing
    uint64_t k = bk.key(j);
    if (k == key) {
      found = j;
      w = bk.weight(j);
    }
    has_empty |= k == SHN_EMPTY_KEY;
  }
  if (found >= 0) {
    *j_out = found;
    *w_out = w;
    return 1;
  }
  if (has_empty || !(bk.weight(0) & SHN_OVERFLOW)) return 0;
  return -1;
}

// Same, also returning the slot's idx word (the claim stamp of the speculative walks).
__device__ __forceinline__ int table_match_bucket2(const ShnBucket& bk, uint64_t key, int* j_out,
                                                   uint32_t* w_out, uint32_t* idx_out) {
  bool has_empty = false;
  int found = -1;
  uint64_t wi = 0;
#pragma unroll
  for (int j = 0; j < SHN_BSLOTS; ++j) {
    uint64_t k = bk.key(j);
    if (k == key) {
      found = j;
      wi = bk.w[2 * j + 1];
    }
    has_empty |= k == SHN_EMPTY_KEY;
  }
  if (found >= 0) {
    *j_out = found;
    *w_out = (uint32_t)wi;
    *idx_out = (uint32_t)(wi >> 32);
    return 1;
  }
  if (has_empty || !(bk.weight(0) & SHN_OVERFLOW)) return 0;
  return -1;
}

// Read-only probe: slot index of `key` or ~0; *w_out = raw weight word (flag bits included).
__device__ __forceinline__ uint64_t table_find(const ShnTableView& t, uint64_t key, uint32_t* w_out) {
  uint64_t b = t.bucket_of(key);
  for (;;) {
    ShnBucket bk;
    table_load_bucket(t, b, &bk);
    int j = 0;
    int r = table_match_bucket(bk, key, &j, w_out);
    if (r == 1) return SHN_BSLOTS * b + j;
    if (r == 0) return ~0ull;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
}

// Latency-critical variant for dependent probe chains (walks): loads the home bucket AND its
// successor in the same memory round, so the ~5 % of probes that have to continue past an
// overflowed bucket do not pay a second DRAM round trip.
__device__ __forceinline__ uint64_t table_find2(const ShnTableView& t, uint64_t key, uint32_t* w_out) {
  uint64_t b = t.bucket_of(key);
  uint64_t b_next = (b + 1 == t.n_buckets) ? 0 : b + 1;
  ShnBucket bk0, bk1;
  table_load_bucket(t, b, &bk0);
  table_load_bucket(t, b_next, &bk1);
  int j = 0;
  int r = table_match_bucket(bk0, key, &j, w_out);
  if (r == 1) return SHN_BSLOTS * b + j;
  if (r == 0) return ~0ull;
  r = table_match_bucket(bk1, key, &j, w_out);
  if (r == 1) return SHN_BSLOTS * b_next + j;
  if (r == 0) return ~0ull;
  b = (b_next + 1 == t.n_buckets) ? 0 : b_next + 1;
  for (;;) {  // third bucket and beyond: ~0.3 % of probes
    table_load_bucket(t, b, &bk0);
    r = table_match_bucket(bk0, key, &j, w_out);
    if (r == 1) return SHN_BSLOTS * b + j;
    if (r == 0) return ~0ull;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
}

// Finds or claims the slot of `key`; returns its global slot index (~0 if the table is full);
// *is_new += 1 if this call claimed a free slot.
__device__ __forceinline__ uint64_t table_upsert_slot(const ShnTableView& t, uint64_t key,
                                                      int* is_new) {
  uint64_t b = t.bucket_of(key);
  for (uint64_t probes = 0; probes < t.n_buckets; ++probes) {
    ShnSlot* s = t.slots + SHN_BSLOTS * b;
    ShnBucket bk;
    table_load_bucket(t, b, &bk);
#pragma unroll
    for (int j = 0; j < SHN_BSLOTS; ++j) {
      uint64_t cur = bk.key(j);
      if (cur == key) return SHN_BSLOTS * b + j;
      if (cur == SHN_EMPTY_KEY) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&s[j].key),
                                           (unsigned long long)SHN_EMPTY_KEY,
                                           (unsigned long long)key);
        if (old == SHN_EMPTY_KEY) {
          *is_new += 1;
          return SHN_BSLOTS * b + j;
        }
        if (old == key) return SHN_BSLOTS * b + j;
        // somebody else took this slot for a different key: keep scanning
      }
    }
    // every slot holds another key: leave the trail marker for lookups, then move on
    if (!(bk.weight(0) & SHN_OVERFLOW)) atomicOr(&s[0].weight, SHN_OVERFLOW);
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return ~0ull;
}
