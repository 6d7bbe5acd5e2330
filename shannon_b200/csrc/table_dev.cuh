// Device-side primitives of the K1-mer table (shared by table.cu, l3.cu, synth.cu), for the key
// width this translation unit is compiled for (common.cuh: shn_key_t, ShnSlot, SHN_BSLOTS).
//
// Bucket = 64 bytes, 64-byte aligned (four 16-byte slots for 64-bit keys, two 32-byte slots for
// 128-bit keys): one DRAM burst / two L2 sectors per probe, read with two 256-bit loads.  A key
// lives in the first bucket of its probe sequence (linear over buckets) that had a free slot when
// it was inserted.  An insert that walks past a full bucket sets that bucket's OVERFLOW flag
// (bit 30 of slot 0's weight word); a lookup stops at the first bucket that holds the key, has a
// free slot, or has no overflow flag -- so ~94 % of the probes for ABSENT keys (three of the four
// successor probes of every walk step) finish after one bucket at load 0.5.
#pragma once
#include "common.cuh"

// 32-byte global load: one L2 sector per request instead of two 16-byte requests that hit the
// same sector twice.  sm_100 has 256-bit LDG (SASS: LDG.E.ENL2.256).
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

// 128-bit compare-and-swap (SASS: ATOMG.E.CAS.128)
__device__ __forceinline__ u128 shn_cas128(u128* addr, u128 cmp, u128 val) {
  uint64_t clo = (uint64_t)cmp, chi = (uint64_t)(cmp >> 64), vlo = (uint64_t)val,
           vhi = (uint64_t)(val >> 64), olo, ohi;
  asm volatile(
      "{\n .reg .b128 c, v, o;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n"
      " atom.global.cas.b128 o, [%6], c, v;\n mov.b128 {%0, %1}, o;\n}"
      : "=l"(olo), "=l"(ohi)
      : "l"(clo), "l"(chi), "l"(vlo), "l"(vhi), "l"(addr)
      : "memory");
  return ((u128)ohi << 64) | olo;
}
__device__ __forceinline__ uint64_t shn_cas_key(uint64_t* addr, uint64_t cmp, uint64_t val) {
  return atomicCAS(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)cmp,
                   (unsigned long long)val);
}
__device__ __forceinline__ u128 shn_cas_key(u128* addr, u128 cmp, u128 val) {
  return shn_cas128(addr, cmp, val);
}

namespace SHN_NS {

struct ShnBucket {
  uint64_t w[8];
#ifdef SHN_WIDE   // slot j = {w[4j] | w[4j+1] << 64 = key, w[4j+2] = weight | idx << 32, pad}
  __device__ __forceinline__ shn_key_t key(int j) const { return ((u128)w[4 * j + 1] << 64) | w[4 * j]; }
  __device__ __forceinline__ uint64_t wi(int j) const { return w[4 * j + 2]; }
#else             // slot j = {w[2j] = key, w[2j+1] = weight | idx << 32}
  __device__ __forceinline__ shn_key_t key(int j) const { return w[2 * j]; }
  __device__ __forceinline__ uint64_t wi(int j) const { return w[2 * j + 1]; }
#endif
  __device__ __forceinline__ uint32_t weight(int j) const { return (uint32_t)wi(j); }
};

__device__ __forceinline__ void table_load_bucket(const ShnTableView& t, uint64_t b, ShnBucket* out) {
  const char* p = reinterpret_cast<const char*>(t.slots + SHN_BSLOTS * b);  // 64-byte aligned
  shn_ld256_cg(p, out->w);                                                  // .cg: data changes under atomics
  shn_ld256_cg(p + 32, out->w + 4);
}

// one slot: key, weight word, idx word
__device__ __forceinline__ void table_load_slot(const ShnSlot* slots, uint64_t s, shn_key_t* key,
                                                uint32_t* weight, uint32_t* idx) {
#ifdef SHN_WIDE
  uint64_t w[4];
  shn_ld256_cg(slots + s, w);
  *key = ((u128)w[1] << 64) | w[0];
  *weight = (uint32_t)w[2];
  *idx = (uint32_t)(w[2] >> 32);
#else
  const uint4 v = __ldcg(reinterpret_cast<const uint4*>(slots) + s);
  *key = ((uint64_t)v.y << 32) | v.x;
  *weight = v.z;
  *idx = v.w;
#endif
}

// Looks `key` up in an already loaded bucket.  Returns 1 = found (slot index in *j_out, raw weight
// word in *w_out, idx word in *idx_out), 0 = definitely absent, -1 = undecided: next bucket.
__device__ __forceinline__ int table_match_bucket2(const ShnBucket& bk, shn_key_t key, int* j_out,
                                                   uint32_t* w_out, uint32_t* idx_out) {
  bool has_empty = false;
  int found = -1;
  uint64_t wi = 0;
#pragma unroll
  for (int j = 0; j < SHN_BSLOTS; ++j) {  // fully unrolled: no dynamic register indexing
    shn_key_t k = bk.key(j);
    if (k == key) {
      found = j;
      wi = bk.wi(j);
    }
    has_empty |= k == SHN_EMPTY;
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
__device__ __forceinline__ int table_match_bucket(const ShnBucket& bk, shn_key_t key, int* j_out,
                                                  uint32_t* w_out) {
  uint32_t idx;
  return table_match_bucket2(bk, key, j_out, w_out, &idx);
}

// Read-only probe: slot index of `key` or ~0; *w_out = raw weight word (flag bits included).
// The all-ones key (only possible as a query: it is low-complexity and never stored) is absent.
__device__ __forceinline__ uint64_t table_find_from(const ShnTableView& t, shn_key_t key, uint64_t b,
                                                    uint32_t* w_out);
__device__ __forceinline__ uint64_t table_find(const ShnTableView& t, shn_key_t key, uint32_t* w_out) {
  if (key == SHN_EMPTY) return ~0ull;
  return table_find_from(t, key, t.bucket_of(key), w_out);
}
// ... starting at a home bucket the caller computed (ShnTableView::bucket_with_min)
__device__ __forceinline__ uint64_t table_find_from(const ShnTableView& t, shn_key_t key, uint64_t b,
                                                    uint32_t* w_out) {
  if (key == SHN_EMPTY) return ~0ull;
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

// The four successors (x[1:] . b) of the K1-mer x share their first K bases, hence their home bucket
// and their probe sequence (ShnTableView::bucket_of): one walk over the buckets finds all of them.
// Returns the mask of the successors that exist; slot[b] / w[b] = slot index / raw weight word.
__device__ __forceinline__ uint32_t table_find_successors(const ShnTableView& t, shn_key_t x, int k1,
                                                           uint64_t* slot, uint32_t* w) {
  const shn_key_t pre = (x << 2) & shn_key_mask(k1);
  uint64_t b = t.bucket_of(pre);
  uint32_t found = 0;
  for (;;) {
    ShnBucket bk;
    table_load_bucket(t, b, &bk);
    bool has_empty = false;
#pragma unroll
    for (int j = 0; j < SHN_BSLOTS; ++j) {
      const shn_key_t k = bk.key(j);
      has_empty |= k == SHN_EMPTY;
      if (k != SHN_EMPTY && (k & ~(shn_key_t)3) == pre) {
        const int c = (int)((uint32_t)k & 3u);
        found |= 1u << c;
#pragma unroll
        for (int q = 0; q < 4; ++q)  // static indices: the arrays stay in registers
          if (q == c) {
            slot[q] = SHN_BSLOTS * b + j;
            w[q] = bk.weight(j);
          }
      }
    }
    if (has_empty || !(bk.weight(0) & SHN_OVERFLOW)) return found;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
}

// Finds or claims the slot of `key`; returns its global slot index (~0 if the table is full);
// *is_new += 1 if this call claimed a free slot.
__device__ __forceinline__ uint64_t table_upsert_slot(const ShnTableView& t, shn_key_t key,
                                                      int* is_new) {
  uint64_t b = t.bucket_of(key);
  for (uint64_t probes = 0; probes < t.n_buckets; ++probes) {
    ShnSlot* s = t.slots + SHN_BSLOTS * b;
    ShnBucket bk;
    table_load_bucket(t, b, &bk);
#pragma unroll
    for (int j = 0; j < SHN_BSLOTS; ++j) {
      shn_key_t cur = bk.key(j);
      if (cur == key) return SHN_BSLOTS * b + j;
      if (cur == SHN_EMPTY) {
        shn_key_t old = shn_cas_key(&s[j].key, SHN_EMPTY, key);
        if (old == SHN_EMPTY) {
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

// Insert (key, +w, min idx) in one pass.  64-bit keys: a NEW key is published with a single
// 128-bit CAS of the whole slot {key, w, idx} against the pristine pattern {EMPTY, 0, 0xFFFFFFFF}
// (one atomic round trip instead of CAS + add + min); only a key that is already present takes
// the add/min path.  Returns the slot (~0: table full); *old_w = weight before this insert.
// min_idx == false: an already present key keeps its idx word (the caller resolves first occurrences).
__device__ __forceinline__ uint64_t table_insert_add(const ShnTableView& t, shn_key_t key, uint32_t w,
                                                     uint32_t idx, int* is_new, uint32_t* old_w,
                                                     bool min_idx = true) {
#ifdef SHN_WIDE
  // 128-bit keys: the key is claimed with a 128-bit CAS; a freshly claimed slot then gets its
  // {weight, idx} pair with ONE 64-bit CAS against the pristine {0, 0xFFFFFFFF} (it fails only if a
  // duplicate line or an overflow marker got there first: then add/min as for a present key)
  const int before = *is_new;
  uint64_t slot = table_upsert_slot(t, key, is_new);
  if (slot == ~0ull) return slot;
  if (*is_new != before) {
    unsigned long long* wi = reinterpret_cast<unsigned long long*>(&t.slots[slot].weight);
    if (atomicCAS(wi, 0xFFFFFFFF00000000ull, ((unsigned long long)idx << 32) | w) == 0xFFFFFFFF00000000ull) {
      *old_w = 0;
      return slot;
    }
  }
  *old_w = atomicAdd(&t.slots[slot].weight, w) & SHN_WEIGHT_MASK;
  // (the thread that claimed the key always records its index, even when a concurrent duplicate
  // beat its {weight, idx} CAS)
  if (min_idx || *is_new != before) atomicMin(&t.slots[slot].idx, idx);
  return slot;
#else
  const u128 pristine = ((u128)0xFFFFFFFF00000000ull << 64) | (u128)SHN_EMPTY_KEY;
  const u128 fresh = ((u128)(((uint64_t)idx << 32) | w) << 64) | (u128)key;
  uint64_t b = t.bucket_of(key);
  for (uint64_t probes = 0; probes < t.n_buckets; ++probes) {
    ShnSlot* s = t.slots + SHN_BSLOTS * b;
    ShnBucket bk;
    table_load_bucket(t, b, &bk);
    uint64_t hit = ~0ull;
#pragma unroll
    for (int j = 0; j < SHN_BSLOTS; ++j) {
      if (hit != ~0ull) continue;
      shn_key_t cur = bk.key(j);
      if (cur == SHN_EMPTY) {
        const u128 old = shn_cas128(reinterpret_cast<u128*>(&s[j]), pristine, fresh);
        if (old == pristine) {
          *is_new += 1;
          *old_w = 0;
          return SHN_BSLOTS * b + j;
        }
        cur = (shn_key_t)old;  // somebody else took the slot meanwhile
      }
      if (cur == key) hit = SHN_BSLOTS * b + j;
    }
    if (hit != ~0ull) {
      *old_w = atomicAdd(&t.slots[hit].weight, w) & SHN_WEIGHT_MASK;
      if (min_idx) atomicMin(&t.slots[hit].idx, idx);
      return hit;
    }
    // every slot holds another key: leave the trail marker for lookups, then move on
    if (!(bk.weight(0) & SHN_OVERFLOW)) atomicOr(&s[0].weight, SHN_OVERFLOW);
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return ~0ull;
#endif
}

// slot initialisation pattern: key = all ones, weight = 0, idx = `idx0`
__device__ __forceinline__ void table_store_empty(ShnSlot* slots, uint64_t s, uint32_t idx0) {
#ifdef SHN_WIDE
  uint4* p = reinterpret_cast<uint4*>(slots + s);
  p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  p[1] = make_uint4(0u, idx0, 0u, 0u);
#else
  reinterpret_cast<uint4*>(slots)[s] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, idx0);
#endif
}

// warp shuffle of a key
__device__ __forceinline__ shn_key_t shfl_key(shn_key_t k, int src_lane) {
#ifdef SHN_WIDE
  uint64_t lo = __shfl_sync(0xFFFFFFFFu, (uint64_t)k, src_lane);
  uint64_t hi = __shfl_sync(0xFFFFFFFFu, (uint64_t)(k >> 64), src_lane);
  return ((u128)hi << 64) | lo;
#else
  return __shfl_sync(0xFFFFFFFFu, k, src_lane);
#endif
}

}  // namespace SHN_NS
