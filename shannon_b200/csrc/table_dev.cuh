// Device-side primitives of the K1-mer table (shared by table.cu, l3.cu, synth.cu).
#pragma once
#include "common.cuh"

// Finds or claims the slot of `key`; returns its global slot index (~0 if the table is full);
// *is_new += 1 if this call claimed a free slot.
__device__ __forceinline__ uint64_t table_upsert_slot(const ShnTableView& t, uint64_t key,
                                                      int* is_new) {
  uint64_t b = t.bucket_of(key);
  for (uint64_t probes = 0; probes < t.n_buckets; ++probes) {
    ShnSlot* s = t.slots + 2 * b;
    // both slots of the bucket in one 32-byte sector; .cg: concurrent CAS traffic lives in L2
    const ulonglong2 s0 = __ldcg(reinterpret_cast<const ulonglong2*>(&s[0]));
    const ulonglong2 s1 = __ldcg(reinterpret_cast<const ulonglong2*>(&s[1]));
    uint64_t k[2] = {s0.x, s1.x};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint64_t cur = k[j];
      if (cur == key) return 2 * b + j;
      if (cur == SHN_EMPTY_KEY) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&s[j].key),
                                           (unsigned long long)SHN_EMPTY_KEY,
                                           (unsigned long long)key);
        if (old == SHN_EMPTY_KEY) {
          *is_new += 1;
          return 2 * b + j;
        }
        if (old == key) return 2 * b + j;
        // somebody else took this slot for a different key: keep probing
      }
    }
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return ~0ull;
}

// Read-only probe: slot index of `key` or ~0; *w_out = raw weight word (bit 31 = traversed).
__device__ __forceinline__ uint64_t table_find(const ShnTableView& t, uint64_t key,
                                               uint32_t* w_out) {
  uint64_t b = t.bucket_of(key);
  for (;;) {
    const ShnSlot* s = t.slots + 2 * b;
    const uint4 s0 = __ldcg(reinterpret_cast<const uint4*>(&s[0]));
    const uint4 s1 = __ldcg(reinterpret_cast<const uint4*>(&s[1]));
    uint64_t k0 = ((uint64_t)s0.y << 32) | s0.x, k1 = ((uint64_t)s1.y << 32) | s1.x;
    if (k0 == key) {
      *w_out = s0.z;
      return 2 * b;
    }
    if (k1 == key) {
      *w_out = s1.z;
      return 2 * b + 1;
    }
    if (k0 == SHN_EMPTY_KEY || k1 == SHN_EMPTY_KEY) return ~0ull;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
}
