// Lock-free union-find on a uint32 parent array (shared by l3.cu: raw components of the K1-mer
// successor graph on one table; shard.cu: the same across ranks).  Parents always have smaller
// indices than their children, so the root of a finished set is its minimum index.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  uint32_t p = __ldcg(&parent[x]);
  while (p != x) {
    uint32_t gp = __ldcg(&parent[p]);
    if (gp != p) parent[x] = gp;  // path halving; x is not a root, so this never races a link
    x = p;
    p = gp;
  }
  return x;
}

// read-only variant (no path compression): safe while other threads overwrite parent[i] <- root
__device__ __forceinline__ uint32_t uf_find_ro(const uint32_t* parent, uint32_t x) {
  uint32_t p = __ldcg(&parent[x]);
  while (p != x) {
    x = p;
    p = __ldcg(&parent[x]);
  }
  return x;
}

// Rem's algorithm with splicing, lock-free (CAS): the two find paths are climbed together, always
// on the side whose parent has the larger index, and every node passed is re-pointed at the other
// side's (smaller) parent; the climb stops as soon as the paths meet instead of walking both to
// their roots.  Parents always have smaller indices than their children, so the forest stays
// acyclic under any interleaving, and the root of a finished set is its minimum slot index.
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  uint32_t pa = __ldcg(&parent[a]), pb = __ldcg(&parent[b]);
  while (pa != pb) {
    if (pa < pb) {  // climb on the side with the larger parent
      uint32_t t = a;
      a = b;
      b = t;
      t = pa;
      pa = pb;
      pb = t;
    }
    if (a == pa) {  // a is a root: link it under the other side
      const uint32_t old = atomicCAS(&parent[a], a, pb);
      if (old == a) return;
      pa = old;  // somebody linked it first
      continue;
    }
    atomicCAS(&parent[a], pa, pb);  // splice (harmless if it fails: the entry only ever decreases)
    a = pa;
    pa = __ldcg(&parent[a]);
  }
}

