// Self-join of (key, owner, pos) entries on equal keys -- the data-parallel closed form of the
// reference's two inverted-list loops:
//   * duplicate_check / rmer_to_contig   (extension_correction.py:247-270, 393-397), r = 15
//   * cmer_to_contig / contig_connections (extension_correction.py:372-389),          C = K1-1
// Both loops walk, for every position i of a later contig j, the list of earlier contigs d that
// contain the same r-mer / C-mer (one list entry per occurrence).  Per ordered pair (j > d) the
// loops only ever need:   count  = number of (position of j, occurrence in d) matches,
//                         min_i / max_i = first / last position of j with a match,
//                         covered = | union of [i, i+r) over matching positions i of j |.
// This routine produces exactly that pair table, sorted by (j, d).
#pragma once
#include "common.cuh"

struct PairTable {
  DevBuf hi, lo;        // uint32 owner ids, hi > lo
  DevBuf count;         // uint32
  DevBuf min_i, max_i;  // uint32 positions in `hi`
  DevBuf covered;       // uint32
  uint64_t n = 0;
};

// keys: n entries generated in (owner ascending, pos ascending) order; key_bits: significant bits.
// r: interval length for `covered`.
void shn_self_join(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_owner,
                   const uint32_t* d_pos, uint64_t n, int key_bits, uint32_t r, PairTable* out,
                   const char* tag);
