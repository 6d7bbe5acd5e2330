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
#include <string>

#include "common.cuh"

struct PairTable {
  DevBuf hi, lo;        // uint32 owner ids, hi > lo
  DevBuf count;         // uint32
  DevBuf min_i, max_i;  // uint32 positions in `hi`
  DevBuf covered;       // uint32
  uint64_t n = 0;
};

// Two phases, so that the duplicate filter can join in rank blocks and skip partners that are
// already known to be rejected (most candidates are near-duplicates of a few accepted contigs:
// joining everything against everything makes ~8 match events per r-mer entry, joining against
// accepted + same-block candidates ~2).
struct SelfJoin {
  shn_ctx* c = nullptr;
  std::string tag;
  uint64_t n = 0;
  uint32_t r = 1;
  int bits_owner = 1, bits_pos = 1;
  DevBuf owner_s, pos_s, run_start, grp_start;  // entries sorted by (key, owner, pos)

  // keys: n entries generated in (owner ascending, pos ascending) order; key_bits: significant
  // bits; r: interval length for `covered`.
  void prepare(shn_ctx* ctx, const char* tag_, const uint64_t* d_keys, const uint32_t* d_owner,
               const uint32_t* d_pos, uint64_t n_, int key_bits, uint32_t r_);
  // Pair table of all (j, d) with j in [lo, hi), d < j sharing a key, restricted to partners with
  // d >= lo or d_status[d] == 1 (d_status == nullptr: no restriction).  Sorted by (j, d).
  void join(uint32_t lo, uint32_t hi, const uint8_t* d_status, PairTable* out);
};
