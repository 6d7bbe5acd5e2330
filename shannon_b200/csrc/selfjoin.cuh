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
#include <vector>

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
  uint64_t n = 0, n_owner = 0;
  uint32_t r = 1, owner_base = 0;
  int bits_owner = 1, bits_pos = 1;
  std::vector<uint64_t> h_ent_off;  // first entry of every owner (generation order)
  DevBuf owner_g, pos_g;            // the caller's entries, generation order (owner, pos ascending)
  DevBuf owner_s;                   // owners in sorted order (key, owner, pos)
  DevBuf range_g;                   // per entry (uint2): its partners are the sorted positions [x, y)

  // n entries generated in (owner ascending, pos ascending) order; owners are owner_base ..
  // owner_base + n_owner - 1 and d_ent_off[k] (device, n_owner + 1 values) is the first entry of owner
  // owner_base + k.  key_bits: significant key bits; r: interval length for `covered`.  The three
  // buffers are consumed (keys freed, owner / pos kept by the join).
  void prepare(shn_ctx* ctx, const char* tag_, DevBuf& keys, DevBuf& owner, DevBuf& pos,
               const uint64_t* d_ent_off, uint64_t n_owner_, uint32_t owner_base_, uint64_t n_, int key_bits,
               uint32_t r_);
  // Pair table of all (j, d) with j in [lo, hi), d < j sharing a key, restricted to partners with
  // d >= lo or d_status[d] == 1 (d_status == nullptr: no restriction, lo must be the first owner).
  // Sorted by (j, d).
  void join(uint32_t lo, uint32_t hi, const uint8_t* d_status, PairTable* out);
};
