// Packed reads of a context: independent of the K1-mer key width (reads.cu, compiled once).
#pragma once
#include "common.cuh"

struct PackedReads {
  DevBuf words;   // uint64, 32 bases per word, MSB first
  DevBuf woff;    // uint64 [n+1] word offsets
  DevBuf len;     // uint32 [n]   length, bit 31 = contains a character outside ACGT
  uint64_t n = 0;
  uint64_t n_words = 0;
};

struct ReadsState {
  PackedReads reads[2];
  DevBuf stage_a, stage_b;
  // reads uploaded ahead of time on the copy stream (overlaps the H2D with the L3 stage)
  DevBuf up_bases[2], up_offs[2];
  uint64_t up_n[2] = {0, 0};
  cudaEvent_t up_done[2] = {nullptr, nullptr};
};

ReadsState* shn_reads_of(shn_ctx* c);
void shn_reads_load(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets, uint64_t n,
                    int on_device);
void shn_reads_upload_async(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets, uint64_t n);
void shn_reads_load_staged(shn_ctx* c, int mate);
