// Entry points of the key-width specific translation units (table.cu, l3.cu, l4.cu, synth.cu's
// counter), declared once per namespace with key arrays as plain uint64 words; api.cu dispatches
// on k1 (<= 32: narrow, 33..64: wide).
#pragma once
#include "common.cuh"

#define SHN_DECLARE_IMPLS(NS)                                                                        \
  namespace NS {                                                                                     \
  void pack_kmers(shn_ctx* c, const char* d_ascii, uint64_t n, int k1, uint64_t* d_keys);            \
  void table_begin(shn_ctx* c, uint64_t n, int k1, int double_stranded);                             \
  void table_insert_chunk(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts,              \
                          const uint32_t* d_line_idx, uint64_t n, uint64_t first_line,               \
                          int double_stranded);                                                      \
  void table_finish(shn_ctx* c);                                                                     \
  void table_build(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n, int k1, \
                   int double_stranded, const uint32_t* d_line_idx);                                 \
  void table_lookup(shn_ctx* c, const uint64_t* d_keys, uint64_t n, uint32_t* d_weights,             \
                    uint8_t* d_found);                                                               \
  void table_dump(shn_ctx* c, uint64_t* h_keys, uint32_t* h_weights, uint32_t* h_idx);               \
  void l3_run(shn_ctx* c, uint32_t min_weight, uint32_t min_length);                                 \
  void l3_walks(shn_ctx* c, uint32_t min_weight, uint32_t min_length);                               \
  void l3_filter(shn_ctx* c, const uint8_t* ext_codes, const uint64_t* ext_offs, uint64_t ext_n,     \
                 int ext, int allow_missing);                                                        \
  void l3_cand_sizes(shn_ctx* c, uint64_t* n_cand, uint64_t* n_bases);                               \
  void l3_cand_export(shn_ctx* c, uint32_t* d_weight, uint64_t* d_line, uint64_t* d_offs,            \
                      uint8_t* d_codes);                                                             \
  void l3_set_allowed_weights(shn_ctx* c, const uint32_t* d_w);                                      \
  void route_lines(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n,         \
                   uint64_t first_line, int ds, int k1, uint32_t nranks, uint64_t* h_counts,         \
                   void* send);                                                                      \
  void table_build_records(shn_ctx* c, const void* d_recs, uint64_t n, int k1);                      \
  void cc_local(shn_ctx* c, uint64_t* n_local);                                                      \
  void cc_cross(shn_ctx* c, uint32_t nranks, uint32_t rank, uint64_t gid_base, uint64_t* h_counts,   \
                void* send);                                                                         \
  void cc_resolve(shn_ctx* c, const void* d_recs, uint64_t n, uint64_t gid_base, uint64_t* d_edges,  \
                  uint64_t* n_edges);                                                                \
  void cc_merge(shn_ctx* c, const uint64_t* d_edges, uint64_t n_edges, uint64_t n_super,             \
                uint64_t* n_final);                                                                  \
  void cc_sizes(shn_ctx* c, uint64_t gid_base, uint64_t* d_sizes);                                   \
  void cc_route(shn_ctx* c, const uint32_t* d_owner_of_final, uint64_t gid_base, uint32_t nranks,    \
                uint64_t* h_counts, void* send);                                                     \
  void cc_free(shn_ctx* c);                                                                          \
  void l3_get_sizes(shn_ctx* c, shn_l3_sizes* out);                                                  \
  void l3_get_walks(shn_ctx* c, uint64_t* seed_keys, uint32_t* n_left, uint32_t* n_right,            \
                    uint64_t* tot_wt, uint8_t* flags);                                               \
  void l3_get_contigs(shn_ctx* c, char* bases, uint64_t* offsets);                                   \
  void l3_get_allowed(shn_ctx* c, uint64_t* keys, uint32_t* weights);                                \
  void l3_get_edges(shn_ctx* c, uint32_t* a, uint32_t* b, uint32_t* weight, uint32_t* fp);           \
  void l3_get_labels(shn_ctx* c, uint32_t* label);                                                   \
  void l3_allowed_dev(shn_ctx* c, const uint64_t** keys, const uint32_t** weights, uint64_t* n);     \
  void l3_contigs_dev(shn_ctx* c, const uint8_t** codes, const uint64_t** offs, uint64_t* n,         \
                      uint64_t* n_allowed);                                                          \
  void l4_map_add_contigs(shn_ctx* c, const char* bases, const uint64_t* offsets,                    \
                          const uint32_t* comp_of_contig, uint64_t n_contigs, int k1, int reset,     \
                          uint64_t expected_total, int on_device, int is_codes);                     \
  void l4_map_set_weights(shn_ctx* c, const uint64_t* keys, const uint32_t* weights, uint64_t n,     \
                          int on_device);                                                            \
  void l4_map_window_weights(shn_ctx* c, const char* bases, const uint64_t* offsets,                 \
                             uint64_t n_contigs, int k1, uint32_t* h_weights);                       \
  void l4_assign(shn_ctx* c, int paired, int k1, uint64_t* n_assign, uint64_t* n_lookups,            \
                 uint64_t* n_valid);                                                                 \
  void l4_get_assignments(shn_ctx* c, uint32_t n_comps, uint64_t* h_offs, uint32_t* h_idx);          \
  void l4_assignments_dev(shn_ctx* c, uint32_t n_comps, uint64_t first_record, uint64_t* d_offs,     \
                          uint32_t* d_idx);                                                          \
  void count_begin(shn_ctx* c, int k1, uint64_t expected_distinct);                                  \
  void count_add_var(shn_ctx* c, const char* d_bases, const uint64_t* d_offs, uint64_t n_reads,      \
                     uint64_t total_bases);                                                          \
  void count_finish(shn_ctx* c, uint32_t min_count, uint64_t** keys_dev, uint32_t** counts_dev,      \
                    uint64_t* n_distinct);                                                           \
  void count_k1mers(shn_ctx* c, const char* const* arrays, const uint64_t* n_reads, int n_arrays,    \
                    int read_len, int k1, uint64_t expected_distinct, uint64_t** keys_dev,           \
                    uint32_t** counts_dev, uint64_t* n_distinct);                                    \
  }

SHN_DECLARE_IMPLS(narrow)
SHN_DECLARE_IMPLS(wide)
