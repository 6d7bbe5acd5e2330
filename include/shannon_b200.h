/* shannon_b200 -- C ABI of the B200-native k-mer front end of Shannon.
 *
 * The reference (sreeramkannan/Shannon) is pure Python and has NO FFI of its own
 * (SURVEY.md 8b): the drop-in boundary is the two Python entry points
 *     extension_correction.extension_correction(arguments, inMem)   extension_correction.py:528
 *     kmers_for_component.kmers_for_component(...)                  kmers_for_component.py:144
 * which this repo re-implements in shannon_b200/{extension_correction,kmers_for_component}.py
 * on top of the entry points declared here (loaded with ctypes, shannon_b200/_lib.py).
 * Each entry point names the reference code it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message is available from
 *    shn_last_error(ctx) (or shn_last_error(NULL) when no ctx exists yet);
 *  - plain pointers and sizes only; `on_device` != 0 means the data pointers are CUDA device
 *    pointers of ctx's device (inputs already resident in HBM), 0 means host pointers and the
 *    call performs the host<->device copies itself;
 *  - device memory for tables and results is owned by the ctx; one ctx per GPU, used by one
 *    thread at a time; all work is issued on the ctx's own CUDA stream;
 *  - getters are two-phase: call shn_*_sizes first, then pass buffers of at least that size.
 *  - 2-bit base code: A=0, G=1, C=2, T=3 (the reference's successor tie order BASES =
 *    ['A','G','C','T'], extension_correction.py:10; complement = 3 - code).  A K1-mer is
 *    packed with its FIRST base in the most significant used pair:
 *    key = sum_i code[i] << 2*(k1-1-i).  1 <= k1 <= 32: one uint64 per key.  k1 = 33 (K = 32,
 *    the largest -K shannon.py accepts): the key is a 66-bit number carried as TWO uint64 words per
 *    key, low word first (keys[2*i] = bits 0..63, keys[2*i+1] = bits 64..65); every `keys`
 *    array below then holds 2*n words.  The library is compiled once per key width and dispatches
 *    on k1.  The hash-routing calls (shn_route_plan / shn_table_build_indexed) take one-word keys
 *    only.
 */
#ifndef SHANNON_B200_H
#define SHANNON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct shn_ctx shn_ctx;

/* ---- context, memory, timing ------------------------------------------------------- */
int shn_create(int device, shn_ctx** out);
void shn_destroy(shn_ctx* ctx);
const char* shn_last_error(shn_ctx* ctx);
const char* shn_version(void);
int shn_device_info(shn_ctx* ctx, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes);

int shn_dev_alloc(shn_ctx* ctx, uint64_t bytes, void** dptr);
int shn_dev_free(shn_ctx* ctx, void* dptr);
int shn_host_alloc_pinned(shn_ctx* ctx, uint64_t bytes, void** hptr);
int shn_host_free_pinned(shn_ctx* ctx, void* hptr);
int shn_memcpy_h2d(shn_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
int shn_memcpy_d2h(shn_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);
int shn_memcpy_d2d(shn_ctx* ctx, void* dst_dev, const void* src_dev, uint64_t bytes); /* asynchronous */
int shn_sync(shn_ctx* ctx);
/* Issue all further work of ctx on the caller's CUDA stream (cudaStream_t; NULL = back to the
 * context's own stream).  The sharded path runs the library and torch.distributed's NCCL calls on
 * one stream, so no host synchronisation separates kernels from collectives. */
int shn_use_stream(shn_ctx* ctx, void* cuda_stream);
/* CUDA-event stopwatch on the ctx stream (the stream every kernel of this library runs on). */
int shn_timer_start(shn_ctx* ctx);
int shn_timer_stop(shn_ctx* ctx, float* elapsed_ms);
/* Per-kernel device time accounting (CUDA events around each launch group).  enable=1 resets
 * and starts collecting; shn_prof_get returns accumulated ms and launch count for `name`. */
int shn_prof_enable(shn_ctx* ctx, int enable);
int shn_prof_get(shn_ctx* ctx, const char* name, double* total_ms, uint64_t* launches);
int shn_prof_dump(shn_ctx* ctx, char* buf, uint64_t buf_bytes);
uint64_t shn_launch_count(shn_ctx* ctx); /* kernels launched by this ctx so far */
/* Return the cached blocks of the context's device allocator to the driver (the library keeps
 * freed blocks for re-use; a second allocator in the process, e.g. torch's, cannot see them). */
int shn_trim(shn_ctx* ctx);
/* Write `bytes` of scratch larger than L2 to evict cached lines between timed iterations. */
int shn_flush_l2(shn_ctx* ctx);

/* ---- host-side text IO (native, multi-threaded; replaces the per-line Python parsing) -- */
/* Parses a `KMER<TAB|SPACE>count` file (jellyfish dump -c -t; load_kmers,
 * extension_correction.py:209-216: upper-cases, float()-parses integer counts).  Allocates
 * *keys / *counts with malloc (free with shn_host_free).  k1 = length of the first k-mer. */
int shn_parse_kmer_file(shn_ctx* ctx, const char* path, uint64_t** keys, uint32_t** counts,
                        uint64_t* n, int* k1);
void shn_host_free(void* p);
/* Loads the sequence lines of a 2-line-per-record FASTA the way kmers_for_component.py:330-339
 * reads it: name = readline()[:-1], stop if empty; read = readline()[:-1]; a record with an empty
 * read is kept and ends the input.  n_fixed >= 0 reads exactly n_fixed records without the stop
 * rule, padding with empty reads (mate 2 is read in lock-step with mate 1, :372-375).
 * Returns malloc'ed concatenated bases (no separators) and n+1 offsets (shn_host_free). */
int shn_load_fasta(shn_ctx* ctx, const char* path, int64_t n_fixed, char** bases, uint64_t** offsets,
                   uint64_t* n);
/* Writes `>{first_index+e}{suffix}\nSEQ\n` records (kmers_for_component.py:351,396-397) for the
 * reads selected by read_idx[0..m) in that order; append != 0 opens in append mode. */
int shn_write_fasta_subset(shn_ctx* ctx, const char* path, int append, const char* bases,
                           const uint64_t* offsets, const uint32_t* read_idx, uint64_t m,
                           uint64_t first_index, const char* suffix);

/* Writes `KMER<TAB>count` lines (the k1mer.dict_org format of `jellyfish dump -c -t`,
 * shannon.py:441) for n packed keys (host arrays), formatted by all host threads. */
int shn_write_kmer_file(shn_ctx* ctx, const char* path, const uint64_t* keys, const uint32_t* counts,
                        uint64_t n, int k1);

/* component{comp}k1mers_allowed.dict (kmers_for_component.py:457-476): for each listed contig (index
 * into bases/offsets), for each K1-mer window, `K1MER\tweight\n`; the weights of contig c start
 * at weights[win_off[c]]. */
int shn_write_k1mer_windows(shn_ctx* ctx, const char* path, const char* bases,
                            const uint64_t* offsets, const uint32_t* contig_ids, uint64_t m, int k1,
                            const uint32_t* weights, const uint64_t* win_off);

/* ---- a1/a2: packing + K1-mer -> weight table ------------------------------------------ */
/* ASCII K1-mers (n * k1 bytes, no separators, upper or lower case) -> packed keys.
 * Non-ACGT characters are an error (2-bit keys cannot hold them; jellyfish never emits N). */
int shn_pack_kmers(shn_ctx* ctx, const char* ascii, uint64_t n, int k1, uint64_t* keys,
                   int on_device);
/* load_kmers + lowComplexity (extension_correction.py:202-221,142-149).  Entry i of
 * (keys,counts) is input line i.  Low-complexity K1-mers (max base count >= k1-2) are
 * dropped; counts of repeated keys accumulate; with double_stranded each line also adds its
 * count to the reverse complement (inserted right after the forward key: first-occurrence
 * index 2i / 2i+1).  Replaces any previous table of this ctx. */
int shn_table_build(shn_ctx* ctx, const uint64_t* keys, const uint32_t* counts, uint64_t n,
                    int k1, int double_stranded, int on_device);
/* Sharded build: like shn_table_build on device pointers, but entry i carries its own
 * global input-line index (the seed tie-break must see the order of the un-sharded input). */
int shn_table_build_indexed(shn_ctx* ctx, const uint64_t* keys_dev, const uint32_t* counts_dev,
                            const uint32_t* line_idx_dev, uint64_t n, int k1);
/* Owner rank of every key (low 32 bits of fmix64(key), independent of the bucket hash) and the
 * stable partition of the batch by owner: perm_dev[j] = source index of the j-th routed element,
 * counts_host[r] = elements owned by rank r.  Feeds an all-to-all whose send buffer must be
 * contiguous per destination. */
int shn_route_plan(shn_ctx* ctx, const uint64_t* keys_dev, uint64_t n, uint32_t nranks,
                   uint32_t* perm_dev, uint64_t* counts_host);
/* dst[i] = src[perm[i]] (scatter == 0) or dst[perm[i]] = src[i] (scatter != 0); elem_bytes in
 * {1,4,8}; all device pointers. */
int shn_permute(shn_ctx* ctx, const void* src_dev, const uint32_t* perm_dev, uint64_t n,
                int elem_bytes, int scatter, void* dst_dev);
int shn_table_stats(shn_ctx* ctx, uint64_t* n_distinct, uint64_t* n_lowcomplexity,
                    uint64_t* n_slots, int* k1);
/* `kmer in kmers` / `kmers[kmer]`: weight (0 if absent) and found flag (0/1) per query. */
int shn_table_lookup(shn_ctx* ctx, const uint64_t* keys, uint64_t n, uint32_t* weights,
                     uint8_t* found, int on_device);
/* All (key, weight, first-occurrence index) triples, sorted by first-occurrence index
 * (= insertion order of the reference's dict).  Buffers sized n_distinct. */
int shn_table_dump(shn_ctx* ctx, uint64_t* keys, uint32_t* weights, uint32_t* first_idx);

/* ---- a3-a9: seeds, greedy walks, shape filter, duplicate filter, contig graph ------------ */
/* run_correction's seed loop and accept logic (extension_correction.py:334-397) on the table
 * built by shn_table_build.  After it returns, the getters below are valid.  The table itself is
 * left exactly as it was (keys and weights are never written; the first-occurrence words are
 * borrowed by the walks and restored before the call returns, also when it fails), so lookups,
 * shn_table_dump and further shn_l3_run calls see the same table. */
int shn_l3_run(shn_ctx* ctx, uint32_t min_weight, uint32_t min_length);
typedef struct shn_l3_sizes {
  uint64_t n_seeds;        /* K1-mers with weight >= min_weight */
  uint64_t n_raw_comps;    /* connected components of the K1-mer successor graph */
  uint64_t n_walks;        /* walks started (seed not yet traversed), in pop order */
  uint64_t n_traversed;    /* K1-mers traversed by all walks */
  uint64_t n_candidates;   /* walks passing length + hyperbola terms */
  uint64_t n_contigs;      /* accepted contigs */
  uint64_t contig_bases;   /* total bases of accepted contigs */
  uint64_t n_allowed;      /* K1-mers of accepted contigs */
  uint64_t n_edges;        /* distinct undirected contig-contig edges */
  uint64_t dup_rounds;     /* frontier rounds the duplicate filter needed */
  uint64_t walk_rounds;    /* longest per-component serial chain (2-step probe rounds) */
  uint64_t n_spec_comps;   /* components walked with speculative windows */
  uint64_t spec_windows;   /* windows those components needed */
} shn_l3_sizes;
int shn_l3_get_sizes(shn_ctx* ctx, shn_l3_sizes* out);
/* The same in two phases, for the sharded path:
 *   shn_l3_walks  = seeds, raw K1-mer components, greedy walks, length + hyperbola terms, candidate
 *                   contigs (extension_correction.py:334-356) -- needs only this context's table,
 *                   which must hold COMPLETE components of the K1-mer successor graph;
 *   shn_l3_filter = duplicate_check in acceptance order, allowed set, contig C-mer graph and its
 *                   components (:358-450) on this context's own candidates (external == 0) or on
 *                   candidates merged from all ranks in global pop order (external != 0: device
 *                   pointers, one base code per byte, n_cand+1 offsets).  allow_missing != 0:
 *                   allowed K1-mers absent from this context's table get weight 0 (their owner rank
 *                   supplies it: shn_l3_allowed_copy / all-reduce / shn_l3_set_allowed_weights).
 * shn_l3_run == shn_l3_walks + shn_l3_filter(NULL, NULL, 0, 0, 0). */
int shn_l3_walks(shn_ctx* ctx, uint32_t min_weight, uint32_t min_length);
int shn_l3_cand_sizes(shn_ctx* ctx, uint64_t* n_candidates, uint64_t* n_bases);
/* candidates of shn_l3_walks in pop order: weight and (global) input line of the seed (the pop
 * order key: weight descending, line descending), offsets, base codes.  Device pointers. */
int shn_l3_cand_export(shn_ctx* ctx, uint32_t* seed_weight_dev, uint64_t* seed_line_dev,
                       uint64_t* offsets_dev, uint8_t* codes_dev);
int shn_l3_filter(shn_ctx* ctx, const uint8_t* codes_dev, const uint64_t* offsets_dev, uint64_t n_cand,
                  int external, int allow_missing);
int shn_l3_allowed_copy(shn_ctx* ctx, uint64_t* keys_dev, uint32_t* weights_dev);
int shn_l3_set_allowed_weights(shn_ctx* ctx, const uint32_t* weights_dev);
/* Per started walk, in pop order: seed key, steps to the left/right, sum of weights, and flags
 * bit0 = passes length+hyperbola, bit1 = duplicate_check() true (only evaluated when bit0),
 * bit2 = accepted. */
int shn_l3_get_walks(shn_ctx* ctx, uint64_t* seed_keys, uint32_t* n_left, uint32_t* n_right,
                     uint64_t* tot_wt, uint8_t* flags);
/* Accepted contigs in acceptance order (contig index 1..n): ASCII bases concatenated,
 * n_contigs+1 offsets. */
int shn_l3_get_contigs(shn_ctx* ctx, char* bases, uint64_t* offsets);
/* allowed K1-mers in contig order with int weights (allowed_kmer_dict, :404-408). */
int shn_l3_get_allowed(shn_ctx* ctx, uint64_t* keys, uint32_t* weights);
/* contig_connections (:372-389) as distinct undirected edges a<b (1-based contig indices) with
 * multiplicity weight and, for the insertion order of b's neighbour dict, the first C-mer
 * position in b shared with a.  Sorted by (b,a). */
int shn_l3_get_edges(shn_ctx* ctx, uint32_t* a, uint32_t* b, uint32_t* weight,
                     uint32_t* first_pos_in_b);
/* component label per contig (index 0 unused) = minimum contig index of its component
 * (GPU union-find; the DFS of :417-434 yields the same partition). */
int shn_l3_get_labels(shn_ctx* ctx, uint32_t* label);

/* ---- a10-a12: K1-mer -> component map, read partition ------------------------------------ */
/* k1mers2component build (kmers_for_component.py:239-305) for contigs given as ASCII
 * (concatenated, n+1 offsets) with one component id per contig.  Every K1-mer window of contig
 * c gets comp_of_contig[c] added to its component set (at most 2 distinct ids per K1-mer: the
 * 'c' partition and its 'r2_c' twin; more is reported as an error) and the weight
 * k1mer_dictionary.get(k1mer, 0) is taken from (dict_keys, dict_weights) when given, or from
 * the L3 allowed set of this ctx when dict_keys == NULL.  Call with reset != 0 first. */
int shn_l4_map_add_contigs(shn_ctx* ctx, const char* bases, const uint64_t* offsets,
                           const uint32_t* comp_of_contig, uint64_t n_contigs, int k1, int reset,
                           uint64_t expected_total_k1mers);
/* Same for the accepted contigs of this ctx's last shn_l3_run, which are still on the device:
 * comp_of_contig[i] is the component of contig i+1 (acceptance order), 0xFFFFFFFF = the contig is
 * not partitioned (single-contig components, extension_correction.py:467-473). */
int shn_l4_map_add_l3_contigs(shn_ctx* ctx, const uint32_t* comp_of_contig, uint64_t n_contigs,
                              int reset);
int shn_l4_map_set_weights(shn_ctx* ctx, const uint64_t* dict_keys, const uint32_t* dict_weights,
                           uint64_t n);
/* weight per K1-mer window of the given contigs, in order (component*k1mers_allowed.dict). */
int shn_l4_map_window_weights(shn_ctx* ctx, const char* bases, const uint64_t* offsets,
                              uint64_t n_contigs, int k1, uint32_t* weights);
/* 2-bit packing of reads: ASCII bases + offsets -> device-resident packed reads held by ctx.
 * A read containing any character outside "ACGT" is flagged invalid (kmers_for_component.py
 * :336,376).  mate = 0 or 1. */
int shn_l4_load_reads(shn_ctx* ctx, int mate, const char* bases, const uint64_t* offsets,
                      uint64_t n_reads, int on_device);
/* Same in two halves, so that the host->device copy overlaps earlier stages: _async starts the
 * copy of the host buffers (pinned memory for real overlap; they must stay valid until _staged
 * returns) on the context's copy stream and returns; _staged waits for it and packs. */
int shn_l4_upload_reads_async(shn_ctx* ctx, int mate, const char* bases, const uint64_t* offsets,
                              uint64_t n_reads);
int shn_l4_load_reads_staged(shn_ctx* ctx, int mate);
/* get_rmers/get_comps/get_comps_paired + the chunk loop (kmers_for_component.py:186-205,
 * 322-423): samples K1-mers at offsets 0,K1,2*K1,... (< len-K1) plus the last K1-mer of every
 * read (both mates when paired), takes the UNION of the component ids hit, and groups the
 * record indices by component preserving input order. */
int shn_l4_assign(shn_ctx* ctx, int paired, int k1, uint64_t* n_assignments, uint64_t* n_lookups,
                  uint64_t* n_valid_records);
/* comp_offsets: n_comps+1 prefix offsets into record_idx (record indices ascending per comp). */
int shn_l4_get_assignments(shn_ctx* ctx, uint32_t n_comps, uint64_t* comp_offsets,
                           uint32_t* record_idx);

/* the same into device buffers (n_comps+1 offsets, n_assignments indices), record indices shifted
 * by first_record: the lists of the ranks of the sharded path are merged on the device. */
int shn_l4_assignments_dev(shn_ctx* ctx, uint32_t n_comps, uint64_t first_record,
                           uint64_t* comp_offsets_dev, uint32_t* record_idx_dev);

/* ---- e: the path on hash-sharded tables, one shard per rank (shannon_b200/dist.py drives the
 * exchanges with torch.distributed all_to_all_single on the same stream) -------------------------
 * Records on the wire have the size of a table slot (16 bytes for k1 <= 32, 32 bytes for k1 = 33):
 * {key (1 or 2 words), payload u64 [, pad u64]}.  Lines and table entries carry payload =
 * global input line << 30 | weight, successor queries the asking component.  owner(K1-mer) = hash
 * of its minimizer (the 11-mer with the smallest hash), so most successor edges are rank-local.
 * Every routing call is made twice: with send_dev == NULL it writes the number of records per
 * destination rank to counts_host; with a send buffer of sum(counts) records it reads counts_host
 * back and fills the buffer contiguously per destination (order inside a destination is free: the
 * records carry their global line).  All pointers are device pointers unless named *_host. */
/* load_kmers' lines (extension_correction.py:209-219) of this rank's slice of k1mer.dict_org:
 * line i is global line first_line + i; double_stranded emits the reverse complement as line 2i+1. */
int shn_route_lines(shn_ctx* ctx, const uint64_t* keys_dev, const uint32_t* counts_dev, uint64_t n,
                    uint64_t first_line, int double_stranded, int k1, uint32_t nranks,
                    uint64_t* counts_host, void* send_dev);
/* shn_table_build from received records.  The context remembers the global input line of every
 * record: the seed order of shn_l3_walks (weight descending, LATER global line first) and
 * shn_cc_route read it; a key received on several lines keeps the smallest one. */
int shn_table_build_records(shn_ctx* ctx, const void* recs_dev, uint64_t n, int k1);
/* connected components of the successor graph restricted to this shard (lock-free union-find). */
int shn_cc_local(shn_ctx* ctx, uint64_t* n_local_components);
/* successor candidates owned by other ranks: {successor key, gid_base + my local component}. */
int shn_cc_cross(shn_ctx* ctx, uint32_t nranks, uint32_t rank, uint64_t gid_base, uint64_t* counts_host,
                 void* send_dev);
/* received queries -> edges (asking component | my component << 32) for the keys this shard holds;
 * edges_dev has room for n entries. */
int shn_cc_resolve(shn_ctx* ctx, const void* recs_dev, uint64_t n, uint64_t gid_base, uint64_t* edges_dev,
                   uint64_t* n_edges);
/* components of the graph whose nodes are the local components of all ranks (all-gathered edges). */
int shn_cc_merge(shn_ctx* ctx, const uint64_t* edges_dev, uint64_t n_edges, uint64_t n_super,
                 uint64_t* n_final_components);
/* sizes_dev[f] = K1-mers of this shard in final component f (n_final entries, zeroed here). */
int shn_cc_sizes(shn_ctx* ctx, uint64_t gid_base, uint64_t* sizes_dev);
/* every table entry {key, global line << 30 | weight} to owner_of_final[its component]. */
int shn_cc_route(shn_ctx* ctx, const uint32_t* owner_of_final_dev, uint64_t gid_base, uint32_t nranks,
                 uint64_t* counts_host, void* send_dev);
int shn_cc_free(shn_ctx* ctx);

/* ---- f1/f2 (inputs of the path): RC doubling, K1-mer counting, synthetic reads ---------- */
/* Generates n_pairs synthetic read pairs on the device (twin of shannon_b200/synth.py::
 * make_pairs) as two ASCII arrays of n_pairs*read_len bytes (device pointers, caller-allocated). */
int shn_synth_pairs(shn_ctx* ctx, const uint8_t* tx_codes, const uint64_t* tx_offs,
                    const uint64_t* thresholds, uint64_t n_tx, uint64_t n_pairs,
                    uint64_t first_pair, uint64_t seed, int read_len, int frag_len,
                    uint32_t err_threshold_24, char* mate1_dev, char* mate2_dev);
/* out[i] = reverse complement of fixed-length read i (rc_gnu.py / rc_s.py; N stays N). */
int shn_revcomp_reads(shn_ctx* ctx, const char* in_dev, char* out_dev, uint64_t n_reads,
                      int read_len);
/* jellyfish count -m k1 + dump -L 1 stand-in (shannon.py:439-441): counts every k1-mer window
 * without non-ACGT characters of the given fixed-length read arrays; results (sorted by the
 * ASCII order of the k-mer, the order of oracle/kmer_count.py) stay on the device and are
 * returned as device pointers owned by the ctx. */
int shn_count_k1mers(shn_ctx* ctx, const char* const* read_arrays_dev, const uint64_t* n_reads,
                     int n_arrays, int read_len, int k1, uint64_t expected_distinct,
                     uint64_t** keys_dev, uint32_t** counts_dev, uint64_t* n_distinct);

/* Variable-length twins for real read files (rows f1/f2 of SURVEY 8f):
 * shn_revcomp_var: out = reverse complement of every read (rc_s.py: A<->T, C<->G, N stays; any other
 * character is an error, as in rc_s.py); same offsets for input and output.
 * shn_count_begin / shn_count_add_reads / shn_count_finish: `jellyfish count -m k1` + `dump -c -t -L
 * min_count` (shannon.py:439-441) over reads added in chunks (only one chunk has to be resident);
 * windows with a character outside ACGT are skipped; result in ascending ASCII order of the k-mer,
 * device arrays owned by the context -- feed them to shn_table_build(on_device = 1) or to
 * shn_write_kmer_file after a copy to the host. */
int shn_revcomp_var(shn_ctx* ctx, const char* bases, const uint64_t* offsets, uint64_t n_reads, char* out,
                    int on_device);
int shn_count_begin(shn_ctx* ctx, int k1, uint64_t expected_distinct);
int shn_count_add_reads(shn_ctx* ctx, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                        int on_device);
int shn_count_finish(shn_ctx* ctx, uint32_t min_count, uint64_t** keys_dev, uint32_t** counts_dev,
                     uint64_t* n_distinct);
/* FASTA with names, as rc_s.py reads it (blank lines dropped, header = stripped line starting with
 * '>', sequence = first field of the next line); malloc'ed arrays (shn_host_free), names without '>'. */
int shn_load_fasta_named(shn_ctx* ctx, const char* path, char** names, uint64_t** name_offsets,
                         char** bases, uint64_t** offsets, uint64_t* n);
/* `>name\nSEQ\n` records (the format rc_s.py writes). */
int shn_write_fasta_named(shn_ctx* ctx, const char* path, int append, const char* names,
                          const uint64_t* name_offsets, const char* bases, const uint64_t* offsets,
                          uint64_t n);
/* ---- f3: multibridging.load_single_jellyfish + Node.condense_all (multibridging.py:145-172,
 * mbgraph.py:479-507,184-258), the first step of the consumer of the per-component k1mer.dict ------
 * Line i of the file is the edge prefix_kmers[i] -> suffix_kmers[i] (the K-mer prefix / suffix of the
 * K1-mer, packed like keys, K <= 32) with prevalence[i].  Nodes are numbered in first-appearance
 * order; every unambiguous edge is condensed, so the result is the unitig graph: per unitig its
 * bases, count (= norm, the number of K-mers) and prevalence (sum over its K-mers of (K-1) x
 * out-degree); per remaining edge (source unitig, destination unitig, copy count), weight K-1; the
 * copy count of an edge touching a condensed node is 0, as in the reference.  Pure cycles of
 * unambiguous edges (order dependent in the reference) are left uncondensed and counted.
 * Host pointers; two-phase: run, then get with buffers of the returned sizes. */
int shn_condense_run(shn_ctx* ctx, const uint64_t* prefix_kmers, const uint64_t* suffix_kmers,
                     const uint32_t* prevalence, uint64_t n, int K, uint64_t* n_unitigs, uint64_t* n_bases,
                     uint64_t* n_edges, uint64_t* n_cycle_nodes);
int shn_condense_get(shn_ctx* ctx, char* bases, uint64_t* offsets, uint32_t* count, uint64_t* prevalence,
                     uint32_t* edge_src, uint32_t* edge_dst, uint32_t* edge_copy_count);

/* ---- f4: faster_reps.py:60-131 (representative selection among the final transcripts) ----------
 * duplicate_out[c] = 1 iff find_reps would drop transcript c: its first and last 24-mer (of the
 * transcript, or of its reverse complement when double_stranded) both occur in one other transcript
 * at a distance within 2 of its own length, and that transcript is longer, or equally long with a
 * smaller name (name_rank[c] = rank of the name in string order).  Host pointers; ACGT only. */
int shn_find_reps(shn_ctx* ctx, const char* bases, const uint64_t* offsets, const uint32_t* name_rank,
                  uint64_t n, int double_stranded, uint8_t* duplicate_out);
/* frees the device arrays the last shn_count_k1mers returned */
int shn_count_release(shn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SHANNON_B200_H */
