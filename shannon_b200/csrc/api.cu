// extern "C" entry points of libshannon_b200.so (see include/shannon_b200.h).
#include "common.cuh"
#include "selfjoin.cuh"

// implemented in the other translation units
#include "impls.h"
#include "reads.cuh"
void shn_parse_kmer_file_impl(const char* path, uint64_t** keys_out, uint32_t** counts_out,
                              uint64_t* n_out, int* k1_out);
void shn_load_fasta_impl(const char* path, int64_t n_fixed, char** bases_out, uint64_t** offs_out,
                         uint64_t* n_out);
void shn_write_fasta_subset_impl(const char* path, int append, const char* bases,
                                 const uint64_t* offsets, const uint32_t* read_idx, uint64_t m,
                                 uint64_t first_index, const char* suffix);
void shn_write_k1mer_windows_impl(const char* path, const char* bases, const uint64_t* offsets,
                                  const uint32_t* contig_ids, uint64_t m, int k1,
                                  const uint32_t* weights, const uint64_t* win_off);
void shn_write_kmer_file_impl(const char* path, const uint64_t* keys, const uint32_t* counts, uint64_t n,
                              int k1);
void shn_condense_run_impl(shn_ctx* c, const uint64_t* h_pre, const uint64_t* h_suf, const uint32_t* h_prev,
                           uint64_t n, int K, uint64_t* n_unitigs, uint64_t* n_bases, uint64_t* n_edges,
                           uint64_t* n_cycle_nodes);
void shn_condense_get_impl(shn_ctx* c, char* bases, uint64_t* offsets, uint32_t* count, uint64_t* prevalence,
                           uint32_t* e_src, uint32_t* e_dst, uint32_t* e_cc);
void shn_find_reps_impl(shn_ctx* c, const char* bases, const uint64_t* offsets, const uint32_t* name_rank,
                        uint64_t n, int ds, uint8_t* dup_out);
void shn_synth_pairs_impl(shn_ctx* c, const uint8_t* tx, const uint64_t* tx_offs, const uint64_t* thr,
                          uint64_t n_tx, uint64_t n_pairs, uint64_t first_pair, uint64_t seed,
                          int read_len, int frag_len, uint32_t err_thr, char* m1, char* m2);
void shn_revcomp_reads_impl(shn_ctx* c, const char* in, char* out, uint64_t n_reads, int read_len);
void shn_revcomp_var_impl(shn_ctx* c, const char* in, const uint64_t* offs, uint64_t n_reads, uint64_t total,
                          char* out);
void shn_load_fasta_named_impl(const char* path, char** names_out, uint64_t** name_offs_out, char** bases_out,
                               uint64_t** offs_out, uint64_t* n_out);
void shn_write_fasta_named_impl(const char* path, int append, const char* names, const uint64_t* name_offs,
                                const char* bases, const uint64_t* offs, uint64_t n);
void shn_route_plan_impl(shn_ctx* c, const uint64_t* d_keys, uint64_t n, uint32_t nranks,
                         uint32_t* d_perm, uint64_t* h_counts);
void shn_permute_impl(shn_ctx* c, const void* src, const uint32_t* d_perm, uint64_t n, int elem_bytes,
                      int scatter, void* dst);

// K1 <= 32: 64-bit keys (namespace narrow); K1 = 33: 128-bit keys (namespace wide).  Key arrays
// cross this ABI as 1 resp. 2 uint64 words per key, low word first.
static inline bool is_wide(int k1) { return k1 > 32; }
static inline uint64_t key_bytes(int k1) { return is_wide(k1) ? 16 : 8; }
#define SHN_DISPATCH(k1, call) \
  do {                         \
    if (is_wide(k1))           \
      wide::call;              \
    else                       \
      narrow::call;            \
  } while (0)

static thread_local std::string g_last_error;
thread_local DevPool* g_shn_pool = nullptr;

[[noreturn]] void shn_throw(const char* file, int line, const std::string& msg) {
  const char* base = strrchr(file, '/');
  throw ShnError(std::string(base ? base + 1 : file) + ":" + std::to_string(line) + ": " + msg);
}

#define SHN_API_BEGIN try {
#define SHN_API_END(ctx)                                   \
  return 0;                                                \
  }                                                        \
  catch (const std::exception& e) {                        \
    g_last_error = e.what();                               \
    if (ctx) (ctx)->last_error = e.what();                 \
    return 1;                                              \
  }                                                        \
  catch (...) {                                            \
    g_last_error = "unknown error";                        \
    if (ctx) (ctx)->last_error = g_last_error;             \
    return 1;                                              \
  }

static void bind(shn_ctx* c) {
  SHN_CHECK(c != nullptr, "null context");
  CUDA_CHECK(cudaSetDevice(c->device));
  g_shn_pool = &c->pool;
}

extern "C" {

const char* shn_version(void) { return "shannon_b200 0.1 (sm_100a)"; }

const char* shn_last_error(shn_ctx* ctx) {
  if (ctx) return ctx->last_error.c_str();
  return g_last_error.c_str();
}

int shn_create(int device, shn_ctx** out) {
  shn_ctx* c = nullptr;
  SHN_API_BEGIN
  SHN_CHECK(out != nullptr, "null out pointer");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    SHN_FAIL(std::string("no CUDA device available (") + cudaGetErrorString(e) +
             "); shannon_b200 has no CPU fallback");
  SHN_CHECK(device >= 0 && device < n_dev, "device index out of range");
  CUDA_CHECK(cudaSetDevice(device));
  c = new shn_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = c->stream;
  CUDA_CHECK(cudaEventCreate(&c->t0));
  CUDA_CHECK(cudaEventCreate(&c->t1));
  CUDA_CHECK(cudaEventCreate(&c->p0));
  CUDA_CHECK(cudaEventCreate(&c->p1));
  *out = c;
  SHN_API_END(c)
}

void shn_destroy(shn_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  g_shn_pool = &c->pool;
  cudaStreamSynchronize(c->stream);
  shn_l3_free(c);
  shn_l4_free(c);
  shn_count_free(c);
  shn_reads_free(c);
  shn_shard_free(c);
  if (c->condense && c->condense_free) c->condense_free(c);
  c->gline_buf.release();
  c->table.release();
  c->cub_tmp.release();
  c->flush_buf.release();
  c->counters.release();
  c->prof_resolve();
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->t0);
  cudaEventDestroy(c->t1);
  cudaEventDestroy(c->p0);
  cudaEventDestroy(c->p1);
  cudaStreamSynchronize(c->stream);
  c->pool.trim();
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  if (c->stream4) cudaStreamDestroy(c->stream4);
  if (c->stream5) cudaStreamDestroy(c->stream5);
  cudaStreamDestroy(c->own_stream);
  g_shn_pool = nullptr;
  delete c;
}

int shn_device_info(shn_ctx* c, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes) {
  SHN_API_BEGIN
  bind(c);
  size_t f = 0, t = 0;
  CUDA_CHECK(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = c->sm_count;
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  SHN_API_END(c)
}

int shn_dev_alloc(shn_ctx* c, uint64_t bytes, void** dptr) {
  SHN_API_BEGIN
  bind(c);
  *dptr = nullptr;
  if (bytes) {
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess)
      SHN_FAIL("cudaMalloc(" + std::to_string(bytes) + ") failed: " + cudaGetErrorString(e));
  }
  SHN_API_END(c)
}
int shn_dev_free(shn_ctx* c, void* dptr) {
  SHN_API_BEGIN
  bind(c);
  if (dptr) CUDA_CHECK(cudaFree(dptr));
  SHN_API_END(c)
}
int shn_host_alloc_pinned(shn_ctx* c, uint64_t bytes, void** hptr) {
  SHN_API_BEGIN
  bind(c);
  *hptr = nullptr;
  if (bytes) CUDA_CHECK(cudaMallocHost(hptr, bytes));
  SHN_API_END(c)
}
int shn_host_free_pinned(shn_ctx* c, void* hptr) {
  SHN_API_BEGIN
  bind(c);
  if (hptr) CUDA_CHECK(cudaFreeHost(hptr));
  SHN_API_END(c)
}
int shn_memcpy_h2d(shn_ctx* c, void* dst, const void* src, uint64_t bytes) {
  SHN_API_BEGIN
  bind(c);
  if (bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  SHN_API_END(c)
}
int shn_memcpy_d2h(shn_ctx* c, void* dst, const void* src, uint64_t bytes) {
  SHN_API_BEGIN
  bind(c);
  if (bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  SHN_API_END(c)
}
int shn_sync(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_API_END(c)
}
int shn_use_stream(shn_ctx* c, void* stream) {
  SHN_API_BEGIN
  bind(c);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));  // nothing of the old stream is left in flight
  c->prof_resolve();
  c->stream = stream ? static_cast<cudaStream_t>(stream) : c->own_stream;
  SHN_API_END(c)
}
int shn_memcpy_d2d(shn_ctx* c, void* dst, const void* src, uint64_t bytes) {
  SHN_API_BEGIN
  bind(c);
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  SHN_API_END(c)
}
int shn_timer_start(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  CUDA_CHECK(cudaEventRecord(c->t0, c->stream));
  SHN_API_END(c)
}
int shn_timer_stop(shn_ctx* c, float* ms) {
  SHN_API_BEGIN
  bind(c);
  CUDA_CHECK(cudaEventRecord(c->t1, c->stream));
  CUDA_CHECK(cudaEventSynchronize(c->t1));
  CUDA_CHECK(cudaEventElapsedTime(ms, c->t0, c->t1));
  SHN_API_END(c)
}
int shn_prof_enable(shn_ctx* c, int enable) {
  SHN_API_BEGIN
  bind(c);
  c->prof_resolve();
  c->prof_on = enable != 0;
  if (enable) c->prof.clear();
  SHN_API_END(c)
}
int shn_prof_get(shn_ctx* c, const char* name, double* total_ms, uint64_t* launches) {
  SHN_API_BEGIN
  SHN_CHECK(c != nullptr, "null context");
  c->prof_resolve();
  auto it = c->prof.find(name);
  *total_ms = it == c->prof.end() ? 0.0 : it->second.ms;
  *launches = it == c->prof.end() ? 0 : it->second.launches;
  SHN_API_END(c)
}
int shn_prof_dump(shn_ctx* c, char* buf, uint64_t buf_bytes) {
  SHN_API_BEGIN
  SHN_CHECK(c != nullptr && buf != nullptr && buf_bytes > 0, "bad arguments");
  c->prof_resolve();
  std::string s;
  for (auto& kv : c->prof)
    s += kv.first + "\t" + std::to_string(kv.second.ms) + "\t" + std::to_string(kv.second.launches) + "\n";
  if (s.size() + 1 > buf_bytes) s.resize(buf_bytes - 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  SHN_API_END(c)
}
uint64_t shn_launch_count(shn_ctx* c) { return c ? c->launches : 0; }

int shn_trim(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  c->pool.trim();
  SHN_API_END(c)
}
int shn_flush_l2(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  const uint64_t bytes = 256ull << 20;  // > 126 MB L2
  c->flush_buf.reserve(bytes);
  CUDA_CHECK(cudaMemsetAsync(c->flush_buf.p, 0x5A, bytes, c->stream));
  SHN_API_END(c)
}

// ---- host IO -----------------------------------------------------------------------------------
int shn_parse_kmer_file(shn_ctx* c, const char* path, uint64_t** keys, uint32_t** counts, uint64_t* n,
                        int* k1) {
  SHN_API_BEGIN
  shn_parse_kmer_file_impl(path, keys, counts, n, k1);
  SHN_API_END(c)
}
void shn_host_free(void* p) { free(p); }
int shn_load_fasta(shn_ctx* c, const char* path, int64_t n_fixed, char** bases, uint64_t** offsets,
                   uint64_t* n) {
  SHN_API_BEGIN
  shn_load_fasta_impl(path, n_fixed, bases, offsets, n);
  SHN_API_END(c)
}
int shn_write_fasta_subset(shn_ctx* c, const char* path, int append, const char* bases,
                           const uint64_t* offsets, const uint32_t* read_idx, uint64_t m,
                           uint64_t first_index, const char* suffix) {
  SHN_API_BEGIN
  shn_write_fasta_subset_impl(path, append, bases, offsets, read_idx, m, first_index, suffix);
  SHN_API_END(c)
}
int shn_write_k1mer_windows(shn_ctx* c, const char* path, const char* bases, const uint64_t* offsets,
                            const uint32_t* contig_ids, uint64_t m, int k1, const uint32_t* weights,
                            const uint64_t* win_off) {
  SHN_API_BEGIN
  shn_write_k1mer_windows_impl(path, bases, offsets, contig_ids, m, k1, weights, win_off);
  SHN_API_END(c)
}

int shn_write_kmer_file(shn_ctx* c, const char* path, const uint64_t* keys, const uint32_t* counts,
                        uint64_t n, int k1) {
  SHN_API_BEGIN
  shn_write_kmer_file_impl(path, keys, counts, n, k1);
  SHN_API_END(c)
}

// ---- table -------------------------------------------------------------------------------------
int shn_pack_kmers(shn_ctx* c, const char* ascii, uint64_t n, int k1, uint64_t* keys, int on_device) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33");
  if (n) {
    DevBuf sa, sk;
    const char* d_ascii = (const char*)InputView::get(c, ascii, n * (uint64_t)k1, on_device, sa);
    uint64_t* d_keys = keys;
    if (!on_device) {
      sk.reserve(n * key_bytes(k1));
      d_keys = sk.as<uint64_t>();
    }
    SHN_DISPATCH(k1, pack_kmers(c, d_ascii, n, k1, d_keys));
    if (!on_device)
      CUDA_CHECK(cudaMemcpyAsync(keys, d_keys, n * key_bytes(k1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  SHN_API_END(c)
}

int shn_table_build(shn_ctx* c, const uint64_t* keys, const uint32_t* counts, uint64_t n, int k1,
                    int double_stranded, int on_device) {
  SHN_API_BEGIN
  bind(c);
  shn_l3_free(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33 (K <= 32)");
  const uint64_t kw = key_bytes(k1) / 8;
  if (on_device || n == 0) {
    SHN_DISPATCH(k1, table_build(c, keys, counts, n, k1, double_stranded, nullptr));
  } else {
    // host input: copy chunk i+1 on the copy stream while chunk i is being inserted
    DevBuf sk, sc;
    sk.reserve(n * 8 * kw);
    sc.reserve(n * 4);
    if (!c->stream2) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    SHN_DISPATCH(k1, table_begin(c, n, k1, double_stranded));
    const uint64_t chunk = 32ull << 20;
    std::vector<cudaEvent_t> evs;
    {
      cudaEvent_t ev = c->prof_event();  // the staging buffers may still be in use on the main stream
      CUDA_CHECK(cudaEventRecord(ev, c->stream));
      CUDA_CHECK(cudaStreamWaitEvent(c->stream2, ev, 0));
      evs.push_back(ev);
    }
    for (uint64_t lo = 0; lo < n; lo += chunk) {
      const uint64_t m = std::min(chunk, n - lo);
      CUDA_CHECK(cudaMemcpyAsync(sk.as<uint64_t>() + lo * kw, keys + lo * kw, m * 8 * kw, cudaMemcpyHostToDevice,
                                 c->stream2));
      CUDA_CHECK(cudaMemcpyAsync(sc.as<uint32_t>() + lo, counts + lo, m * 4, cudaMemcpyHostToDevice, c->stream2));
      cudaEvent_t ev = c->prof_event();
      CUDA_CHECK(cudaEventRecord(ev, c->stream2));
      CUDA_CHECK(cudaStreamWaitEvent(c->stream, ev, 0));
      evs.push_back(ev);
      SHN_DISPATCH(k1, table_insert_chunk(c, sk.as<uint64_t>() + lo * kw, sc.as<uint32_t>() + lo, nullptr,
                                          m, lo, double_stranded));
    }
    SHN_DISPATCH(k1, table_finish(c));
    for (cudaEvent_t ev : evs) c->prof_pool.push_back(ev);
  }
  SHN_API_END(c)
}

int shn_table_build_indexed(shn_ctx* c, const uint64_t* keys_dev, const uint32_t* counts_dev,
                            const uint32_t* line_idx_dev, uint64_t n, int k1) {
  SHN_API_BEGIN
  bind(c);
  shn_l3_free(c);
  SHN_DISPATCH(k1, table_build(c, keys_dev, counts_dev, n, k1, 0, line_idx_dev));
  SHN_API_END(c)
}

int shn_route_plan(shn_ctx* c, const uint64_t* keys_dev, uint64_t n, uint32_t nranks, uint32_t* perm_dev,
                   uint64_t* counts_host) {
  SHN_API_BEGIN
  bind(c);
  shn_route_plan_impl(c, keys_dev, n, nranks, perm_dev, counts_host);
  SHN_API_END(c)
}

int shn_permute(shn_ctx* c, const void* src_dev, const uint32_t* perm_dev, uint64_t n, int elem_bytes,
                int scatter, void* dst_dev) {
  SHN_API_BEGIN
  bind(c);
  shn_permute_impl(c, src_dev, perm_dev, n, elem_bytes, scatter, dst_dev);
  SHN_API_END(c)
}

int shn_table_stats(shn_ctx* c, uint64_t* n_distinct, uint64_t* n_lowcomplexity, uint64_t* n_slots,
                    int* k1) {
  SHN_API_BEGIN
  SHN_CHECK(c != nullptr, "null context");
  if (n_distinct) *n_distinct = c->n_distinct;
  if (n_lowcomplexity) *n_lowcomplexity = c->n_lowcomplexity;
  if (n_slots) *n_slots = c->n_buckets * (is_wide(c->k1) ? 2 : 4);
  if (k1) *k1 = c->k1;
  SHN_API_END(c)
}

int shn_table_lookup(shn_ctx* c, const uint64_t* keys, uint64_t n, uint32_t* weights, uint8_t* found,
                     int on_device) {
  SHN_API_BEGIN
  bind(c);
  if (n) {
    DevBuf sk, sw, sf;
    const uint64_t* dk = (const uint64_t*)InputView::get(c, keys, n * key_bytes(c->k1), on_device, sk);
    uint32_t* dw = weights;
    uint8_t* df = found;
    if (!on_device) {
      sw.reserve(n * 4);
      sf.reserve(n);
      dw = sw.as<uint32_t>();
      df = sf.as<uint8_t>();
    }
    SHN_DISPATCH(c->k1, table_lookup(c, dk, n, dw, df));
    if (!on_device) {
      if (weights) CUDA_CHECK(cudaMemcpyAsync(weights, dw, n * 4, cudaMemcpyDeviceToHost, c->stream));
      if (found) CUDA_CHECK(cudaMemcpyAsync(found, df, n, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  SHN_API_END(c)
}

int shn_table_dump(shn_ctx* c, uint64_t* keys, uint32_t* weights, uint32_t* first_idx) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, table_dump(c, keys, weights, first_idx));
  SHN_API_END(c)
}

// ---- L3 ----------------------------------------------------------------------------------------
int shn_l3_run(shn_ctx* c, uint32_t min_weight, uint32_t min_length) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_run(c, min_weight, min_length));
  SHN_API_END(c)
}
int shn_l3_get_sizes(shn_ctx* c, shn_l3_sizes* out) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_sizes(c, out));
  SHN_API_END(c)
}
int shn_l3_get_walks(shn_ctx* c, uint64_t* seed_keys, uint32_t* n_left, uint32_t* n_right,
                     uint64_t* tot_wt, uint8_t* flags) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_walks(c, seed_keys, n_left, n_right, tot_wt, flags));
  SHN_API_END(c)
}
int shn_l3_get_contigs(shn_ctx* c, char* bases, uint64_t* offsets) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_contigs(c, bases, offsets));
  SHN_API_END(c)
}
int shn_l3_get_allowed(shn_ctx* c, uint64_t* keys, uint32_t* weights) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_allowed(c, keys, weights));
  SHN_API_END(c)
}
int shn_l3_get_edges(shn_ctx* c, uint32_t* a, uint32_t* b, uint32_t* weight, uint32_t* first_pos_in_b) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_edges(c, a, b, weight, first_pos_in_b));
  SHN_API_END(c)
}
int shn_l3_get_labels(shn_ctx* c, uint32_t* label) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_get_labels(c, label));
  SHN_API_END(c)
}

int shn_l3_walks(shn_ctx* c, uint32_t min_weight, uint32_t min_length) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_walks(c, min_weight, min_length));
  SHN_API_END(c)
}
int shn_l3_cand_sizes(shn_ctx* c, uint64_t* n_cand, uint64_t* n_bases) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_cand_sizes(c, n_cand, n_bases));
  SHN_API_END(c)
}
int shn_l3_cand_export(shn_ctx* c, uint32_t* weight_dev, uint64_t* line_dev, uint64_t* offs_dev,
                       uint8_t* codes_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_cand_export(c, weight_dev, line_dev, offs_dev, codes_dev));
  SHN_API_END(c)
}
int shn_l3_filter(shn_ctx* c, const uint8_t* codes_dev, const uint64_t* offs_dev, uint64_t n_cand,
                  int external, int allow_missing) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_filter(c, codes_dev, offs_dev, n_cand, external, allow_missing));
  SHN_API_END(c)
}
int shn_l3_allowed_copy(shn_ctx* c, uint64_t* keys_dev, uint32_t* weights_dev) {
  SHN_API_BEGIN
  bind(c);
  const uint64_t* dk;
  const uint32_t* dw;
  uint64_t dn;
  SHN_DISPATCH(c->k1, l3_allowed_dev(c, &dk, &dw, &dn));
  if (dn && keys_dev)
    CUDA_CHECK(cudaMemcpyAsync(keys_dev, dk, dn * key_bytes(c->k1), cudaMemcpyDeviceToDevice, c->stream));
  if (dn && weights_dev) CUDA_CHECK(cudaMemcpyAsync(weights_dev, dw, dn * 4, cudaMemcpyDeviceToDevice, c->stream));
  SHN_API_END(c)
}
int shn_l3_set_allowed_weights(shn_ctx* c, const uint32_t* weights_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, l3_set_allowed_weights(c, weights_dev));
  SHN_API_END(c)
}

// ---- sharded tables --------------------------------------------------------------------
int shn_route_lines(shn_ctx* c, const uint64_t* keys_dev, const uint32_t* counts_dev, uint64_t n,
                    uint64_t first_line, int double_stranded, int k1, uint32_t nranks, uint64_t* counts_host,
                    void* send_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33");
  SHN_DISPATCH(k1, route_lines(c, keys_dev, counts_dev, n, first_line, double_stranded, k1, nranks, counts_host,
                               send_dev));
  SHN_API_END(c)
}
int shn_table_build_records(shn_ctx* c, const void* recs_dev, uint64_t n, int k1) {
  SHN_API_BEGIN
  bind(c);
  shn_l3_free(c);
  shn_shard_free(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33");
  SHN_DISPATCH(k1, table_build_records(c, recs_dev, n, k1));
  SHN_API_END(c)
}
int shn_cc_local(shn_ctx* c, uint64_t* n_local) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_local(c, n_local));
  SHN_API_END(c)
}
int shn_cc_cross(shn_ctx* c, uint32_t nranks, uint32_t rank, uint64_t gid_base, uint64_t* counts_host,
                 void* send_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_cross(c, nranks, rank, gid_base, counts_host, send_dev));
  SHN_API_END(c)
}
int shn_cc_resolve(shn_ctx* c, const void* recs_dev, uint64_t n, uint64_t gid_base, uint64_t* edges_dev,
                   uint64_t* n_edges) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_resolve(c, recs_dev, n, gid_base, edges_dev, n_edges));
  SHN_API_END(c)
}
int shn_cc_merge(shn_ctx* c, const uint64_t* edges_dev, uint64_t n_edges, uint64_t n_super, uint64_t* n_final) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_merge(c, edges_dev, n_edges, n_super, n_final));
  SHN_API_END(c)
}
int shn_cc_sizes(shn_ctx* c, uint64_t gid_base, uint64_t* sizes_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_sizes(c, gid_base, sizes_dev));
  SHN_API_END(c)
}
int shn_cc_route(shn_ctx* c, const uint32_t* owner_of_final_dev, uint64_t gid_base, uint32_t nranks,
                 uint64_t* counts_host, void* send_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->k1, cc_route(c, owner_of_final_dev, gid_base, nranks, counts_host, send_dev));
  SHN_API_END(c)
}
int shn_cc_free(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  shn_shard_free(c);
  SHN_API_END(c)
}

// ---- L4 ----------------------------------------------------------------------------------------
int shn_l4_map_add_contigs(shn_ctx* c, const char* bases, const uint64_t* offsets,
                           const uint32_t* comp_of_contig, uint64_t n_contigs, int k1, int reset,
                           uint64_t expected_total_k1mers) {
  SHN_API_BEGIN
  bind(c);
  if (reset) c->l4_k1 = k1;
  SHN_CHECK(k1 == c->l4_k1, "k1 differs from the component map's");
  SHN_DISPATCH(k1, l4_map_add_contigs(c, bases, offsets, comp_of_contig, n_contigs, k1, reset,
                                      expected_total_k1mers, 0, 0));
  SHN_API_END(c)
}
int shn_l4_map_add_l3_contigs(shn_ctx* c, const uint32_t* comp_of_contig, uint64_t n_contigs, int reset) {
  SHN_API_BEGIN
  bind(c);
  const uint8_t* codes;
  const uint64_t* offs;
  uint64_t n, n_allowed;
  SHN_DISPATCH(c->k1, l3_contigs_dev(c, &codes, &offs, &n, &n_allowed));
  SHN_CHECK(n == n_contigs, "component id array does not match the number of accepted contigs");
  if (reset) c->l4_k1 = c->k1;
  SHN_DISPATCH(c->k1, l4_map_add_contigs(c, (const char*)codes, offs, comp_of_contig, n, c->k1, reset,
                                         n_allowed, 1, 1));
  SHN_API_END(c)
}
int shn_l4_map_set_weights(shn_ctx* c, const uint64_t* dict_keys, const uint32_t* dict_weights,
                           uint64_t n) {
  SHN_API_BEGIN
  bind(c);
  if (dict_keys == nullptr) {  // k1mer_dictionary = the allowed set this ctx's shn_l3_run produced
    const uint64_t* dk;
    const uint32_t* dw;
    uint64_t dn;
    SHN_DISPATCH(c->k1, l3_allowed_dev(c, &dk, &dw, &dn));
    SHN_CHECK(c->k1 == c->l4_k1, "the allowed set and the component map use different k1");
    SHN_DISPATCH(c->l4_k1, l4_map_set_weights(c, dk, dw, dn, 1));
  } else {
    SHN_DISPATCH(c->l4_k1, l4_map_set_weights(c, dict_keys, dict_weights, n, 0));
  }
  SHN_API_END(c)
}
int shn_l4_map_window_weights(shn_ctx* c, const char* bases, const uint64_t* offsets,
                              uint64_t n_contigs, int k1, uint32_t* weights) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 == c->l4_k1, "k1 differs from the component map's");
  SHN_DISPATCH(k1, l4_map_window_weights(c, bases, offsets, n_contigs, k1, weights));
  SHN_API_END(c)
}
int shn_l4_load_reads(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets,
                      uint64_t n_reads, int on_device) {
  SHN_API_BEGIN
  bind(c);
  shn_reads_load(c, mate, bases, offsets, n_reads, on_device);
  SHN_API_END(c)
}
int shn_l4_upload_reads_async(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets,
                              uint64_t n_reads) {
  SHN_API_BEGIN
  bind(c);
  shn_reads_upload_async(c, mate, bases, offsets, n_reads);
  SHN_API_END(c)
}
int shn_l4_load_reads_staged(shn_ctx* c, int mate) {
  SHN_API_BEGIN
  bind(c);
  shn_reads_load_staged(c, mate);
  SHN_API_END(c)
}
int shn_l4_assign(shn_ctx* c, int paired, int k1, uint64_t* n_assignments, uint64_t* n_lookups,
                  uint64_t* n_valid_records) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 == c->l4_k1, "k1 differs from the component map's");
  SHN_DISPATCH(k1, l4_assign(c, paired, k1, n_assignments, n_lookups, n_valid_records));
  SHN_API_END(c)
}
int shn_l4_get_assignments(shn_ctx* c, uint32_t n_comps, uint64_t* comp_offsets, uint32_t* record_idx) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->l4_k1, l4_get_assignments(c, n_comps, comp_offsets, record_idx));
  SHN_API_END(c)
}

int shn_l4_assignments_dev(shn_ctx* c, uint32_t n_comps, uint64_t first_record, uint64_t* comp_offsets_dev,
                           uint32_t* record_idx_dev) {
  SHN_API_BEGIN
  bind(c);
  SHN_DISPATCH(c->l4_k1, l4_assignments_dev(c, n_comps, first_record, comp_offsets_dev, record_idx_dev));
  SHN_API_END(c)
}

// ---- inputs of the path --------------------------------------------------------------------------
int shn_synth_pairs(shn_ctx* c, const uint8_t* tx_codes, const uint64_t* tx_offs,
                    const uint64_t* thresholds, uint64_t n_tx, uint64_t n_pairs, uint64_t first_pair,
                    uint64_t seed, int read_len, int frag_len, uint32_t err_threshold_24,
                    char* mate1_dev, char* mate2_dev) {
  SHN_API_BEGIN
  bind(c);
  shn_synth_pairs_impl(c, tx_codes, tx_offs, thresholds, n_tx, n_pairs, first_pair, seed, read_len,
                       frag_len, err_threshold_24, mate1_dev, mate2_dev);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_API_END(c)
}
int shn_revcomp_reads(shn_ctx* c, const char* in_dev, char* out_dev, uint64_t n_reads, int read_len) {
  SHN_API_BEGIN
  bind(c);
  shn_revcomp_reads_impl(c, in_dev, out_dev, n_reads, read_len);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_API_END(c)
}
int shn_revcomp_var(shn_ctx* c, const char* bases, const uint64_t* offsets, uint64_t n_reads, char* out,
                    int on_device) {
  SHN_API_BEGIN
  bind(c);
  if (n_reads) {
    uint64_t total = 0;
    DevBuf sb, so, sout;
    const uint64_t* d_offs = offsets;
    if (on_device) {
      CUDA_CHECK(cudaMemcpyAsync(&total, offsets + n_reads, 8, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
      total = offsets[n_reads];
      d_offs = (const uint64_t*)InputView::get(c, offsets, (n_reads + 1) * 8, 0, so);
    }
    const char* d_in = (const char*)InputView::get(c, bases, total, on_device, sb);
    char* d_out = out;
    if (!on_device) {
      sout.reserve(std::max<uint64_t>(total, 1));
      d_out = sout.as<char>();
    }
    shn_revcomp_var_impl(c, d_in, d_offs, n_reads, total, d_out);
    if (!on_device && total) {
      CUDA_CHECK(cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
  }
  SHN_API_END(c)
}
int shn_count_begin(shn_ctx* c, int k1, uint64_t expected_distinct) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33");
  SHN_DISPATCH(k1, count_begin(c, k1, expected_distinct));
  c->count_k1 = k1;
  SHN_API_END(c)
}
int shn_count_add_reads(shn_ctx* c, const char* bases, const uint64_t* offsets, uint64_t n_reads, int on_device) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(c->count_k1 > 0, "shn_count_begin has not been called");
  if (n_reads) {
    uint64_t total = 0;
    DevBuf sb, so;
    const uint64_t* d_offs = offsets;
    if (on_device) {
      CUDA_CHECK(cudaMemcpyAsync(&total, offsets + n_reads, 8, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
      total = offsets[n_reads];
      d_offs = (const uint64_t*)InputView::get(c, offsets, (n_reads + 1) * 8, 0, so);
    }
    const char* d_in = (const char*)InputView::get(c, bases, total, on_device, sb);
    SHN_DISPATCH(c->count_k1, count_add_var(c, d_in, d_offs, n_reads, total));
  }
  SHN_API_END(c)
}
int shn_count_finish(shn_ctx* c, uint32_t min_count, uint64_t** keys_dev, uint32_t** counts_dev,
                     uint64_t* n_distinct) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(c->count_k1 > 0, "shn_count_begin has not been called");
  SHN_DISPATCH(c->count_k1, count_finish(c, min_count, keys_dev, counts_dev, n_distinct));
  SHN_API_END(c)
}
int shn_load_fasta_named(shn_ctx* c, const char* path, char** names, uint64_t** name_offsets, char** bases,
                         uint64_t** offsets, uint64_t* n) {
  SHN_API_BEGIN
  shn_load_fasta_named_impl(path, names, name_offsets, bases, offsets, n);
  SHN_API_END(c)
}
int shn_write_fasta_named(shn_ctx* c, const char* path, int append, const char* names,
                          const uint64_t* name_offsets, const char* bases, const uint64_t* offsets, uint64_t n) {
  SHN_API_BEGIN
  shn_write_fasta_named_impl(path, append, names, name_offsets, bases, offsets, n);
  SHN_API_END(c)
}
int shn_condense_run(shn_ctx* c, const uint64_t* prefix_kmers, const uint64_t* suffix_kmers,
                     const uint32_t* prevalence, uint64_t n, int K, uint64_t* n_unitigs, uint64_t* n_bases,
                     uint64_t* n_edges, uint64_t* n_cycle_nodes) {
  SHN_API_BEGIN
  bind(c);
  shn_condense_run_impl(c, prefix_kmers, suffix_kmers, prevalence, n, K, n_unitigs, n_bases, n_edges, n_cycle_nodes);
  SHN_API_END(c)
}
int shn_condense_get(shn_ctx* c, char* bases, uint64_t* offsets, uint32_t* count, uint64_t* prevalence,
                     uint32_t* edge_src, uint32_t* edge_dst, uint32_t* edge_copy_count) {
  SHN_API_BEGIN
  bind(c);
  shn_condense_get_impl(c, bases, offsets, count, prevalence, edge_src, edge_dst, edge_copy_count);
  SHN_API_END(c)
}
int shn_find_reps(shn_ctx* c, const char* bases, const uint64_t* offsets, const uint32_t* name_rank,
                  uint64_t n, int double_stranded, uint8_t* duplicate_out) {
  SHN_API_BEGIN
  bind(c);
  shn_find_reps_impl(c, bases, offsets, name_rank, n, double_stranded, duplicate_out);
  SHN_API_END(c)
}
int shn_count_release(shn_ctx* c) {
  SHN_API_BEGIN
  bind(c);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  shn_count_free(c);
  SHN_API_END(c)
}
int shn_count_k1mers(shn_ctx* c, const char* const* read_arrays_dev, const uint64_t* n_reads,
                     int n_arrays, int read_len, int k1, uint64_t expected_distinct,
                     uint64_t** keys_dev, uint32_t** counts_dev, uint64_t* n_distinct) {
  SHN_API_BEGIN
  bind(c);
  SHN_CHECK(k1 >= 1 && k1 <= 33, "k1 must be in 1..33");
  SHN_DISPATCH(k1, count_k1mers(c, read_arrays_dev, n_reads, n_arrays, read_len, k1, expected_distinct,
                                keys_dev, counts_dev, n_distinct));
  SHN_API_END(c)
}

}  // extern "C"
