// Shared declarations for the shannon_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/shannon_b200.h"

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
struct ShnError : public std::runtime_error {
  explicit ShnError(const std::string& m) : std::runtime_error(m) {}
};

[[noreturn]] void shn_throw(const char* file, int line, const std::string& msg);

#define SHN_FAIL(msg) shn_throw(__FILE__, __LINE__, (msg))
#define SHN_CHECK(cond, msg) \
  do {                       \
    if (!(cond)) SHN_FAIL(msg); \
  } while (0)
#define CUDA_CHECK(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      SHN_FAIL(std::string(#expr) + " -> " + cudaGetErrorName(_e) + ": " +          \
               cudaGetErrorString(_e));                                             \
  } while (0)

// ---------------------------------------------------------------------------------------
// 2-bit k-mer arithmetic (host + device).  Code: A=0 G=1 C=2 T=3, complement = 3-code.
// ---------------------------------------------------------------------------------------
#define SHN_HD __host__ __device__ __forceinline__

static const uint64_t SHN_EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

// ASCII -> code, 4 for anything that is not ACGT/acgt.
SHN_HD uint32_t shn_code_of(uint32_t c) {
  c &= 0xDFu;  // upper-case
  // (c>>1)&3 : A->0 C->1 T->2 G->3 ; remap to A0 G1 C2 T3 with the constant 0b01'11'10'00
  uint32_t x = (c >> 1) & 3u;
  uint32_t code = (0x78u >> (2u * x)) & 3u;
  bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
  return ok ? code : 4u;
}
// strict variant used for reads: lower case is NOT a base (read.strip('ACTG'))
SHN_HD uint32_t shn_code_of_strict(uint32_t c) {
  uint32_t x = (c >> 1) & 3u;
  uint32_t code = (0x78u >> (2u * x)) & 3u;
  bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
  return ok ? code : 4u;
}
SHN_HD char shn_base_of(uint32_t code) { return "AGCT"[code & 3u]; }

SHN_HD uint64_t shn_kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// reverse complement of a k-base key
SHN_HD uint64_t shn_revcomp(uint64_t x, int k) {
  x = ~x;  // complement every pair (3 - code)
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  x = (x >> 32) | (x << 32);
  return x >> (64 - 2 * k);
}

#if defined(__CUDA_ARCH__)
#define SHN_POPC64(x) __popcll(x)
#else
#define SHN_POPC64(x) __builtin_popcountll(x)
#endif

// lowComplexity (extension_correction.py:142-149): max base count >= k-2
SHN_HD bool shn_low_complexity(uint64_t x, int k) {
  const uint64_t m = 0x5555555555555555ull & shn_kmer_mask(k);
  uint64_t lo = x & m, hi = (x >> 1) & m;
  int nT = SHN_POPC64(hi & lo);
  int nC = SHN_POPC64(hi & ~lo);
  int nG = SHN_POPC64(~hi & lo & m);
  int nA = k - nT - nC - nG;
  int mx = nA > nC ? nA : nC;
  mx = mx > nG ? mx : nG;
  mx = mx > nT ? mx : nT;
  return mx >= k - 2;
}

// murmur3 fmix64
SHN_HD uint64_t shn_mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xFF51AFD7ED558CCDull;
  x ^= x >> 33;
  x *= 0xC4CEB9FE1A85EC53ull;
  x ^= x >> 33;
  return x;
}

// map A0 G1 C2 T3 pairs to A0 C1 G2 T3 pairs (swap the two bits of every pair): integer order of
// the result is the ASCII order of the k-mer string.
SHN_HD uint64_t shn_ascii_order_key(uint64_t x) {
  return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
}

// ---- the same for K1-mers of 33..64 bases (K = 32..63): two 64-bit words --------------------------
typedef unsigned __int128 u128;

SHN_HD u128 shn_kmer_mask128(int k) { return k >= 64 ? ~(u128)0 : (((u128)1 << (2 * k)) - 1); }
SHN_HD u128 shn_revcomp(u128 x, int k) {
  // reverse-complement both halves as 32-base words, swap them, drop the unused low pairs
  uint64_t lo = shn_revcomp((uint64_t)x, 32), hi = shn_revcomp((uint64_t)(x >> 64), 32);
  u128 y = ((u128)lo << 64) | hi;
  return y >> (128 - 2 * k);
}
SHN_HD bool shn_low_complexity(u128 x, int k) {
  const uint64_t m0 = 0x5555555555555555ull;
  const uint64_t m1 = 0x5555555555555555ull & shn_kmer_mask(k - 32);
  uint64_t a = (uint64_t)x, b = (uint64_t)(x >> 64);
  uint64_t lo0 = a & m0, hi0 = (a >> 1) & m0, lo1 = b & m1, hi1 = (b >> 1) & m1;
  int nT = SHN_POPC64(hi0 & lo0) + SHN_POPC64(hi1 & lo1);
  int nC = SHN_POPC64(hi0 & ~lo0) + SHN_POPC64(hi1 & ~lo1);
  int nG = SHN_POPC64(~hi0 & lo0) + SHN_POPC64(~hi1 & lo1 & m1);
  int nA = k - nT - nC - nG;
  int mx = nA > nC ? nA : nC;
  mx = mx > nG ? mx : nG;
  mx = mx > nT ? mx : nT;
  return mx >= k - 2;
}
SHN_HD u128 shn_ascii_order_key(u128 x) {
  return ((u128)shn_ascii_order_key((uint64_t)(x >> 64)) << 64) | shn_ascii_order_key((uint64_t)x);
}
SHN_HD uint64_t shn_key_hash(uint64_t x) { return shn_mix64(x); }
SHN_HD uint64_t shn_key_hash(u128 x) {
  return shn_mix64((uint64_t)x ^ shn_mix64((uint64_t)(x >> 64) + 0x9E3779B97F4A7C15ull));
}

// Placement hash of the K1-mer table: only the TOP bits are used (bucket inside a region, or
// umulhi with the bucket count), and the top bits of a product depend on every bit of the operand,
// so one multiplication by an odd constant is enough (the walks pay for this hash once per probe on
// their critical path; the two-multiply finalizer cost 17 dependent instructions).
SHN_HD uint64_t shn_bucket_hash(uint64_t x) { return x * 0x9E3779B97F4A7C15ull; }
SHN_HD uint64_t shn_bucket_hash(u128 x) {
  return ((uint64_t)x ^ ((uint64_t)(x >> 64) * 0xC4CEB9FE1A85EC53ull)) * 0x9E3779B97F4A7C15ull;
}

// ---------------------------------------------------------------------------------------
// K1-mer table: open addressing, 64-byte buckets (one DRAM burst per probe), linear probing over
// buckets; see table_dev.cuh.  Every key-dependent translation unit is compiled twice: once with
// 64-bit keys (K1 <= 32: four 16-byte slots per bucket) into namespace `narrow`, once with
// -DSHN_WIDE (K1 = 33..64: two 32-byte slots per bucket, 128-bit keys) into namespace `wide`;
// api.cu dispatches on k1.  Key arrays cross the C-ABI as SHN_KEY_WORDS uint64 words per key,
// low word first.
// ---------------------------------------------------------------------------------------
static const uint32_t SHN_TRAVERSED = 0x80000000u;
static const uint32_t SHN_OVERFLOW = 0x40000000u;
static const uint32_t SHN_WEIGHT_MASK = 0x3FFFFFFFu;
static const uint32_t SHN_NONE32 = 0xFFFFFFFFu;

#ifdef SHN_WIDE
#define SHN_NS wide
#define SHN_KEY_WORDS 2
#define SHN_BSLOTS 2
#define SHN_MAX_K1 33 /* contig-overlap words (C = K1-1 bases) must fit 64 bits */
typedef u128 shn_key_t;
struct __align__(32) ShnSlot {
  u128 key;         // all ones when free
  uint32_t weight;  // as below
  uint32_t idx;
  uint64_t pad;
};
SHN_HD shn_key_t shn_key_mask(int k) { return shn_kmer_mask128(k); }
#else
#define SHN_NS narrow
#define SHN_KEY_WORDS 1
#define SHN_BSLOTS 4
#define SHN_MAX_K1 32
typedef uint64_t shn_key_t;
struct __align__(16) ShnSlot {
  uint64_t key;     // SHN_EMPTY_KEY when free
  uint32_t weight;  // sum of counts (30 bits); bit 30 of the bucket's slot 0 = an insert walked
                    // past this full bucket
  uint32_t idx;     // first-occurrence index in the input (dict insertion order); parked in a side
                    // array during shn_l3_run, when the word is the walks' aux word (l3.cu)
};
SHN_HD shn_key_t shn_key_mask(int k) { return shn_kmer_mask(k); }
#endif
#define SHN_EMPTY ((shn_key_t)~(shn_key_t)0)

// Placement by the K-base PREFIX of the K1-mer.  The home bucket is a function of key >> 2 alone, so
// the four successors (x[1:] . b) of a K1-mer x -- which share their first K bases -- share one home
// bucket and one probe sequence: uf_edges finds all of them with ONE walk over the buckets instead
// of four, and the four first-level probes of a right extension read the same 64 bytes.  (A bucket
// then holds whole successor families; at load 0.5 8 % of the buckets overflow against 5 % for
// independent keys.)
// Minimizer-clustered regions on top: the table is cut into regions of 2^region_shift buckets
// (8 MB) and a K1-mer goes to region hash(minimizer of its prefix), bucket hash(prefix) inside the
// region.  Consecutive K1-mers of a chain share their minimizer -- the 12-mer with the smallest
// hash -- about 7 times out of 8, so successor / predecessor probes mostly stay inside the 8 MB the
// kernel is streaming through (L2 hits instead of one DRAM burst each: uf_edges), and so do the
// parent words of the union-find.  Overflow still probes linearly over buckets, across regions.
// Tables smaller than one region, and k1 < 13, use the plain hash of the prefix.
constexpr int kRegionM = 12;
// hash of the 12-mer at base offset p counted from the END of the K1-mer (bits 2p .. 2p+23)
SHN_HD uint32_t shn_mmer_hash(shn_key_t key, int p) {
  return ((uint32_t)(key >> (2 * p)) & ((1u << (2 * kRegionM)) - 1u)) * 0x9E3779B1u;  // odd multiplier: a bijection
}
// smallest 12-mer hash over the offsets [p_lo, p_hi]
SHN_HD uint32_t shn_minimizer_hash_range(shn_key_t key, int p_lo, int p_hi) {
  uint32_t best = 0xFFFFFFFFu;
  for (int p = p_lo; p <= p_hi; ++p) {
    const uint32_t h = shn_mmer_hash(key, p);
    best = h < best ? h : best;
  }
  return best;
}
// minimizer of the K-base prefix: the 12-mers that do not contain the last base (offsets 1 .. k1-12)
SHN_HD uint32_t shn_minimizer_hash(shn_key_t key, int k1) { return shn_minimizer_hash_range(key, 1, k1 - kRegionM); }

struct ShnTableView {
  ShnSlot* slots;      // SHN_BSLOTS * n_buckets
  uint64_t n_buckets;
  uint32_t n_regions = 0;  // 0: plain hashing
  int k1 = 0;
  int region_shift = 17;   // log2(buckets per region): 2^17 x 64 B = 8 MB
  __device__ __forceinline__ uint64_t bucket_of(shn_key_t key) const {
    if (n_regions == 0) return __umul64hi(shn_bucket_hash(key >> 2), n_buckets);
    return bucket_with_min(key, shn_minimizer_hash(key, k1));
  }
  // the same when the caller already knows the minimizer hash of the K1-mer's prefix (neighbouring
  // K1-mers share all 12-mers but one or two: the walks compute the shared minimum once)
  __device__ __forceinline__ uint64_t bucket_with_min(shn_key_t key, uint32_t min_hash) const {
    const uint64_t h = shn_bucket_hash(key >> 2);
    if (n_regions == 0) return __umul64hi(h, n_buckets);
    // the minimum of a dozen 12-mer hashes is a small number: re-spread it before taking top bits
    const uint32_t region = __umulhi(min_hash * 0x85EBCA6Bu, n_regions);
    return ((uint64_t)region << region_shift) | (h >> (64 - region_shift));
  }
};

// key arrays at the C-ABI: SHN_KEY_WORDS uint64 per key, low word first
SHN_HD shn_key_t shn_load_key(const uint64_t* keys, uint64_t i) {
#ifdef SHN_WIDE
  return ((u128)keys[2 * i + 1] << 64) | keys[2 * i];
#else
  return keys[i];
#endif
}
SHN_HD void shn_store_key(uint64_t* keys, uint64_t i, shn_key_t k) {
#ifdef SHN_WIDE
  keys[2 * i] = (uint64_t)k;
  keys[2 * i + 1] = (uint64_t)(k >> 64);
#else
  keys[i] = k;
#endif
}

// ---------------------------------------------------------------------------------------
// device buffers and the context
// ---------------------------------------------------------------------------------------
// Caching device allocator, one per context.  Every kernel of a context runs on that context's
// single stream, so a block handed back by one buffer can be reused by the next allocation without
// synchronisation (stream order protects it).  Blocks are never returned to the driver between
// steps: after the first pass of a workload every allocation is a free-list hit (cudaMalloc /
// cudaFree of multi-GB blocks cost tens to hundreds of ms and serialise the device).
struct DevPool {
  std::multimap<uint64_t, void*> free_blocks;   // size -> block
  std::map<void*, uint64_t> live;                // block -> size
  uint64_t cached_bytes = 0, live_bytes = 0;
  static uint64_t round_up(uint64_t n) {
    const uint64_t g = n >= (64ull << 20) ? (2ull << 20) : 512ull;
    return (n + g - 1) / g * g;
  }
  void* alloc(uint64_t nbytes) {
    const uint64_t want = round_up(nbytes);
    auto it = free_blocks.lower_bound(want);
    if (it != free_blocks.end() && it->first <= want + want / 4 + (1ull << 20)) {
      void* p = it->second;
      uint64_t sz = it->first;
      free_blocks.erase(it);
      cached_bytes -= sz;
      live[p] = sz;
      live_bytes += sz;
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim();  // give cached blocks back to the driver and retry once
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      SHN_FAIL("device allocation of " + std::to_string(want) + " bytes failed: " +
               cudaGetErrorString(e));
    }
    live[p] = want;
    live_bytes += want;
    return p;
  }
  void release(void* p) {
    auto it = live.find(p);
    if (it == live.end()) {
      cudaFree(p);
      return;
    }
    free_blocks.emplace(it->second, p);
    cached_bytes += it->second;
    live_bytes -= it->second;
    live.erase(it);
  }
  void trim() {
    cudaDeviceSynchronize();
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
};

extern thread_local DevPool* g_shn_pool;  // set by bind() in api.cu

struct DevBuf {
  void* p = nullptr;
  uint64_t bytes = 0;
  DevPool* pool = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      if (pool)
        pool->release(p);
      else
        cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
    pool = nullptr;
  }
  // grow-only (re)allocation; contents are NOT preserved
  void reserve(uint64_t nbytes) {
    if (nbytes <= bytes) return;
    release();
    if (nbytes == 0) return;
    pool = g_shn_pool;
    if (pool) {
      p = pool->alloc(nbytes);
    } else {
      cudaError_t e = cudaMalloc(&p, nbytes);
      if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        SHN_FAIL("device allocation of " + std::to_string(nbytes) + " bytes failed: " +
                 cudaGetErrorString(e));
      }
    }
    bytes = nbytes;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct ProfEntry {
  double ms = 0;
  uint64_t launches = 0;
};
struct ProfPending {  // an event pair recorded on the stream, resolved lazily (no sync per kernel)
  cudaEvent_t e0, e1;
  std::string name;
  uint64_t launches;
};


struct shn_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // asynchronous uploads (created on demand)
  bool explicit_idx = false;  // the table was built with caller-supplied first-occurrence indices (any 32-bit value)
  cudaStream_t stream3 = nullptr, stream4 = nullptr, stream5 = nullptr;  // side streams of the walk stage (created on demand)
  std::string last_error;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  // profiling
  bool prof_on = false;
  cudaEvent_t p0 = nullptr, p1 = nullptr;
  std::map<std::string, ProfEntry> prof;
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  uint64_t launches = 0;
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) {
      cudaEvent_t e = prof_pool.back();
      prof_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  // fold all finished event pairs into `prof` (synchronises the stream)
  void prof_resolve() {
    if (prof_pending.empty()) return;
    cudaStreamSynchronize(stream);
    const bool gaps = getenv("SHN_PROF_GAPS") != nullptr;  // GPU-idle time in front of each scope
    cudaEvent_t prev = nullptr;
    for (auto& p : prof_pending) {
      float ms = 0;
      if (gaps && prev) {
        cudaEventElapsedTime(&ms, prev, p.e0);
        if (ms > 0) prof[std::string("gap<") + p.name].ms += ms;
      }
      prev = p.e1;
      cudaEventElapsedTime(&ms, p.e0, p.e1);
      ProfEntry& e = prof[p.name];
      e.ms += ms;
      e.launches += p.launches;
      prof_pool.push_back(p.e0);
      prof_pool.push_back(p.e1);
    }
    prof_pending.clear();
  }
  DevPool pool;
  // scratch
  DevBuf cub_tmp;
  DevBuf flush_buf;
  DevBuf counters;  // small device array of uint64 counters
  // K1-mer weight table
  DevBuf table;
  uint64_t n_buckets = 0;
  uint32_t n_regions = 0;  // minimizer-clustered placement (ShnTableView); 0 = plain hashing
  int region_shift = 17;
  int k1 = 0;
  int l4_k1 = 0;  // k1 of the component map (may differ from the table's in a fresh process)
  uint64_t n_distinct = 0, n_lowcomplexity = 0, n_items = 0;
  // sharded path: the idx word of a slot is the position of its record in the receive buffer and
  // gline_dev[position] the global input line (nullptr: idx IS the line)
  DevBuf gline_buf;
  const uint64_t* gline_dev = nullptr;
  // stage states, owned by the key-width specific code that created them
  void* l3 = nullptr;
  void (*l3_free)(shn_ctx*) = nullptr;
  void* l4 = nullptr;
  void (*l4_free)(shn_ctx*) = nullptr;
  int count_k1 = 0;  // k1 of the counting table between shn_count_begin and shn_count_finish
  void* count_state = nullptr;
  void (*count_free)(shn_ctx*) = nullptr;
  void* reads = nullptr;  // packed reads (reads.cu), independent of the key width
  void (*reads_free)(shn_ctx*) = nullptr;
  void* condense = nullptr;  // unitig condensation of a component (condense.cu)
  void (*condense_free)(shn_ctx*) = nullptr;
  void* shard = nullptr;  // cross-rank component labelling (shard.cu)
  void (*shard_free)(shn_ctx*) = nullptr;
  cudaStream_t own_stream = nullptr;  // the stream shn_create made (stream may point elsewhere: shn_use_stream)

  void* tmp(uint64_t bytes) {
    cub_tmp.reserve(bytes);
    return cub_tmp.p;
  }
};

// RAII profiling scope: brackets a launch group with a pair of CUDA events on the ctx stream
// when profiling is enabled.  Nothing synchronises here; times are read in prof_resolve().
struct ProfScope {
  shn_ctx* c;
  const char* name;
  uint64_t n;
  cudaEvent_t e0 = nullptr;
  ProfScope(shn_ctx* ctx, const char* nm, uint64_t launches = 1) : c(ctx), name(nm), n(launches) {
    c->launches += n;
    if (c->prof_on) {
      e0 = c->prof_event();
      cudaEventRecord(e0, c->stream);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEvent_t e1 = c->prof_event();
      cudaEventRecord(e1, c->stream);
      c->prof_pending.push_back(ProfPending{e0, e1, name, n});
    }
  }
};

static inline unsigned shn_grid(uint64_t n, unsigned block) {
  uint64_t g = (n + block - 1) / block;
  if (g == 0) g = 1;
  if (g > 0x7FFFFFFFull) SHN_FAIL("grid too large");
  return (unsigned)g;
}

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// copies with on_device switch ------------------------------------------------------------
struct InputView {
  // Makes `n_bytes` of caller data available on the device: either the pointer itself
  // (on_device) or a staged copy in `stage`.
  static const void* get(shn_ctx* c, const void* p, uint64_t n_bytes, int on_device, DevBuf& stage) {
    if (on_device || n_bytes == 0) return p;
    stage.reserve(n_bytes);
    CUDA_CHECK(cudaMemcpyAsync(stage.p, p, n_bytes, cudaMemcpyHostToDevice, c->stream));
    return stage.p;
  }
};

// releases the L3 / L4 / counter state of a context, whichever key width created it
static inline void shn_l3_free(shn_ctx* c) {
  if (c->l3 && c->l3_free) c->l3_free(c);
  c->l3 = nullptr;
}
static inline void shn_l4_free(shn_ctx* c) {
  if (c->l4 && c->l4_free) c->l4_free(c);
  c->l4 = nullptr;
}
static inline void shn_reads_free(shn_ctx* c) {
  if (c->reads && c->reads_free) c->reads_free(c);
  c->reads = nullptr;
}
static inline void shn_shard_free(shn_ctx* c) {
  if (c->shard && c->shard_free) c->shard_free(c);
  c->shard = nullptr;
}
static inline void shn_count_free(shn_ctx* c) {
  if (c->count_state && c->count_free) c->count_free(c);
  c->count_state = nullptr;
}

namespace SHN_NS {
static inline ShnTableView table_view(const shn_ctx* c) {
  return ShnTableView{c->table.as<ShnSlot>(), c->n_buckets, c->n_regions, c->k1, c->region_shift};
}
}  // namespace SHN_NS
