// Inputs of the hot path produced on the device (SURVEY 8f rows f1/f2 + the synthetic configs):
//  - shn_synth_pairs   : seeded synthetic read pairs (twin of shannon_b200/synth.py::make_pairs)
//  - shn_revcomp_reads : RC doubling of fixed-length reads (rc_gnu.py / rc_s.py, shannon.py:395-424)
//  - shn_count_k1mers  : jellyfish count/dump stand-in (shannon.py:439-441), ASCII-sorted output
#include <cub/cub.cuh>

#include "common.cuh"
#include "impls.h"
#include "table_dev.cuh"

#ifndef SHN_WIDE  // the generators do not depend on the key width: compiled once
namespace {

constexpr int kBlock = 256;
constexpr uint64_t kGold = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ uint64_t stream64(uint64_t seed, uint64_t index, uint64_t lane) {
  uint64_t a = shn_mix64(seed * kGold + index);
  return shn_mix64(a ^ (lane * kGold));
}

// one thread per (pair, base position): both mates of that position
__global__ void __launch_bounds__(kBlock)
    synth_pairs_kernel(const uint8_t* __restrict__ tx, const uint64_t* __restrict__ tx_offs,
                       const uint64_t* __restrict__ thr, uint64_t n_tx, uint64_t n_pairs,
                       uint64_t first_pair, uint64_t seed, int read_len, int frag_len,
                       uint32_t err_thr, char* __restrict__ m1, char* __restrict__ m2) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_pairs * (uint64_t)read_len) return;
  uint64_t pi = g / read_len;
  int j = (int)(g - pi * read_len);
  uint64_t p = first_pair + pi;
  // transcript: first t with r < thr[t]  (numpy searchsorted(thr, r, side='right'))
  uint64_t r = stream64(seed, p, 0) >> 1;
  uint64_t lo = 0, hi = n_tx;  // answer in [lo, hi]
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&thr[mid]) <= r)
      lo = mid + 1;
    else
      hi = mid;
  }
  uint64_t t = lo < n_tx ? lo : n_tx - 1;
  uint64_t t0 = __ldg(&tx_offs[t]), tlen = __ldg(&tx_offs[t + 1]) - t0;
  uint64_t start = stream64(seed, p, 1) % (tlen - (uint64_t)frag_len + 1);
  uint64_t base = t0 + start;
  uint32_t c1 = __ldg(&tx[base + j]);
  uint32_t c2 = 3u - __ldg(&tx[base + (frag_len - 1 - j)]);
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    uint64_t e = stream64(seed, p, (uint64_t)(2 + m * read_len + j));
    uint32_t code = m == 0 ? c1 : c2;
    if ((uint32_t)(e & 0xFFFFFFu) < err_thr) code = (code + 1u + (uint32_t)((e >> 24) % 3u)) & 3u;
    (m == 0 ? m1 : m2)[g] = shn_base_of(code);
  }
}

__global__ void __launch_bounds__(kBlock)
    revcomp_reads_kernel(const char* __restrict__ in, char* __restrict__ out, uint64_t n_reads,
                         int read_len) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_reads * (uint64_t)read_len) return;
  uint64_t r = g / read_len;
  int j = (int)(g - r * read_len);
  char c = in[r * read_len + (read_len - 1 - j)];
  char o;
  switch (c) {
    case 'A': o = 'T'; break;
    case 'C': o = 'G'; break;
    case 'G': o = 'C'; break;
    case 'T': o = 'A'; break;
    default: o = c; break;  // N stays N (rc_s.py maps only ACGT)
  }
  out[g] = o;
}

// variable-length reads (rc_s.py: D = {A:T, C:G, G:C, T:A, N:N}; anything else is a KeyError there)
__global__ void __launch_bounds__(kBlock)
    revcomp_var_kernel(const char* __restrict__ in, const uint64_t* __restrict__ offs, uint64_t n_reads,
                       uint64_t total, char* __restrict__ out, unsigned long long* bad) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t lo = 0, hi = n_reads;  // last read r with offs[r] <= g
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  const uint64_t b = __ldg(&offs[lo]), e = __ldg(&offs[lo + 1]);
  const char ch = in[g];
  char o;
  switch (ch) {
    case 'A': o = 'T'; break;
    case 'C': o = 'G'; break;
    case 'G': o = 'C'; break;
    case 'T': o = 'A'; break;
    case 'N': o = 'N'; break;
    default: o = ch; atomicAdd(bad, 1ull); break;
  }
  out[b + (e - 1 - g)] = o;
}

}  // namespace

void shn_revcomp_var_impl(shn_ctx* c, const char* in, const uint64_t* offs, uint64_t n_reads, uint64_t total,
                          char* out) {
  if (n_reads == 0 || total == 0) return;
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, c->stream));
  {
    ProfScope ps(c, "revcomp_reads");
    revcomp_var_kernel<<<shn_grid(total, kBlock), kBlock, 0, c->stream>>>(in, offs, n_reads, total, out, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, ctr, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h == 0, "read contains a character outside ACGTN (rc_s.py raises KeyError on it)");
}

void shn_synth_pairs_impl(shn_ctx* c, const uint8_t* tx, const uint64_t* tx_offs, const uint64_t* thr,
                          uint64_t n_tx, uint64_t n_pairs, uint64_t first_pair, uint64_t seed,
                          int read_len, int frag_len, uint32_t err_thr, char* m1, char* m2) {
  SHN_CHECK(read_len > 0 && frag_len >= read_len, "bad read/fragment length");
  if (n_pairs == 0) return;
  ProfScope ps(c, "synth_pairs");
  synth_pairs_kernel<<<shn_grid(n_pairs * read_len, kBlock), kBlock, 0, c->stream>>>(
      tx, tx_offs, thr, n_tx, n_pairs, first_pair, seed, read_len, frag_len, err_thr, m1, m2);
  KERNEL_CHECK();
}

void shn_revcomp_reads_impl(shn_ctx* c, const char* in, char* out, uint64_t n_reads, int read_len) {
  if (n_reads == 0) return;
  ProfScope ps(c, "revcomp_reads");
  revcomp_reads_kernel<<<shn_grid(n_reads * read_len, kBlock), kBlock, 0, c->stream>>>(in, out, n_reads,
                                                                                      read_len);
  KERNEL_CHECK();
}

#endif  // !SHN_WIDE

namespace SHN_NS {
namespace {

constexpr int kCBlock = 256;

// one thread per window; counters: [0]=new keys [1]=table full
__global__ void __launch_bounds__(kCBlock)
    count_windows_kernel(ShnTableView t, const char* __restrict__ reads, uint64_t n_reads,
                         int read_len, int k1, unsigned long long* counters) {
  const int wins = read_len - k1 + 1;
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, full = 0;
  if (g < n_reads * (uint64_t)wins) {
    uint64_t r = g / wins;
    int w = (int)(g - r * wins);
    const char* p = reads + r * read_len + w;
    shn_key_t key = 0;
    bool ok = true;
    for (int j = 0; j < k1; ++j) {
      uint32_t code = shn_code_of_strict((uint8_t)__ldg(&p[j]));
      ok &= code < 4;
      key = (key << 2) | (shn_key_t)(code & 3u);
    }
    if (ok && key != SHN_EMPTY) {  // (the all-T K1-mer of 32/64 bases is the table's empty marker;
                                   //  it is low-complexity and dropped by load_kmers anyway)
      uint64_t slot = table_upsert_slot(t, key, &n_new);
      if (slot == ~0ull)
        full = 1;
      else
        atomicAdd(&t.slots[slot].weight, 1u);
    }
  }
  int t_new = __syncthreads_count(n_new), t_full = __syncthreads_count(full);
  if (threadIdx.x == 0) {
    if (t_new) atomicAdd(&counters[0], (unsigned long long)t_new);
    if (t_full) atomicAdd(&counters[1], (unsigned long long)t_full);
  }
}

__global__ void __launch_bounds__(kCBlock) count_clear_kernel(ShnSlot* slots, uint64_t n_slots) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n_slots; i += stride) table_store_empty(slots, i, 0u);
}

// compaction of occupied slots: ASCII-order sort words of the key, count, running index
__global__ void __launch_bounds__(kCBlock)
    count_compact_kernel(const ShnSlot* __restrict__ slots, uint64_t n_slots, uint32_t min_count,
                         uint64_t* __restrict__ ord_lo, uint64_t* __restrict__ ord_hi,
                         uint32_t* __restrict__ counts, uint32_t* __restrict__ iota,
                         unsigned long long* cursor) {
  __shared__ unsigned long long block_base;
  __shared__ int warp_off[kCBlock / 32];
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  shn_key_t key = SHN_EMPTY;
  uint32_t wz = 0, wi = 0;
  if (i < n_slots) table_load_slot(slots, i, &key, &wz, &wi);
  bool occ = key != SHN_EMPTY && (wz & SHN_WEIGHT_MASK) >= min_count;   // jellyfish dump -L min_count
  unsigned b = __ballot_sync(0xFFFFFFFFu, occ);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_off[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kCBlock / 32; ++w) {
      int c = warp_off[w];
      warp_off[w] = tot;
      tot += c;
    }
    block_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (occ) {
    uint64_t o = block_base + warp_off[warp] + __popc(b & ((1u << lane) - 1u));
    shn_key_t ok = shn_ascii_order_key(key);
    ord_lo[o] = (uint64_t)ok;
#ifdef SHN_WIDE
    ord_hi[o] = (uint64_t)(ok >> 64);
#endif
    counts[o] = wz & SHN_WEIGHT_MASK;
    iota[o] = (uint32_t)o;
  }
}

__global__ void __launch_bounds__(kCBlock)
    gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, uint64_t n,
                      uint64_t* __restrict__ dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

// final order: keys back in table code order (the pair swap is an involution), counts permuted
__global__ void __launch_bounds__(kCBlock)
    count_emit_kernel(const uint64_t* __restrict__ ord_lo, const uint64_t* __restrict__ ord_hi,
                      const uint32_t* __restrict__ counts, const uint32_t* __restrict__ perm, uint64_t n,
                      uint64_t* __restrict__ keys_out, uint32_t* __restrict__ counts_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t p = perm[i];
#ifdef SHN_WIDE
  shn_key_t k = ((u128)ord_hi[p] << 64) | ord_lo[p];
#else
  shn_key_t k = ord_lo[p];
#endif
  shn_store_key(keys_out, i, shn_ascii_order_key(k));
  counts_out[i] = counts[p];
}

struct CountState {
  DevBuf table;              // counting table (weight word = count), alive between begin and finish
  uint64_t nb = 0;
  int k1 = 0;
  DevBuf keys, counts;       // result of the last finish
};

void count_state_free(shn_ctx* c) {
  delete static_cast<CountState*>(c->count_state);
  c->count_state = nullptr;
}

CountState* count_state_of(shn_ctx* c) {
  if (c->count_state && c->count_free != &count_state_free) shn_count_free(c);
  if (!c->count_state) {
    c->count_state = new CountState();
    c->count_free = &count_state_free;
  }
  return static_cast<CountState*>(c->count_state);
}

// last segment index s with offs[s] <= g
__device__ __forceinline__ uint64_t segment_of(const uint64_t* __restrict__ offs, uint64_t n, uint64_t g) {
  uint64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// variable-length reads: one thread per base position = window start
__global__ void __launch_bounds__(kCBlock)
    count_windows_var_kernel(ShnTableView t, const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                             uint64_t n_reads, uint64_t total, int k1, unsigned long long* counters) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, full = 0;
  if (g < total) {
    const uint64_t r = segment_of(offs, n_reads, g);
    if (g + k1 <= __ldg(&offs[r + 1])) {
      shn_key_t key = 0;
      bool ok = true;
      for (int j = 0; j < k1; ++j) {
        uint32_t code = shn_code_of_strict((uint8_t)__ldg(&bases[g + j]));
        ok &= code < 4;
        key = (key << 2) | (shn_key_t)(code & 3u);
      }
      if (ok && key != SHN_EMPTY) {
        uint64_t slot = table_upsert_slot(t, key, &n_new);
        if (slot == ~0ull)
          full = 1;
        else
          atomicAdd(&t.slots[slot].weight, 1u);
      }
    }
  }
  int t_new = __syncthreads_count(n_new), t_full = __syncthreads_count(full);
  if (threadIdx.x == 0) {
    if (t_new) atomicAdd(&counters[0], (unsigned long long)t_new);
    if (t_full) atomicAdd(&counters[1], (unsigned long long)t_full);
  }
}

void count_check_full(shn_ctx* c, unsigned long long* h) {
  CUDA_CHECK(cudaMemcpyAsync(h, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h[1] == 0, "k-mer counting table full: raise expected_distinct");
}

}  // namespace

// jellyfish count (shannon.py:439): the counting table lives in the context from count_begin to
// count_finish, so reads can be added in chunks (only one chunk of ASCII has to be resident)
void count_begin(shn_ctx* c, int k1, uint64_t expected_distinct) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1 && (SHN_KEY_WORDS == 1 || k1 > 32), "k1 out of range for this key width");
  CountState* st = count_state_of(c);
  st->keys.release();
  st->counts.release();
  const uint64_t nb = std::max<uint64_t>(256, (2 * expected_distinct + SHN_BSLOTS - 1) / SHN_BSLOTS);
  const uint64_t bytes = nb * SHN_BSLOTS * sizeof(ShnSlot);
  size_t fr = 0, tot = 0;
  CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
  SHN_CHECK(bytes <= st->table.bytes + fr + c->pool.cached_bytes,
            "k-mer counting table of " + std::to_string(bytes >> 20) +
                " MB does not fit the free device memory: lower expected_distinct or count in more shards");
  st->table.reserve(bytes);
  st->nb = nb;
  st->k1 = k1;
  c->counters.reserve(64 * sizeof(unsigned long long));
  CUDA_CHECK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
  ProfScope ps(c, "count_clear");
  unsigned grid =
      (unsigned)std::min<uint64_t>((nb * SHN_BSLOTS + kCBlock - 1) / kCBlock, (uint64_t)c->sm_count * 32);
  count_clear_kernel<<<grid, kCBlock, 0, c->stream>>>(st->table.as<ShnSlot>(), nb * SHN_BSLOTS);
  KERNEL_CHECK();
}

// fixed-length reads (the synthetic generators): n_reads x read_len bytes
void count_add_fixed(shn_ctx* c, const char* d_reads, uint64_t n_reads, int read_len) {
  CountState* st = count_state_of(c);
  SHN_CHECK(st->nb > 0, "shn_count_begin has not been called");
  if (read_len < st->k1 || n_reads == 0) return;
  const uint64_t nw = n_reads * (uint64_t)(read_len - st->k1 + 1);
  ProfScope ps(c, "count_windows");
  count_windows_kernel<<<shn_grid(nw, kCBlock), kCBlock, 0, c->stream>>>(
      ShnTableView{st->table.as<ShnSlot>(), st->nb}, d_reads, n_reads, read_len, st->k1,
      c->counters.as<unsigned long long>());
  KERNEL_CHECK();
}

// variable-length reads: concatenated bases + n_reads+1 offsets (device pointers).  Returns after
// the chunk has been counted: the caller may re-use its buffers.
void count_add_var(shn_ctx* c, const char* d_bases, const uint64_t* d_offs, uint64_t n_reads,
                   uint64_t total_bases) {
  CountState* st = count_state_of(c);
  SHN_CHECK(st->nb > 0, "shn_count_begin has not been called");
  if (n_reads == 0 || total_bases == 0) return;
  {
    ProfScope ps(c, "count_windows");
    count_windows_var_kernel<<<shn_grid(total_bases, kCBlock), kCBlock, 0, c->stream>>>(
        ShnTableView{st->table.as<ShnSlot>(), st->nb}, d_bases, d_offs, n_reads, total_bases, st->k1,
        c->counters.as<unsigned long long>());
    KERNEL_CHECK();
  }
  unsigned long long h[2];
  count_check_full(c, h);
}

// jellyfish dump -c -t -L min_count (shannon.py:441): (k-mer, count) in ascending ASCII order of the
// k-mer, the documented order of oracle/kmer_count.py; device arrays owned by the context
void count_finish(shn_ctx* c, uint32_t min_count, uint64_t** keys_dev, uint32_t** counts_dev,
                  uint64_t* n_distinct) {
  CountState* st = count_state_of(c);
  SHN_CHECK(st->nb > 0, "shn_count_begin has not been called");
  const int k1 = st->k1;
  const uint64_t nb = st->nb;
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  unsigned long long h[2];
  count_check_full(c, h);
  const uint64_t n_all = h[0];
  SHN_CHECK(n_all < 0xFFFFFFFFull, "more than 2^32-1 distinct K1-mers");
  const uint64_t n1 = std::max<uint64_t>(n_all, 1);
  DevBuf ord_lo, ord_hi, cnt, iota, lo_s, perm;
  uint64_t n = 0;
  if (n_all) {
    ord_lo.reserve(n1 * 8);
    ord_hi.reserve(SHN_KEY_WORDS == 2 ? n1 * 8 : 8);
    cnt.reserve(n1 * 4);
    iota.reserve(n1 * 4);
    CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, c->stream));
    {
      ProfScope ps(c, "count_compact");
      count_compact_kernel<<<shn_grid(nb * SHN_BSLOTS, kCBlock), kCBlock, 0, c->stream>>>(
          st->table.as<ShnSlot>(), nb * SHN_BSLOTS, min_count, ord_lo.as<uint64_t>(), ord_hi.as<uint64_t>(),
          cnt.as<uint32_t>(), iota.as<uint32_t>(), ctr);
      KERNEL_CHECK();
    }
    unsigned long long kept = 0;
    CUDA_CHECK(cudaMemcpyAsync(&kept, ctr, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    n = kept;
  }
  st->table.release();  // the counting table is the largest buffer: free it before the sort buffers
  st->nb = 0;
  st->keys.reserve(std::max<uint64_t>(n, 1) * 8 * SHN_KEY_WORDS);
  st->counts.reserve(std::max<uint64_t>(n, 1) * 4);
  if (n) {
    lo_s.reserve(n * 8);
    perm.reserve(n * 4);
    ProfScope ps(c, "count_sort", 4);
    // LSD radix over the ASCII-order key: low word first, then (stable) the high word
    size_t tb = 0;
    const int lo_bits = k1 > 32 ? 64 : 2 * k1;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, ord_lo.as<uint64_t>(), lo_s.as<uint64_t>(),
                                               iota.as<uint32_t>(), perm.as<uint32_t>(), (int64_t)n, 0,
                                               lo_bits, c->stream));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, ord_lo.as<uint64_t>(), lo_s.as<uint64_t>(),
                                               iota.as<uint32_t>(), perm.as<uint32_t>(), (int64_t)n, 0,
                                               lo_bits, c->stream));
    uint32_t* final_perm = perm.as<uint32_t>();
#ifdef SHN_WIDE
    DevBuf hi_g, hi_s, perm2;
    hi_g.reserve(n * 8);
    hi_s.reserve(n * 8);
    perm2.reserve(n * 4);
    gather_u64_kernel<<<shn_grid(n, kCBlock), kCBlock, 0, c->stream>>>(ord_hi.as<uint64_t>(),
                                                                     perm.as<uint32_t>(), n,
                                                                     hi_g.as<uint64_t>());
    KERNEL_CHECK();
    tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, hi_g.as<uint64_t>(), hi_s.as<uint64_t>(),
                                               perm.as<uint32_t>(), perm2.as<uint32_t>(), (int64_t)n, 0,
                                               2 * (k1 - 32), c->stream));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, hi_g.as<uint64_t>(), hi_s.as<uint64_t>(),
                                               perm.as<uint32_t>(), perm2.as<uint32_t>(), (int64_t)n, 0,
                                               2 * (k1 - 32), c->stream));
    final_perm = perm2.as<uint32_t>();
#endif
    count_emit_kernel<<<shn_grid(n, kCBlock), kCBlock, 0, c->stream>>>(
        ord_lo.as<uint64_t>(), ord_hi.as<uint64_t>(), cnt.as<uint32_t>(), final_perm, n,
        st->keys.as<uint64_t>(), st->counts.as<uint32_t>());
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  *keys_dev = st->keys.as<uint64_t>();
  *counts_dev = st->counts.as<uint32_t>();
  *n_distinct = n;
}

// the three steps in one call for fixed-length read arrays (bench / test inputs)
void count_k1mers(shn_ctx* c, const char* const* arrays, const uint64_t* n_reads, int n_arrays,
                  int read_len, int k1, uint64_t expected_distinct, uint64_t** keys_dev,
                  uint32_t** counts_dev, uint64_t* n_distinct) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1 && read_len >= k1, "bad k1 / read length");
  uint64_t total_windows = 0;
  for (int a = 0; a < n_arrays; ++a) total_windows += n_reads[a] * (uint64_t)(read_len - k1 + 1);
  count_begin(c, k1, std::min(expected_distinct, total_windows));
  for (int a = 0; a < n_arrays; ++a) count_add_fixed(c, arrays[a], n_reads[a], read_len);
  count_finish(c, 1, keys_dev, counts_dev, n_distinct);
}

}  // namespace SHN_NS
