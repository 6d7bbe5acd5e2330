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

}  // namespace

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
    count_compact_kernel(const ShnSlot* __restrict__ slots, uint64_t n_slots,
                         uint64_t* __restrict__ ord_lo, uint64_t* __restrict__ ord_hi,
                         uint32_t* __restrict__ counts, uint32_t* __restrict__ iota,
                         unsigned long long* cursor) {
  __shared__ unsigned long long block_base;
  __shared__ int warp_off[kCBlock / 32];
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  shn_key_t key = SHN_EMPTY;
  uint32_t wz = 0, wi = 0;
  if (i < n_slots) table_load_slot(slots, i, &key, &wz, &wi);
  bool occ = key != SHN_EMPTY;
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
  DevBuf keys, counts;
};

void count_state_free(shn_ctx* c) {
  delete static_cast<CountState*>(c->count_state);
  c->count_state = nullptr;
}

}  // namespace

void count_k1mers(shn_ctx* c, const char* const* arrays, const uint64_t* n_reads, int n_arrays,
                  int read_len, int k1, uint64_t expected_distinct, uint64_t** keys_dev,
                  uint32_t** counts_dev, uint64_t* n_distinct) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1 && read_len >= k1, "bad k1 / read length");
  uint64_t total_windows = 0;
  for (int a = 0; a < n_arrays; ++a) total_windows += n_reads[a] * (uint64_t)(read_len - k1 + 1);
  uint64_t nb = std::max<uint64_t>(256, (2 * std::min(expected_distinct, total_windows) + SHN_BSLOTS - 1) /
                                            SHN_BSLOTS);
  DevBuf table;
  table.reserve(nb * SHN_BSLOTS * sizeof(ShnSlot));
  ShnTableView view{table.as<ShnSlot>(), nb};
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), c->stream));
  {
    ProfScope ps(c, "count_clear");
    unsigned grid =
        (unsigned)std::min<uint64_t>((nb * SHN_BSLOTS + kCBlock - 1) / kCBlock, (uint64_t)c->sm_count * 32);
    count_clear_kernel<<<grid, kCBlock, 0, c->stream>>>(view.slots, nb * SHN_BSLOTS);
    KERNEL_CHECK();
  }
  for (int a = 0; a < n_arrays; ++a) {
    uint64_t nw = n_reads[a] * (uint64_t)(read_len - k1 + 1);
    if (nw == 0) continue;
    ProfScope ps(c, "count_windows");
    count_windows_kernel<<<shn_grid(nw, kCBlock), kCBlock, 0, c->stream>>>(view, arrays[a], n_reads[a],
                                                                           read_len, k1, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h[2];
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h[1] == 0, "k-mer counting table full: raise expected_distinct");
  uint64_t n = h[0];
  SHN_CHECK(n < 0xFFFFFFFFull, "more than 2^32-1 distinct K1-mers");
  if (c->count_state && c->count_free != &count_state_free) shn_count_free(c);
  if (!c->count_state) {
    c->count_state = new CountState();
    c->count_free = &count_state_free;
  }
  CountState* st = static_cast<CountState*>(c->count_state);
  const uint64_t n1 = std::max<uint64_t>(n, 1);
  st->keys.reserve(n1 * 8 * SHN_KEY_WORDS);
  st->counts.reserve(n1 * 4);
  if (n) {
    DevBuf ord_lo, ord_hi, cnt, iota, lo_s, perm;
    ord_lo.reserve(n1 * 8);
    ord_hi.reserve(SHN_KEY_WORDS == 2 ? n1 * 8 : 8);
    cnt.reserve(n1 * 4);
    iota.reserve(n1 * 4);
    CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, c->stream));
    {
      ProfScope ps(c, "count_compact");
      count_compact_kernel<<<shn_grid(nb * SHN_BSLOTS, kCBlock), kCBlock, 0, c->stream>>>(
          view.slots, nb * SHN_BSLOTS, ord_lo.as<uint64_t>(), ord_hi.as<uint64_t>(), cnt.as<uint32_t>(),
          iota.as<uint32_t>(), ctr);
      KERNEL_CHECK();
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    table.release();  // the counting table is the largest buffer: free it before the sort buffers
    lo_s.reserve(n1 * 8);
    perm.reserve(n1 * 4);
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
    hi_g.reserve(n1 * 8);
    hi_s.reserve(n1 * 8);
    perm2.reserve(n1 * 4);
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

}  // namespace SHN_NS
