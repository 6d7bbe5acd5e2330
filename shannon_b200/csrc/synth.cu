// Inputs of the hot path produced on the device (SURVEY 8f rows f1/f2 + the synthetic configs):
//  - shn_synth_pairs   : seeded synthetic read pairs (twin of shannon_b200/synth.py::make_pairs)
//  - shn_revcomp_reads : RC doubling of fixed-length reads (rc_gnu.py / rc_s.py, shannon.py:395-424)
//  - shn_count_k1mers  : jellyfish count/dump stand-in (shannon.py:439-441), ASCII-sorted output
#include <cub/cub.cuh>

#include "common.cuh"
#include "table_dev.cuh"

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

// one thread per window; counters: [0]=new keys [1]=table full
__global__ void __launch_bounds__(kBlock)
    count_windows_kernel(ShnTableView t, const char* __restrict__ reads, uint64_t n_reads,
                         int read_len, int k1, unsigned long long* counters) {
  const int wins = read_len - k1 + 1;
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, full = 0;
  if (g < n_reads * (uint64_t)wins) {
    uint64_t r = g / wins;
    int w = (int)(g - r * wins);
    const char* p = reads + r * read_len + w;
    uint64_t key = 0;
    bool ok = true;
    for (int j = 0; j < k1; ++j) {
      uint32_t code = shn_code_of_strict((uint8_t)__ldg(&p[j]));
      ok &= code < 4;
      key = (key << 2) | (code & 3u);
    }
    if (ok) {
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

__global__ void __launch_bounds__(kBlock)
    count_clear_kernel(ShnSlot* slots, uint64_t n_slots) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
  for (; i < n_slots; i += stride) reinterpret_cast<uint4*>(slots)[i] = v;
}

// compaction of occupied slots into (ascii-order sort key, count), block-aggregated append
__global__ void __launch_bounds__(kBlock)
    count_compact_kernel(const ShnSlot* __restrict__ slots, uint64_t n_slots,
                         uint64_t* __restrict__ skeys, uint32_t* __restrict__ counts,
                         unsigned long long* cursor) {
  __shared__ unsigned long long block_base;
  __shared__ int warp_off[kBlock / 32];
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0);
  if (i < n_slots) v = __ldg(reinterpret_cast<const uint4*>(slots) + i);
  uint64_t key = ((uint64_t)v.y << 32) | v.x;
  bool occ = key != SHN_EMPTY_KEY;
  unsigned b = __ballot_sync(0xFFFFFFFFu, occ);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_off[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kBlock / 32; ++w) {
      int c = warp_off[w];
      warp_off[w] = tot;
      tot += c;
    }
    block_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (occ) {
    uint64_t o = block_base + warp_off[warp] + __popc(b & ((1u << lane) - 1u));
    skeys[o] = shn_ascii_order_key(key);
    counts[o] = v.z & SHN_WEIGHT_MASK;
  }
}

__global__ void __launch_bounds__(kBlock)
    unorder_keys_kernel(uint64_t* __restrict__ keys, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = shn_ascii_order_key(keys[i]);  // the pair swap is an involution
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

struct CountState {
  DevBuf keys, counts;
};
static std::map<shn_ctx*, CountState*> g_count_state;

void shn_count_free(shn_ctx* c) {
  auto it = g_count_state.find(c);
  if (it != g_count_state.end()) {
    delete it->second;
    g_count_state.erase(it);
  }
}

void shn_count_k1mers_impl(shn_ctx* c, const char* const* arrays, const uint64_t* n_reads,
                           int n_arrays, int read_len, int k1, uint64_t expected_distinct,
                           uint64_t** keys_dev, uint32_t** counts_dev, uint64_t* n_distinct) {
  SHN_CHECK(k1 >= 1 && k1 <= 32 && read_len >= k1, "bad k1 / read length");
  uint64_t total_windows = 0;
  for (int a = 0; a < n_arrays; ++a) total_windows += n_reads[a] * (uint64_t)(read_len - k1 + 1);
  uint64_t nb = std::max<uint64_t>(256, (std::min(expected_distinct, total_windows) + 1) / 2);
  DevBuf table;
  table.reserve(nb * SHN_BSLOTS * sizeof(ShnSlot));
  ShnTableView view{table.as<ShnSlot>(), nb};
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), c->stream));
  {
    ProfScope ps(c, "count_clear");
    unsigned grid =
        (unsigned)std::min<uint64_t>((nb * SHN_BSLOTS + kBlock - 1) / kBlock, (uint64_t)c->sm_count * 32);
    count_clear_kernel<<<grid, kBlock, 0, c->stream>>>(view.slots, nb * SHN_BSLOTS);
    KERNEL_CHECK();
  }
  for (int a = 0; a < n_arrays; ++a) {
    uint64_t nw = n_reads[a] * (uint64_t)(read_len - k1 + 1);
    if (nw == 0) continue;
    ProfScope ps(c, "count_windows");
    count_windows_kernel<<<shn_grid(nw, kBlock), kBlock, 0, c->stream>>>(view, arrays[a], n_reads[a],
                                                                         read_len, k1, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h[2];
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h[1] == 0, "k-mer counting table full: raise expected_distinct");
  uint64_t n = h[0];
  CountState*& st = g_count_state[c];
  if (!st) st = new CountState();
  DevBuf skeys, cnt;
  skeys.reserve(std::max<uint64_t>(n, 1) * 8);
  cnt.reserve(std::max<uint64_t>(n, 1) * 4);
  st->keys.reserve(std::max<uint64_t>(n, 1) * 8);
  st->counts.reserve(std::max<uint64_t>(n, 1) * 4);
  if (n) {
    CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, c->stream));
    {
      ProfScope ps(c, "count_compact");
      count_compact_kernel<<<shn_grid(nb * SHN_BSLOTS, kBlock), kBlock, 0, c->stream>>>(
          view.slots, nb * SHN_BSLOTS, skeys.as<uint64_t>(), cnt.as<uint32_t>(), ctr);
      KERNEL_CHECK();
    }
    ProfScope ps(c, "count_sort", 2);
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, skeys.as<uint64_t>(), st->keys.as<uint64_t>(),
                                               cnt.as<uint32_t>(), st->counts.as<uint32_t>(), (int64_t)n,
                                               0, 2 * k1, c->stream));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, skeys.as<uint64_t>(),
                                               st->keys.as<uint64_t>(), cnt.as<uint32_t>(),
                                               st->counts.as<uint32_t>(), (int64_t)n, 0, 2 * k1,
                                               c->stream));
    unorder_keys_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(st->keys.as<uint64_t>(), n);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  *keys_dev = st->keys.as<uint64_t>();
  *counts_dev = st->counts.as<uint32_t>();
  *n_distinct = n;
}
