// a1/a2: K1-mer -> weight table (load_kmers + lowComplexity, extension_correction.py:202-221,
// 142-149) as an open-addressing hash table in HBM.  Compiled twice (common.cuh): 64-bit keys into
// namespace `narrow`, 128-bit keys (-DSHN_WIDE) into namespace `wide`.
//
// Layout: 64-byte buckets {key, weight u32, first_idx u32} x 4 (2 for wide keys); a probe moves
// one 64-byte DRAM burst.  Bucket = mulhi(hash(key), n_buckets), linear probing over buckets with
// an overflow flag per bucket (table_dev.cuh).  Load factor <= 0.5.
//
// Algorithmic bytes: insert = 8 B key + 4 B count streamed + one bucket
// read-modify-write (2 x 64 B) = 140 B per input line; lookup = 8 B + 64 B + 5 B out = 77 B.
#include <cub/cub.cuh>

#include "common.cuh"
#include "impls.h"
#include "table_dev.cuh"

namespace SHN_NS {
namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) table_clear_kernel(ShnSlot* slots, uint64_t n_slots) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // key = EMPTY, weight = 0, first_idx = +inf for atomicMin; coalesced 16/32-byte stores
  for (; i < n_slots; i += stride) table_store_empty(slots, i, 0xFFFFFFFFu);
}

// counters: [0]=new keys [1]=low-complexity lines [3]=bad key / index overflow
__global__ void __launch_bounds__(kBlock)
    table_insert_kernel(ShnTableView t, const uint64_t* __restrict__ keys,
                        const uint32_t* __restrict__ counts, const uint32_t* __restrict__ line_idx,
                        uint64_t n, uint64_t first_line, int k1, int ds,
                        unsigned long long* counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, n_low = 0, n_bad = 0;
  if (i < n) {
    shn_key_t key = shn_load_key(keys, i);
    uint32_t w = counts[i];
    // explicit line indices (sharded build: the global input line of every routed key)
    uint64_t base_idx = line_idx ? (uint64_t)line_idx[i] : (ds ? 2 * (first_line + i) : first_line + i);
    if ((key & ~shn_key_mask(k1)) || base_idx + 1 >= 0xFFFFFFFFull || w >= SHN_WEIGHT_MASK) {
      n_bad = 1;
    } else if (shn_low_complexity(key, k1)) {  // rc(kmer) is low-complexity iff kmer is
      n_low = 1;
    } else {
      const int reps = ds ? 2 : 1;
      for (int r = 0; r < reps; ++r) {
        shn_key_t kk = r == 0 ? key : shn_revcomp(key, k1);
        uint32_t old = 0;
        uint64_t slot = table_insert_add(t, kk, w, (uint32_t)(base_idx + r), &n_new, &old);
        if (slot == ~0ull) {
          n_bad = 1;
          break;
        }
        if ((uint64_t)old + w >= (uint64_t)SHN_WEIGHT_MASK) n_bad = 1;
      }
    }
  }
  // one atomic per block and counter instead of one per thread
  int tot_new = __syncthreads_count(n_new & 1) + 2 * __syncthreads_count(n_new >> 1);
  int tot_low = __syncthreads_count(n_low);
  int tot_bad = __syncthreads_count(n_bad);
  if (threadIdx.x == 0) {
    if (tot_new) atomicAdd(&counters[0], (unsigned long long)tot_new);
    if (tot_low) atomicAdd(&counters[1], (unsigned long long)tot_low);
    if (tot_bad) atomicAdd(&counters[3], (unsigned long long)tot_bad);
  }
}

__global__ void __launch_bounds__(kBlock)
    table_lookup_kernel(ShnTableView t, const uint64_t* __restrict__ keys, uint64_t n,
                        uint32_t* __restrict__ weights, uint8_t* __restrict__ found) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t wraw = 0;
  uint64_t slot = table_find(t, shn_load_key(keys, i), &wraw);
  uint32_t w = slot == ~0ull ? 0u : (wraw & SHN_WEIGHT_MASK);
  uint8_t f = slot == ~0ull ? 0 : 1;
  if (weights) weights[i] = w;
  if (found) found[i] = f;
}

struct OccupiedSlot {
  const ShnSlot* slots;
  __device__ bool operator()(uint64_t i) const { return slots[i].key != SHN_EMPTY; }
};

__global__ void __launch_bounds__(kBlock)
    table_dump_gather_kernel(const ShnSlot* __restrict__ slots, const uint64_t* __restrict__ sel,
                             uint64_t n, uint64_t* keys, uint32_t* weights, uint32_t* idx,
                             uint32_t* order) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ShnSlot s = slots[sel[i]];
  shn_store_key(keys, i, s.key);
  weights[i] = s.weight & SHN_WEIGHT_MASK;
  idx[i] = s.idx;
  order[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    table_dump_permute_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ weights,
                              const uint32_t* __restrict__ order, uint64_t n, uint64_t* keys_out,
                              uint32_t* weights_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t o = order[i];
  shn_store_key(keys_out, i, shn_load_key(keys, o));
  weights_out[i] = weights[o];
}

__global__ void __launch_bounds__(kBlock)
    pack_kmers_kernel(const char* __restrict__ ascii, uint64_t n, int k1, uint64_t* __restrict__ keys,
                      unsigned long long* counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < n) {
    const char* p = ascii + i * (uint64_t)k1;
    shn_key_t x = 0;
    for (int j = 0; j < k1; ++j) {
      uint32_t code = shn_code_of((uint8_t)__ldg(&p[j]));  // kmer.upper(), :214
      bad |= code > 3;
      x = (x << 2) | (shn_key_t)(code & 3u);
    }
    shn_store_key(keys, i, x);
  }
  int tot = __syncthreads_count(bad);
  if (threadIdx.x == 0 && tot) atomicAdd(&counters[0], (unsigned long long)tot);
}

}  // namespace

void pack_kmers(shn_ctx* c, const char* d_ascii, uint64_t n, int k1, uint64_t* d_keys) {
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, c->stream));
  {
    ProfScope ps(c, "pack_kmers");
    pack_kmers_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(d_ascii, n, k1, d_keys, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, ctr, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h == 0, "k-mer contains a character outside ACGT");
}

// The build in three parts so that a host-resident input can be copied in chunks on the copy
// stream while earlier chunks are being inserted (shn_table_build, api.cu).
void table_begin(shn_ctx* c, uint64_t n, int k1, int double_stranded) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1 && (SHN_KEY_WORDS == 1 || k1 > 32), "k1 out of range for this key width");
  uint64_t items = n * (double_stranded ? 2 : 1);
  SHN_CHECK(items < 0xFFFFFFFEull, "more than 2^32-2 input K1-mers per table");
  // slots >= 2 * items  (load factor <= 0.5)
  // slots = items / load factor.  0.5 by default: at 0.4 the walks need fewer probe continuations
  // (walk stage 110 -> 103 ms at 10 M pairs) but every pass over the slots grows (uf_edges 50 -> 55 ms):
  // no net gain, measured.  SHN_TABLE_LOAD in 0.1 .. 0.9.
  const char* envl = getenv("SHN_TABLE_LOAD");
  double load = envl ? atof(envl) : 0.5;
  if (!(load >= 0.1 && load <= 0.9)) load = 0.5;
  uint64_t n_buckets = items < 1024 ? 1024 / SHN_BSLOTS
                                    : (uint64_t)((double)items / load + SHN_BSLOTS - 1) / SHN_BSLOTS;
  // minimizer-clustered regions (common.cuh) once the table spans at least a few of them
  const char* envs = getenv("SHN_REGION_SHIFT");
  // 2^17 buckets = 8 MB per region: measured at 10 M pairs (uf_edges 94 -> 57 ms); 256 KB regions
  // have a load spread of +-30 % and triple the length of the walks' probe sequences
  c->region_shift = envs ? std::max(4, std::min(24, atoi(envs))) : 17;
  const uint64_t region = 1ull << c->region_shift;
  const char* envr = getenv("SHN_TABLE_REGIONS");  // 0 = plain hashing (A/B measurements, tests)
  const bool regions = (!envr || atoi(envr) != 0) && k1 > kRegionM && n_buckets >= 4 * region;
  c->n_regions = 0;
  if (regions) {
    n_buckets = (n_buckets + region - 1) / region * region;
    c->n_regions = (uint32_t)(n_buckets >> c->region_shift);
  }
  c->table.reserve(n_buckets * SHN_BSLOTS * sizeof(ShnSlot));
  c->n_buckets = n_buckets;
  c->k1 = k1;
  c->n_items = items;
  c->gline_dev = nullptr;  // set again by table_build_records
  c->explicit_idx = false;
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), c->stream));
  ProfScope ps(c, "table_clear");
  unsigned grid = (unsigned)std::min<uint64_t>((n_buckets * SHN_BSLOTS + kBlock - 1) / kBlock,
                                               (uint64_t)c->sm_count * 32);
  table_clear_kernel<<<grid, kBlock, 0, c->stream>>>(c->table.as<ShnSlot>(), n_buckets * SHN_BSLOTS);
  KERNEL_CHECK();
}

// lines [first_line, first_line + n) of the input; d_keys/d_counts point at this chunk
void table_insert_chunk(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts,
                        const uint32_t* d_line_idx, uint64_t n, uint64_t first_line,
                        int double_stranded) {
  if (n == 0) return;
  if (d_line_idx) c->explicit_idx = true;
  ProfScope ps(c, "table_insert");
  table_insert_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(
      table_view(c), d_keys, d_counts, d_line_idx, n, first_line, c->k1, double_stranded,
      c->counters.as<unsigned long long>());
  KERNEL_CHECK();
}

void table_finish(shn_ctx* c) {
  unsigned long long h[4];
  CUDA_CHECK(cudaMemcpyAsync(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  SHN_CHECK(h[3] == 0, "table build: key wider than 2*k1 bits, a K1-mer weight above 2^30-2, or input index overflow");
  c->n_distinct = h[0];
  c->n_lowcomplexity = h[1];
}

void table_build(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n, int k1,
                 int double_stranded, const uint32_t* d_line_idx) {
  SHN_CHECK(!(double_stranded && d_line_idx), "explicit line indices exclude double_stranded");
  table_begin(c, n, k1, double_stranded);
  table_insert_chunk(c, d_keys, d_counts, d_line_idx, n, 0, double_stranded);
  table_finish(c);
}

void table_lookup(shn_ctx* c, const uint64_t* d_keys, uint64_t n, uint32_t* d_weights,
                  uint8_t* d_found) {
  SHN_CHECK(c->n_buckets > 0, "no table built");
  if (n == 0) return;
  ProfScope ps(c, "table_lookup");
  table_lookup_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(table_view(c), d_keys, n,
                                                                     d_weights, d_found);
  KERNEL_CHECK();
}

// dump sorted by first-occurrence index ------------------------------------------------------
void table_dump(shn_ctx* c, uint64_t* h_keys, uint32_t* h_weights, uint32_t* h_idx) {
  SHN_CHECK(c->n_buckets > 0, "no table built");
  uint64_t n_slots = c->n_buckets * SHN_BSLOTS, n = c->n_distinct;
  if (n == 0) return;
  const uint64_t kb = 8 * SHN_KEY_WORDS;
  DevBuf sel, nsel, keys, w, idx, order, keys2, w2, idx2, order2;
  sel.reserve(n * 8);
  nsel.reserve(8);
  keys.reserve(n * kb);
  w.reserve(n * 4);
  idx.reserve(n * 4);
  order.reserve(n * 4);
  keys2.reserve(n * kb);
  w2.reserve(n * 4);
  idx2.reserve(n * 4);
  order2.reserve(n * 4);
  cub::CountingInputIterator<uint64_t> it(0);
  OccupiedSlot pred{c->table.as<ShnSlot>()};
  size_t tb = 0;
  SHN_CHECK(n_slots < 0x7FFFFFFFull, "table dump limited to < 2^31 slots");
  CUDA_CHECK(cub::DeviceSelect::If(nullptr, tb, it, sel.as<uint64_t>(), nsel.as<uint64_t>(),
                                   (int)n_slots, pred, c->stream));
  CUDA_CHECK(cub::DeviceSelect::If(c->tmp(tb), tb, it, sel.as<uint64_t>(), nsel.as<uint64_t>(),
                                   (int)n_slots, pred, c->stream));
  table_dump_gather_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(
      c->table.as<ShnSlot>(), sel.as<uint64_t>(), n, keys.as<uint64_t>(), w.as<uint32_t>(),
      idx.as<uint32_t>(), order.as<uint32_t>());
  KERNEL_CHECK();
  // sort the permutation by first-occurrence index, then permute keys and weights
  tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, idx.as<uint32_t>(), idx2.as<uint32_t>(),
                                             order.as<uint32_t>(), order2.as<uint32_t>(), (int)n, 0, 32,
                                             c->stream));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, idx.as<uint32_t>(), idx2.as<uint32_t>(),
                                             order.as<uint32_t>(), order2.as<uint32_t>(), (int)n, 0, 32,
                                             c->stream));
  table_dump_permute_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(
      keys.as<uint64_t>(), w.as<uint32_t>(), order2.as<uint32_t>(), n, keys2.as<uint64_t>(),
      w2.as<uint32_t>());
  KERNEL_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(h_keys, keys2.p, n * kb, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(h_weights, w2.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(h_idx, idx2.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace SHN_NS
