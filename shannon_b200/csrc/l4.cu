// a10-a12: K1-mer -> component map and read partition (kmers_for_component.py:186-205,239-305,
// 322-423,452-477).
//
// Component map: open-addressing table, 32-byte buckets of two 16-byte slots
// {key u64, comp0 u32, comp1 u32}; a K1-mer belongs to at most two components (its 'c' partition
// and the 'r2_c' twin of the repartition pass).  Weights for the per-component k1mer files live
// in a side array indexed by slot and are touched only by shn_l4_map_window_weights.
//
// Reads: 2-bit packed, MSB-first 32 bases per uint64 word, ceil(len/32) words per read.
// Assignment: 8 lanes per record (4 samples x 2 mates for 100 bp, K=24); each lane extracts one
// sampled K1-mer, probes one 32-byte sector; the union over the record is formed with shuffles.
//
// Algorithmic bytes per record: 32 B packed words + 8 B (offset,len) per read,
// 32 B per probe, 8 B per assignment written.
#include <cub/cub.cuh>

#include "common.cuh"
#include "impls.h"
#include "table_dev.cuh"

#include "reads.cuh"

namespace SHN_NS {

#ifdef SHN_WIDE
struct __align__(32) CompSlot {
  u128 key;
  uint32_t comp0, comp1;
  uint64_t pad;
};
#else
struct __align__(16) CompSlot {
  uint64_t key;
  uint32_t comp0, comp1;
};
#endif

struct L4State {
  DevBuf map;       // CompSlot[2*n_buckets]
  DevBuf map_w;     // uint32 weight per slot
  uint64_t n_buckets = 0;
  int k1 = 0;
  uint64_t n_keys = 0;
  DevBuf assign;    // uint64 entries (comp << 32 | record), sorted + unique after shn_l4_assign
  uint64_t n_assign = 0;
  DevBuf stage_a, stage_b, stage_c;
};

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kLenBad = 0x80000000u;

struct CompMapView {
  CompSlot* slots;
  uint64_t n_buckets;
  __device__ __forceinline__ uint64_t bucket_of(shn_key_t key) const {
    return __umul64hi(shn_key_hash(key), n_buckets);
  }
};

__global__ void __launch_bounds__(kBlock) map_clear_kernel(CompSlot* slots, uint32_t* w, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  for (; i < n; i += stride) {  // key = EMPTY, comp0 = comp1 = NONE
#ifdef SHN_WIDE
    reinterpret_cast<uint4*>(slots)[2 * i] = v;
    reinterpret_cast<uint4*>(slots)[2 * i + 1] = v;
#else
    reinterpret_cast<uint4*>(slots)[i] = v;
#endif
    w[i] = 0;
  }
}

// index of the contig that contains concatenated-base position g: last c with offs[c] <= g
__device__ __forceinline__ uint64_t find_segment(const uint64_t* __restrict__ offs, uint64_t n,
                                                 uint64_t g) {
  uint64_t lo = 0, hi = n;  // invariant offs[lo] <= g < offs[hi]
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// packs bases[p .. p+k) ; returns false if a non-ACGT character is met
__device__ __forceinline__ bool pack_window(const char* __restrict__ bases, uint64_t p, int k,
                                            shn_key_t* out) {
  shn_key_t x = 0;
  bool ok = true;
  for (int j = 0; j < k; ++j) {
    uint32_t code = shn_code_of_strict((uint8_t)__ldg(&bases[p + j]));
    ok &= code < 4;
    x = (x << 2) | (shn_key_t)(code & 3u);
  }
  *out = x;
  return ok;
}

// bucket of two slots (32 bytes = one 256-bit load for 64-bit keys, 64 bytes for 128-bit keys);
// *comps = comp0 | comp1 << 32 of the matching slot
__device__ __forceinline__ uint64_t map_find(const CompMapView& m, shn_key_t key, uint64_t* comps) {
  if (key == SHN_EMPTY) return ~0ull;
  uint64_t b = m.bucket_of(key);
  for (;;) {
#ifdef SHN_WIDE
    uint64_t w[8];
    shn_ld256_nc(m.slots + 2 * b, w);
    shn_ld256_nc(m.slots + 2 * b + 1, w + 4);
    const shn_key_t k0 = ((u128)w[1] << 64) | w[0], k1 = ((u128)w[5] << 64) | w[4];
    const uint64_t c0 = w[2], c1 = w[6];
#else
    uint64_t w[4];
    shn_ld256_nc(m.slots + 2 * b, w);
    const shn_key_t k0 = w[0], k1 = w[2];
    const uint64_t c0 = w[1], c1 = w[3];
#endif
    if (k0 == key) {
      *comps = c0;
      return 2 * b;
    }
    if (k1 == key) {
      *comps = c1;
      return 2 * b + 1;
    }
    if (k0 == SHN_EMPTY || k1 == SHN_EMPTY) return ~0ull;
    b = (b + 1 == m.n_buckets) ? 0 : b + 1;
  }
}

// counters: [0]=new keys [1]=non-ACGT windows [2]=more than two components for a K1-mer
// [3]=windows equal to the free-slot marker (poly-T of 32 / 64 bases)
__global__ void __launch_bounds__(kBlock)
    map_add_kernel(CompMapView m, const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                   const uint32_t* __restrict__ comp_of, uint64_t n_contigs, uint64_t total_bases,
                   int k1, int is_codes, unsigned long long* counters) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, n_bad = 0, n_over = 0, n_empty = 0;
  if (g < total_bases) {
    uint64_t c = find_segment(offs, n_contigs, g);
    uint64_t end = __ldg(&offs[c + 1]);
    uint32_t comp = __ldg(&comp_of[c]);
    if (g + k1 <= end && comp != SHN_NONE32) {  // comp == NONE: contig is not partitioned (single)
      shn_key_t key = 0;
      bool okw = true;
      if (is_codes) {
        for (int j = 0; j < k1; ++j) key = (key << 2) | (shn_key_t)((uint8_t)__ldg(&bases[g + j]) & 3u);
      } else {
        okw = pack_window(bases, g, k1, &key);
      }
      if (!okw) {
        n_bad = 1;
      } else if (key == SHN_EMPTY) {
        n_empty = 1;  // the all-T K1-mer of 32 (64) bases is the table's free-slot marker
      } else {
        uint64_t b = m.bucket_of(key);
        CompSlot* slot = nullptr;
        for (uint64_t probes = 0; !slot && probes <= m.n_buckets; ++probes) {
          CompSlot* s = m.slots + 2 * b;
          const ulonglong2 s0 = __ldcg(reinterpret_cast<const ulonglong2*>(&s[0]));
          const ulonglong2 s1 = __ldcg(reinterpret_cast<const ulonglong2*>(&s[1]));
#ifdef SHN_WIDE
          shn_key_t k[2] = {((u128)s0.y << 64) | s0.x, ((u128)s1.y << 64) | s1.x};
#else
          shn_key_t k[2] = {s0.x, s1.x};
#endif
#pragma unroll
          for (int j = 0; j < 2 && !slot; ++j) {
            if (k[j] == key) {
              slot = &s[j];
            } else if (k[j] == SHN_EMPTY) {
              shn_key_t old = shn_cas_key(&s[j].key, SHN_EMPTY, key);
              if (old == SHN_EMPTY) {
                n_new = 1;
                slot = &s[j];
              } else if (old == key) {
                slot = &s[j];
              }
            }
          }
          b = (b + 1 == m.n_buckets) ? 0 : b + 1;
        }
        if (!slot) {
          n_bad = 1;  // map full (cannot happen when the capacity check of the host holds)
        } else {
          uint32_t old = atomicCAS(&slot->comp0, SHN_NONE32, comp);
          if (old != SHN_NONE32 && old != comp) {
            old = atomicCAS(&slot->comp1, SHN_NONE32, comp);
            if (old != SHN_NONE32 && old != comp) n_over = 1;
          }
        }
      }
    }
  }
  int t_new = __syncthreads_count(n_new), t_bad = __syncthreads_count(n_bad),
      t_over = __syncthreads_count(n_over), t_empty = __syncthreads_count(n_empty);
  if (threadIdx.x == 0) {
    if (t_new) atomicAdd(&counters[0], (unsigned long long)t_new);
    if (t_bad) atomicAdd(&counters[1], (unsigned long long)t_bad);
    if (t_over) atomicAdd(&counters[2], (unsigned long long)t_over);
    if (t_empty) atomicAdd(&counters[3], (unsigned long long)t_empty);
  }
}

__global__ void __launch_bounds__(kBlock)
    map_set_weights_kernel(CompMapView m, uint32_t* __restrict__ map_w,
                           const uint64_t* __restrict__ keys, const uint32_t* __restrict__ weights,
                           uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t comps;
  uint64_t slot = map_find(m, shn_load_key(keys, i), &comps);
  if (slot != ~0ull) map_w[slot] = weights[i];
}

__global__ void __launch_bounds__(kBlock)
    window_count_kernel(const uint64_t* __restrict__ offs, uint64_t n_contigs, int k1,
                        uint64_t* __restrict__ nwin) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_contigs) return;
  uint64_t len = offs[c + 1] - offs[c];
  nwin[c] = len >= (uint64_t)k1 ? len - k1 + 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
    map_window_weights_kernel(CompMapView m, const uint32_t* __restrict__ map_w,
                              const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                              const uint64_t* __restrict__ win_off, uint64_t n_contigs,
                              uint64_t total_bases, int k1, uint32_t* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_bases) return;
  uint64_t c = find_segment(offs, n_contigs, g);
  uint64_t start = __ldg(&offs[c]), end = __ldg(&offs[c + 1]);
  if (g + k1 > end) return;
  shn_key_t key;
  uint32_t w = 0;
  if (pack_window(bases, g, k1, &key)) {
    uint64_t comps;
    uint64_t slot = map_find(m, key, &comps);
    if (slot != ~0ull) w = map_w[slot];
  }
  out[win_off[c] + (g - start)] = w;
}

struct ReadsView {
  const uint64_t* words;
  const uint64_t* woff;
  const uint32_t* len;
};

__device__ __forceinline__ shn_key_t extract_kmer(const uint64_t* __restrict__ words, uint64_t wbase,
                                                  uint32_t st, int k1) {
  uint64_t wi = wbase + (st >> 5);
  int sh = (int)(st & 31);
#ifdef SHN_WIDE   // 33..64 bases: up to three 32-base words
  u128 x = (((u128)__ldg(&words[wi]) << 64) | __ldg(&words[wi + 1])) << (2 * sh);
  if (sh + k1 > 64) x |= (u128)(__ldg(&words[wi + 2]) >> (64 - 2 * sh));
  return x >> (128 - 2 * k1);
#else
  uint64_t x = __ldg(&words[wi]) << (2 * sh);
  if (sh + k1 > 32) x |= __ldg(&words[wi + 1]) >> (64 - 2 * sh);
  return x >> (64 - 2 * k1);
#endif
}

// number of sampled K1-mers of a read of length len (get_rmers, kmers_for_component.py:186-192)
__device__ __forceinline__ uint32_t n_samples(uint32_t len, uint32_t k1) {
  if (len < k1) return 0;  // one too-short key that can never be in the map
  return (len - k1 + k1 - 1) / k1 + 1;
}

// counters: [0]=entries appended [1]=lookups [2]=valid records [3]=overflow
__global__ void __launch_bounds__(kBlock)
    assign_kernel(CompMapView m, ReadsView r0, ReadsView r1, int paired, uint64_t n_records, int k1,
                  uint64_t* __restrict__ out, uint64_t out_cap, unsigned long long* counters) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t rec = t >> 3;
  int lane8 = (int)(t & 7);
  int lane = threadIdx.x & 31;
  int grp_base = lane & ~7;
  bool active = rec < n_records;
  uint32_t len0 = 0, len1 = 0;
  uint64_t wb0 = 0, wb1 = 0;
  bool valid = false;
  if (active) {
    len0 = __ldg(&r0.len[rec]);
    wb0 = __ldg(&r0.woff[rec]);
    if (paired) {
      len1 = __ldg(&r1.len[rec]);
      wb1 = __ldg(&r1.woff[rec]);
    }
    valid = !((len0 | len1) & kLenBad);
  }
  uint32_t ns0 = valid ? n_samples(len0, k1) : 0;
  uint32_t ns1 = (valid && paired) ? n_samples(len1, k1) : 0;
  uint32_t ns = ns0 + ns1;
  // all 32 lanes iterate the same number of rounds so that the shuffles stay converged
  uint32_t max_ns = ns;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_ns = max(max_ns, __shfl_xor_sync(0xFFFFFFFFu, max_ns, o));
  unsigned long long my_lookups = 0;
  for (uint32_t base_s = 0; base_s < max_ns; base_s += 8) {
    uint32_t s = base_s + lane8;
    uint32_t v0 = SHN_NONE32, v1 = SHN_NONE32;
    if (s < ns) {
      bool second = s >= ns0;
      uint32_t si = second ? s - ns0 : s;
      uint32_t len = second ? len1 : len0;
      uint32_t nsm = second ? ns1 : ns0;
      // offsets 0, k1, 2*k1, ... and the last window (read[-k1:])
      uint32_t st = (si + 1 == nsm) ? len - k1 : si * k1;
      shn_key_t key = extract_kmer(second ? r1.words : r0.words, second ? wb1 : wb0, st, k1);
      uint64_t comps;
      uint64_t slot = map_find(m, key, &comps);
      my_lookups++;
      if (slot != ~0ull) {
        v0 = (uint32_t)comps;
        v1 = (uint32_t)(comps >> 32);
      }
    }
    // dedup inside the 8-lane group: keep the first occurrence of every component id
    if (v1 == v0) v1 = SHN_NONE32;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t o0 = __shfl_sync(0xFFFFFFFFu, v0, grp_base + j);
      uint32_t o1 = __shfl_sync(0xFFFFFFFFu, v1, grp_base + j);
      if (j < lane8) {
        if (v0 == o0 || v0 == o1) v0 = SHN_NONE32;
        if (v1 == o0 || v1 == o1) v1 = SHN_NONE32;
      }
    }
    // warp-aggregated append
    unsigned b0 = __ballot_sync(0xFFFFFFFFu, v0 != SHN_NONE32);
    unsigned b1 = __ballot_sync(0xFFFFFFFFu, v1 != SHN_NONE32);
    int total = __popc(b0) + __popc(b1);
    if (total) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&counters[0], (unsigned long long)total);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      unsigned lower = (1u << lane) - 1u;
      uint64_t p0 = base + __popc(b0 & lower);
      uint64_t p1 = base + __popc(b0) + __popc(b1 & lower);
      if (base + total <= out_cap) {
        if (v0 != SHN_NONE32) out[p0] = ((uint64_t)v0 << 32) | rec;
        if (v1 != SHN_NONE32) out[p1] = ((uint64_t)v1 << 32) | rec;
      } else if (lane == 0) {
        atomicAdd(&counters[3], 1ull);
      }
    }
  }
  // per-block totals
  __shared__ unsigned long long sh_lookups;
  if (threadIdx.x == 0) sh_lookups = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_lookups += __shfl_xor_sync(0xFFFFFFFFu, my_lookups, o);
  if (lane == 0 && my_lookups) atomicAdd(&sh_lookups, my_lookups);
  int nvalid = __syncthreads_count(valid && lane8 == 0);
  if (threadIdx.x == 0) {
    if (sh_lookups) atomicAdd(&counters[1], sh_lookups);
    if (nvalid) atomicAdd(&counters[2], (unsigned long long)nvalid);
  }
}

__global__ void __launch_bounds__(kBlock)
    comp_offsets_kernel(const uint64_t* __restrict__ entries, uint64_t n, uint32_t n_comps,
                        uint64_t* __restrict__ offs) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  // offs[c] = first entry index with comp >= c
  uint32_t prev = i == 0 ? 0u : (uint32_t)(entries[i - 1] >> 32) + 1u;
  uint32_t cur = i == n ? n_comps + 1u : (uint32_t)(entries[i] >> 32) + 1u;
  if (cur > n_comps + 1u) cur = n_comps + 1u;
  for (uint32_t c = prev; c < cur; ++c) offs[c] = i;
}

__global__ void __launch_bounds__(kBlock)
    entries_low32_kernel(const uint64_t* __restrict__ entries, uint64_t n, uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)entries[i];
}

__global__ void __launch_bounds__(kBlock)
    entries_low32_add_kernel(const uint64_t* __restrict__ entries, uint64_t n, uint32_t add,
                             uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)entries[i] + add;
}

void l4_state_free(shn_ctx* c) {
  delete static_cast<L4State*>(c->l4);
  c->l4 = nullptr;
}

L4State* l4_of(shn_ctx* c) {
  if (c->l4 && c->l4_free != &l4_state_free) shn_l4_free(c);  // state of the other key width
  if (!c->l4) {
    c->l4 = new L4State();
    c->l4_free = &l4_state_free;
  }
  return static_cast<L4State*>(c->l4);
}

CompMapView map_view(L4State* s) { return CompMapView{s->map.as<CompSlot>(), s->n_buckets}; }

unsigned long long* zero_counters(shn_ctx* c) {
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), c->stream));
  return ctr;
}

void read_counters(shn_ctx* c, unsigned long long* h, int n) {
  CUDA_CHECK(cudaMemcpyAsync(h, c->counters.p, n * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace

void l4_map_add_contigs(shn_ctx* c, const char* bases, const uint64_t* offsets,
                                 const uint32_t* comp_of_contig, uint64_t n_contigs, int k1,
                                 int reset, uint64_t expected_total, int on_device, int is_codes) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1, "k1 out of range for this key width");
  L4State* s = l4_of(c);
  if (reset) {
    // two slots per bucket; load 0.8 by default: longer probe sequences but a smaller footprint next
    // to the 126 MB L2 (assign 7.2 -> 6.3 ms at 10 M pairs against load 0.5).  SHN_MAP_LOAD in 0.3 .. 0.9
    const char* envl = getenv("SHN_MAP_LOAD");
    double load = envl ? atof(envl) : 0.8;
    if (!(load >= 0.3 && load <= 0.9)) load = 0.8;
    uint64_t nb = expected_total < 1024 ? 1024 : (uint64_t)((double)expected_total * 0.5 / load) + 1;
    s->map.reserve(nb * 2 * sizeof(CompSlot));
    s->map_w.reserve(nb * 2 * sizeof(uint32_t));
    s->n_buckets = nb;
    s->k1 = k1;
    s->n_keys = 0;
    ProfScope ps(c, "l4_map_clear");
    unsigned grid =
        (unsigned)std::min<uint64_t>((nb * 2 + kBlock - 1) / kBlock, (uint64_t)c->sm_count * 32);
    map_clear_kernel<<<grid, kBlock, 0, c->stream>>>(s->map.as<CompSlot>(), s->map_w.as<uint32_t>(),
                                                     nb * 2);
    KERNEL_CHECK();
  }
  SHN_CHECK(s->n_buckets > 0, "component map not initialised (call with reset != 0 first)");
  SHN_CHECK(k1 == s->k1, "k1 differs from the component map's");
  if (n_contigs == 0) return;
  uint64_t total = 0;
  const uint64_t* d_offs;
  if (on_device) {
    d_offs = offsets;
    CUDA_CHECK(cudaMemcpyAsync(&total, offsets + n_contigs, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  } else {
    total = offsets[n_contigs];
    d_offs = (const uint64_t*)InputView::get(c, offsets, (n_contigs + 1) * 8, 0, s->stage_b);
  }
  const char* d_bases = (const char*)InputView::get(c, bases, total, on_device, s->stage_a);
  // the component ids always come from the host (they are decided by the host-side packing)
  const uint32_t* d_comp =
      (const uint32_t*)InputView::get(c, comp_of_contig, n_contigs * 4, 0, s->stage_c);
  // K1-mer windows of this batch: exact from host offsets; device-resident contigs come with the
  // exact count in expected_total (the allowed set of shn_l3_run)
  uint64_t windows = expected_total;
  if (!on_device) {
    windows = 0;
    for (uint64_t i = 0; i < n_contigs; ++i) {
      const uint64_t len = offsets[i + 1] - offsets[i];
      if (len >= (uint64_t)k1) windows += len - k1 + 1;
    }
  }
  SHN_CHECK(s->n_keys + windows <= s->n_buckets * 2,
            "component map too small: expected_total_k1mers was underestimated");
  unsigned long long* ctr = zero_counters(c);
  if (total) {
    ProfScope ps(c, "l4_map_add");
    map_add_kernel<<<shn_grid(total, kBlock), kBlock, 0, c->stream>>>(
        map_view(s), d_bases, d_offs, d_comp, n_contigs, total, k1, is_codes, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h[4];
  read_counters(c, h, 4);
  SHN_CHECK(h[1] == 0, "contig contains a character outside ACGT (or the component map is full)");
  SHN_CHECK(h[2] == 0, "a K1-mer belongs to more than two components (unsupported)");
  SHN_CHECK(h[3] == 0, "a contig holds the poly-T K1-mer of 32 (64) bases, which the component map cannot store "
                       "(low-complexity: never produced by extension_correction)");
  s->n_keys += h[0];
}

void l4_map_set_weights(shn_ctx* c, const uint64_t* keys, const uint32_t* weights,
                                 uint64_t n, int on_device) {
  L4State* s = l4_of(c);
  SHN_CHECK(s->n_buckets > 0, "component map not initialised");
  if (n == 0) return;
  const uint64_t* d_keys =
      (const uint64_t*)InputView::get(c, keys, n * 8 * SHN_KEY_WORDS, on_device, s->stage_a);
  const uint32_t* d_w = (const uint32_t*)InputView::get(c, weights, n * 4, on_device, s->stage_b);
  ProfScope ps(c, "l4_map_set_weights");
  map_set_weights_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(
      map_view(s), s->map_w.as<uint32_t>(), d_keys, d_w, n);
  KERNEL_CHECK();
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void l4_map_window_weights(shn_ctx* c, const char* bases, const uint64_t* offsets,
                                    uint64_t n_contigs, int k1, uint32_t* h_weights) {
  L4State* s = l4_of(c);
  SHN_CHECK(s->n_buckets > 0, "component map not initialised");
  if (n_contigs == 0) return;
  uint64_t total = offsets[n_contigs];
  const uint64_t* d_offs = (const uint64_t*)InputView::get(c, offsets, (n_contigs + 1) * 8, 0, s->stage_b);
  const char* d_bases = (const char*)InputView::get(c, bases, total, 0, s->stage_a);
  DevBuf nwin, winoff, out;
  nwin.reserve((n_contigs + 1) * 8);
  winoff.reserve((n_contigs + 1) * 8);
  CUDA_CHECK(cudaMemsetAsync(nwin.p, 0, (n_contigs + 1) * 8, c->stream));
  window_count_kernel<<<shn_grid(n_contigs, kBlock), kBlock, 0, c->stream>>>(d_offs, n_contigs, k1,
                                                                             nwin.as<uint64_t>());
  KERNEL_CHECK();
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, nwin.as<uint64_t>(), winoff.as<uint64_t>(),
                                           (int)(n_contigs + 1), c->stream));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, nwin.as<uint64_t>(), winoff.as<uint64_t>(),
                                           (int)(n_contigs + 1), c->stream));
  uint64_t total_win = 0;
  CUDA_CHECK(cudaMemcpyAsync(&total_win, winoff.as<uint64_t>() + n_contigs, 8, cudaMemcpyDeviceToHost,
                             c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (total_win == 0) return;
  out.reserve(total_win * 4);
  {
    ProfScope ps(c, "l4_window_weights");
    map_window_weights_kernel<<<shn_grid(total, kBlock), kBlock, 0, c->stream>>>(
        map_view(s), s->map_w.as<uint32_t>(), d_bases, d_offs, winoff.as<uint64_t>(), n_contigs, total,
        k1, out.as<uint32_t>());
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(h_weights, out.p, total_win * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void l4_assign(shn_ctx* c, int paired, int k1, uint64_t* n_assign, uint64_t* n_lookups,
                        uint64_t* n_valid) {
  L4State* s = l4_of(c);
  SHN_CHECK(s->n_buckets > 0, "component map not initialised");
  SHN_CHECK(k1 == s->k1, "k1 differs from the component map's");
  ReadsState* rs = shn_reads_of(c);
  uint64_t n = rs->reads[0].n;
  if (paired) SHN_CHECK(rs->reads[1].n == n, "mate files hold different numbers of records");
  s->n_assign = 0;
  *n_assign = *n_lookups = *n_valid = 0;
  if (n == 0) return;
  ReadsView r0{rs->reads[0].words.as<uint64_t>(), rs->reads[0].woff.as<uint64_t>(),
               rs->reads[0].len.as<uint32_t>()};
  ReadsView r1 = r0;
  if (paired)
    r1 = ReadsView{rs->reads[1].words.as<uint64_t>(), rs->reads[1].woff.as<uint64_t>(),
                   rs->reads[1].len.as<uint32_t>()};
  uint64_t cap = std::max<uint64_t>(2 * n, 1024);
  unsigned long long h[4];
  for (int attempt = 0;; ++attempt) {
    s->assign.reserve(cap * 8);
    unsigned long long* ctr = zero_counters(c);
    {
      ProfScope ps(c, "l4_assign");
      assign_kernel<<<shn_grid(n * 8, kBlock), kBlock, 0, c->stream>>>(
          map_view(s), r0, r1, paired, n, k1, s->assign.as<uint64_t>(), cap, ctr);
      KERNEL_CHECK();
    }
    read_counters(c, h, 4);
    if (h[3] == 0) break;
    SHN_CHECK(attempt < 2, "assignment buffer overflow after resize");
    cap = h[0] + 1024;  // exact size known now
  }
  uint64_t m = h[0];
  *n_lookups = h[1];
  *n_valid = h[2];
  if (m > 0) {
    // sort by (component, record) and drop duplicates (records longer than 8 samples)
    DevBuf sorted, nuniq;
    sorted.reserve(m * 8);
    nuniq.reserve(8);
    ProfScope ps(c, "l4_sort_unique", 3);
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb, s->assign.as<uint64_t>(),
                                              sorted.as<uint64_t>(), (int64_t)m, 0, 64, c->stream));
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(c->tmp(tb), tb, s->assign.as<uint64_t>(),
                                              sorted.as<uint64_t>(), (int64_t)m, 0, 64, c->stream));
    tb = 0;
    CUDA_CHECK(cub::DeviceSelect::Unique(nullptr, tb, sorted.as<uint64_t>(), s->assign.as<uint64_t>(),
                                         nuniq.as<uint64_t>(), (int64_t)m, c->stream));
    CUDA_CHECK(cub::DeviceSelect::Unique(c->tmp(tb), tb, sorted.as<uint64_t>(),
                                         s->assign.as<uint64_t>(), nuniq.as<uint64_t>(), (int64_t)m,
                                         c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&m, nuniq.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  s->n_assign = m;
  *n_assign = m;
}

void l4_get_assignments(shn_ctx* c, uint32_t n_comps, uint64_t* h_offs, uint32_t* h_idx) {
  L4State* s = l4_of(c);
  uint64_t m = s->n_assign;
  DevBuf offs, idx;
  offs.reserve((uint64_t)(n_comps + 2) * 8);
  idx.reserve((m + 1) * 4);
  comp_offsets_kernel<<<shn_grid(m + 1, kBlock), kBlock, 0, c->stream>>>(s->assign.as<uint64_t>(), m,
                                                                         n_comps, offs.as<uint64_t>());
  KERNEL_CHECK();
  if (m) {
    entries_low32_kernel<<<shn_grid(m, kBlock), kBlock, 0, c->stream>>>(s->assign.as<uint64_t>(), m,
                                                                        idx.as<uint32_t>());
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(h_idx, idx.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK(cudaMemcpyAsync(h_offs, offs.p, (uint64_t)(n_comps + 1) * 8, cudaMemcpyDeviceToHost,
                             c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// the same into caller-owned DEVICE buffers (sharded path: the lists of the ranks are merged on
// the device); record indices are shifted by `first_record` (this rank's first global record)
void l4_assignments_dev(shn_ctx* c, uint32_t n_comps, uint64_t first_record, uint64_t* d_offs,
                        uint32_t* d_idx) {
  L4State* s = l4_of(c);
  uint64_t m = s->n_assign;
  comp_offsets_kernel<<<shn_grid(m + 1, kBlock), kBlock, 0, c->stream>>>(s->assign.as<uint64_t>(), m,
                                                                         n_comps, d_offs);
  KERNEL_CHECK();
  if (m) {
    entries_low32_add_kernel<<<shn_grid(m, kBlock), kBlock, 0, c->stream>>>(s->assign.as<uint64_t>(), m,
                                                                            (uint32_t)first_record, d_idx);
    KERNEL_CHECK();
  }
}

}  // namespace SHN_NS
