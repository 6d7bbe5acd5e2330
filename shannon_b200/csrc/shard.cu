// e: the front end on hash-sharded K1-mer tables, one table shard per rank.
//
// The walks of extension_correction.py:343-397 never leave a connected component of the K1-mer
// successor graph, so the unit of independence is that component.  The sharded path therefore
//   1. routes every input line to owner(K1-mer) = hash of the K1-mer's MINIMIZER (the 11-mer with
//      the smallest hash): consecutive K1-mers of a chain share their minimizer ~7 times out of 8,
//      so most successor edges stay rank-local                                 (route_lines_*)
//   2. builds the local table shard from the routed records                      (table_build_records)
//   3. labels the LOCAL components with the lock-free union-find                  (cc_local)
//   4. asks the owners of the non-local successor candidates whether they exist   (cc_cross /
//      cc_resolve: one all-to-all of (successor key, my local component)), which yields the edges
//      of the "super-node" graph whose nodes are the local components of all ranks
//   5. all-gathers those edges and labels the super-node graph, replicated         (cc_merge)
//   6. assigns whole K1-mer graph components to ranks (balanced by size, cc_sizes) and routes
//      every table entry to the rank that owns its component                      (cc_route)
// after which every rank holds complete components and runs the unchanged single-GPU walks.
//
// Records on the wire have the size of a table slot: {key, payload}.  Lines / table entries carry
// payload = global input line << 30 | weight (the seed tie-break needs the order of the un-sharded
// k1mer.dict_org, extension_correction.py:334); successor queries carry the asking component.
// The exchanges themselves are torch.distributed all_to_all_single calls on the same CUDA stream
// (shannon_b200/dist.py); nothing here synchronises except to hand split sizes to the host.
#include <cub/cub.cuh>

#include "common.cuh"
#include "impls.h"
#include "table_dev.cuh"
#include "uf_dev.cuh"

namespace SHN_NS {

struct ShnRec {
  shn_key_t key;
  uint64_t payload;
#ifdef SHN_WIDE
  uint64_t pad;
#endif
};
static_assert(sizeof(ShnRec) == sizeof(ShnSlot), "a routing record has the size of a table slot");

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kMaxRanks = 64;
constexpr int kMinimizerM = 11;
constexpr int kGlineShift = 30;                       // payload = gline << 30 | weight
constexpr uint64_t kGlineMax = 1ull << (64 - kGlineShift);

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // a bijection on 32-bit words
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}

// smallest mix32 over the m-mers at base offsets [p_lo, p_hi] counted from the END of the K1-mer
// (offset p = bits 2p .. 2p+2m-1 of the key)
__device__ __forceinline__ uint32_t min_mmer_hash(shn_key_t key, int p_lo, int p_hi) {
  const uint32_t mmask = (1u << (2 * kMinimizerM)) - 1u;
  uint32_t best = 0xFFFFFFFFu;
  for (int p = p_lo; p <= p_hi; ++p) best = min(best, mix32((uint32_t)(key >> (2 * p)) & mmask));
  return best;
}
__device__ __forceinline__ uint32_t owner_from_hash(uint32_t h, uint32_t nranks) {
  return (uint32_t)(((uint64_t)mix32(h ^ 0x5BD1E995u) * nranks) >> 32);  // min-of-many is not uniform: re-mix
}
__device__ __forceinline__ uint32_t owner_of(shn_key_t key, int k1, uint32_t nranks) {
  if (k1 <= kMinimizerM) return owner_from_hash(mix32((uint32_t)shn_key_hash(key)), nranks);
  return owner_from_hash(min_mmer_hash(key, 0, k1 - kMinimizerM), nranks);
}

// ---- routing: count per destination, then scatter into a send buffer contiguous per rank -----
// ctr[0..63] = records per destination (count pass) / write cursors (fill pass); ctr[64] = errors
template <typename Src, bool kFill>
__global__ void __launch_bounds__(kBlock)
    route_kernel(Src src, uint64_t n_items, uint32_t nranks, unsigned long long* ctr,
                 const unsigned long long* __restrict__ seg_start, ShnRec* __restrict__ send) {
  __shared__ uint32_t hist[kMaxRanks];
  __shared__ unsigned long long base[kMaxRanks];
  if (threadIdx.x < kMaxRanks) hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t dest[Src::kMaxEmit], loc[Src::kMaxEmit];
  ShnRec rec[Src::kMaxEmit];
  int ne = 0;
  if (i < n_items) ne = src.emit(i, dest, rec, ctr + kMaxRanks);
#pragma unroll
  for (int e = 0; e < Src::kMaxEmit; ++e)
    if (e < ne) loc[e] = atomicAdd(&hist[dest[e]], 1u);
  __syncthreads();
  if (threadIdx.x < nranks) {
    const uint32_t hcnt = hist[threadIdx.x];
    base[threadIdx.x] = hcnt ? atomicAdd(&ctr[threadIdx.x], (unsigned long long)hcnt) : 0ull;
  }
  if (!kFill) return;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < Src::kMaxEmit; ++e)
    if (e < ne) send[seg_start[dest[e]] + base[dest[e]] + loc[e]] = rec[e];
}

__device__ __forceinline__ ShnRec make_rec(shn_key_t key, uint64_t payload) {
  ShnRec r;
  r.key = key;
  r.payload = payload;
#ifdef SHN_WIDE
  r.pad = 0;
#endif
  return r;
}

// input lines (load_kmers, extension_correction.py:209-219): one record per line, two with -d
struct LinesSrc {
  static constexpr int kMaxEmit = 2;
  const uint64_t* keys;
  const uint32_t* counts;
  uint64_t first_line;
  int ds, k1;
  uint32_t nranks;
  __device__ __forceinline__ int emit(uint64_t i, uint32_t* dest, ShnRec* rec, unsigned long long* err) const {
    const shn_key_t key = shn_load_key(keys, i);
    const uint32_t w = counts[i];
    const uint64_t gline = ds ? 2 * (first_line + i) : first_line + i;
    if ((key & ~shn_key_mask(k1)) || w >= SHN_WEIGHT_MASK || gline + 1 >= kGlineMax) {
      atomicAdd(err, 1ull);
      return 0;
    }
    dest[0] = owner_of(key, k1, nranks);
    rec[0] = make_rec(key, (gline << kGlineShift) | w);
    if (!ds) return 1;
    const shn_key_t rc = shn_revcomp(key, k1);
    dest[1] = owner_of(rc, k1, nranks);
    rec[1] = make_rec(rc, ((gline + 1) << kGlineShift) | w);
    return 2;
  }
};

// component id of a table slot in the global numbering of local components ("super-nodes")
struct CcView {
  const ShnSlot* slots;
  const uint32_t* parent;   // root slot per slot after flatten (SHN_NONE32: free slot)
  const uint32_t* root_id;  // dense local id of a root slot
  uint32_t gid_base;
  __device__ __forceinline__ uint32_t gid(uint64_t slot) const { return gid_base + root_id[parent[slot]]; }
};

// successor candidates owned by another rank: "does (x[1:] . b) exist?  I am component gid"
struct CrossSrc {
  static constexpr int kMaxEmit = 4;
  CcView cc;
  int k1;
  uint32_t nranks, rank;
  __device__ __forceinline__ int emit(uint64_t i, uint32_t* dest, ShnRec* rec, unsigned long long*) const {
    const shn_key_t key = cc.slots[i].key;
    if (key == SHN_EMPTY) return 0;
    const shn_key_t pre = (key << 2) & shn_key_mask(k1);
    int ne = 0;
    if (k1 <= kMinimizerM) {
      for (uint32_t b = 0; b < 4; ++b) {
        const uint32_t o = owner_of(pre | (shn_key_t)b, k1, nranks);
        if (o != rank) {
          dest[ne] = o;
          rec[ne++] = make_rec(pre | (shn_key_t)b, cc.gid(i));
        }
      }
      return ne;
    }
    // the successor shares all m-mers but its last with x: minimum over the shared ones once
    const uint32_t shared = k1 - kMinimizerM >= 1 ? min_mmer_hash(pre, 1, k1 - kMinimizerM) : 0xFFFFFFFFu;
    const uint32_t mmask = (1u << (2 * kMinimizerM)) - 1u;
    uint32_t gid = 0;
    bool have_gid = false;
    for (uint32_t b = 0; b < 4; ++b) {
      const shn_key_t succ = pre | (shn_key_t)b;
      const uint32_t o = owner_from_hash(min(shared, mix32((uint32_t)succ & mmask)), nranks);
      if (o != rank) {
        if (!have_gid) {
          gid = cc.gid(i);
          have_gid = true;
        }
        dest[ne] = o;
        rec[ne++] = make_rec(succ, gid);
      }
    }
    return ne;
  }
};

// every table entry to the rank that owns its K1-mer graph component
struct ByCompSrc {
  static constexpr int kMaxEmit = 1;
  CcView cc;
  const uint32_t* final_of_super;
  const uint32_t* owner_of_final;
  const uint64_t* gline;  // global input line of local first-occurrence index
  __device__ __forceinline__ int emit(uint64_t i, uint32_t* dest, ShnRec* rec, unsigned long long*) const {
    shn_key_t key;
    uint32_t wz, idx;
    table_load_slot(cc.slots, i, &key, &wz, &idx);
    if (key == SHN_EMPTY) return 0;
    dest[0] = owner_of_final[final_of_super[cc.gid(i)]];
    rec[0] = make_rec(key, (gline[idx] << kGlineShift) | (wz & SHN_WEIGHT_MASK));
    return 1;
  }
};

struct ShardState {
  DevBuf parent, root_id;   // uint32 [n_slots], [n_slots + 1]
  uint64_t n_slots = 0, n_local = 0;
  DevBuf final_of_super;    // uint32 [n_super]
  uint64_t n_super = 0, n_final = 0;
  DevBuf ctr;               // 65 counters + 65 segment starts
};

void shard_state_free(shn_ctx* c) {
  delete static_cast<ShardState*>(c->shard);
  c->shard = nullptr;
}

ShardState* shard_of(shn_ctx* c) {
  if (c->shard && c->shard_free != &shard_state_free) {
    c->shard_free(c);
    c->shard = nullptr;
  }
  if (!c->shard) {
    c->shard = new ShardState();
    c->shard_free = &shard_state_free;
  }
  return static_cast<ShardState*>(c->shard);
}

// count pass (send == nullptr: h_counts is written) or fill pass (h_counts is read)
template <typename Src>
void route(shn_ctx* c, const char* name, const Src& src, uint64_t n_items, uint32_t nranks,
           uint64_t* h_counts, void* send) {
  SHN_CHECK(nranks >= 1 && nranks <= kMaxRanks, "nranks out of range (1..64)");
  ShardState* s = shard_of(c);
  s->ctr.reserve((2 * kMaxRanks + 2) * 8);
  unsigned long long* ctr = s->ctr.as<unsigned long long>();
  unsigned long long* seg = ctr + kMaxRanks + 1;
  cudaStream_t st = c->stream;
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, (kMaxRanks + 1) * 8, st));
  unsigned long long h[kMaxRanks + 1];
  if (send) {
    unsigned long long hs[kMaxRanks + 1];
    hs[0] = 0;
    for (uint32_t r = 0; r < nranks; ++r) hs[r + 1] = hs[r] + h_counts[r];
    CUDA_CHECK(cudaMemcpyAsync(seg, hs, (nranks + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));  // hs lives on this stack frame
  }
  if (n_items) {
    ProfScope ps(c, name);
    if (send)
      route_kernel<Src, true><<<shn_grid(n_items, kBlock), kBlock, 0, st>>>(src, n_items, nranks, ctr, seg,
                                                                           static_cast<ShnRec*>(send));
    else
      route_kernel<Src, false><<<shn_grid(n_items, kBlock), kBlock, 0, st>>>(src, n_items, nranks, ctr, seg,
                                                                            nullptr);
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, (kMaxRanks + 1) * 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  SHN_CHECK(h[kMaxRanks] == 0,
            "routing: key wider than 2*k1 bits, a count above 2^30-2, or more than 2^34 input lines");
  if (send) {
    for (uint32_t r = 0; r < nranks; ++r)
      SHN_CHECK(h[r] == h_counts[r], "routing: the fill pass disagrees with the count pass");
  } else {
    for (uint32_t r = 0; r < nranks; ++r) h_counts[r] = h[r];
  }
}

// ---- table shard from routed records ------------------------------------------------------------
// The first-occurrence word of a slot holds the POSITION of the record in the receive buffer;
// gline[position] is its global input line, and the seed order of the walks reads the line through
// that table (l3.cu seed_emit_kernel), so nothing has to be sorted here.  A key that arrives on
// several lines (repeated lines, palindromes under -d) keeps the record with the smallest global
// line: the later arrivals are listed and resolved by dup_resolve_kernel.
// counters as table_insert_kernel: [0]=new keys [1]=low-complexity [2]=later arrivals [3]=bad
__global__ void __launch_bounds__(kBlock)
    insert_records_kernel(ShnTableView t, const ShnRec* __restrict__ recs, uint64_t n, int k1,
                          uint64_t* __restrict__ gline, uint32_t* __restrict__ dups,
                          unsigned long long* counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_new = 0, n_low = 0, n_bad = 0;
  if (i < n) {
    const ShnRec r = recs[i];
    const uint32_t w = (uint32_t)(r.payload & SHN_WEIGHT_MASK);
    gline[i] = r.payload >> kGlineShift;
    if (r.key & ~shn_key_mask(k1)) {
      n_bad = 1;
    } else if (shn_low_complexity(r.key, k1)) {
      n_low = 1;
    } else {
      uint32_t old = 0;
      const int before = n_new;
      const uint64_t slot = table_insert_add(t, r.key, w, (uint32_t)i, &n_new, &old, false);
      if (slot == ~0ull || (uint64_t)old + w >= (uint64_t)SHN_WEIGHT_MASK) n_bad = 1;
      else if (n_new == before) dups[atomicAdd(&counters[2], 1ull)] = (uint32_t)i;
    }
  }
  int tot_new = __syncthreads_count(n_new), tot_low = __syncthreads_count(n_low),
      tot_bad = __syncthreads_count(n_bad);
  if (threadIdx.x == 0) {
    if (tot_new) atomicAdd(&counters[0], (unsigned long long)tot_new);
    if (tot_low) atomicAdd(&counters[1], (unsigned long long)tot_low);
    if (tot_bad) atomicAdd(&counters[3], (unsigned long long)tot_bad);
  }
}

// runs after insert_records_kernel has finished: every slot's idx word is initialised
__global__ void __launch_bounds__(kBlock)
    dup_resolve_kernel(ShnTableView t, const ShnRec* __restrict__ recs, const uint32_t* __restrict__ dups,
                       uint64_t n_dups, const uint64_t* __restrict__ gline) {
  uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_dups) return;
  const uint32_t i = dups[d];
  uint32_t w;
  const uint64_t slot = table_find(t, recs[i].key, &w);
  if (slot == ~0ull) return;
  uint32_t* p = &t.slots[slot].idx;
  uint32_t cur = *reinterpret_cast<volatile uint32_t*>(p);
  while (gline[i] < gline[cur]) {
    const uint32_t old = atomicCAS(p, cur, i);
    if (old == cur) break;
    cur = old;
  }
}

// ---- local components ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) iota_kernel(uint32_t* p, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    uf_local_kernel(ShnTableView t, uint32_t* parent, uint64_t n_slots, int k1) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const shn_key_t key = t.slots[i].key;
  if (key == SHN_EMPTY) {
    parent[i] = SHN_NONE32;
    return;
  }
  // the four successors share one home bucket and probe sequence (placement by the K-base prefix)
  uint64_t slot[4];
  uint32_t w[4];
  const uint32_t sm = table_find_successors(t, key, k1, slot, w);
  for (uint32_t m = sm; m; m &= m - 1) {  // k-th existing successor of every lane together (cf. uf_edges_kernel)
    const int b = __ffs(m) - 1;
    const uint64_t s = b == 0 ? slot[0] : (b == 1 ? slot[1] : (b == 2 ? slot[2] : slot[3]));
    if (s != i) uf_union(parent, (uint32_t)i, (uint32_t)s);
  }
}

__global__ void __launch_bounds__(kBlock)
    flatten_kernel(uint32_t* parent, uint64_t n, uint32_t* __restrict__ is_root) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = __ldcg(&parent[i]);
  const bool occ = p != SHN_NONE32;
  uint32_t root = (uint32_t)i;
  if (occ && p != (uint32_t)i) root = uf_find_ro(parent, p);
  is_root[i] = (occ && root == (uint32_t)i) ? 1u : 0u;
  if (occ && root != p) parent[i] = root;
}

void exclusive_sum_u32(shn_ctx* c, const uint32_t* in, uint32_t* out, uint64_t n) {
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, c->stream));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, in, out, (int64_t)n, c->stream));
}

// ---- cross-rank edges ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
    resolve_kernel(ShnTableView t, CcView cc, const ShnRec* __restrict__ q, uint64_t n,
                   uint64_t* __restrict__ edges, unsigned long long* cursor) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t e = 0;
  bool have = false;
  if (i < n) {
    uint32_t w;
    const uint64_t s = table_find(t, q[i].key, &w);
    if (s != ~0ull) {
      have = true;
      e = (uint64_t)(uint32_t)q[i].payload | ((uint64_t)cc.gid(s) << 32);
    }
  }
  const unsigned b = __ballot_sync(0xFFFFFFFFu, have);
  if (b) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(b));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (have) edges[base + __popc(b & ((1u << lane) - 1u))] = e;
  }
}

__global__ void __launch_bounds__(kBlock)
    union_edges_kernel(const uint64_t* __restrict__ edges, uint64_t n, uint32_t* parent, uint32_t n_nodes,
                       unsigned long long* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t a = (uint32_t)edges[i], b = (uint32_t)(edges[i] >> 32);
  if (a >= n_nodes || b >= n_nodes) {
    atomicAdd(err, 1ull);
    return;
  }
  if (a != b) uf_union(parent, a, b);
}

__global__ void __launch_bounds__(kBlock)
    final_ids_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ root_id, uint64_t n,
                     uint32_t* __restrict__ final_of_super) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) final_of_super[i] = root_id[parent[i]];
}

// K1-mers per local component, then per final component (two levels: the second pass issues one
// atomic per LOCAL component, so a few giant final components do not serialise 10^8 atomics)
__global__ void __launch_bounds__(kBlock)
    local_sizes_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ root_id, uint64_t n_slots,
                       uint32_t* __restrict__ local_size) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // neighbouring slots mostly belong to the same local component (minimizer regions): one atomic per
  // distinct component of a warp -- one per slot serialised on the large components (measured: 80 ms
  // of a 230 ms step on four ranks)
  uint32_t r = SHN_NONE32;
  if (i < n_slots) {
    const uint32_t p = parent[i];
    if (p != SHN_NONE32) r = root_id[p];
  }
  const unsigned same = __match_any_sync(0xFFFFFFFFu, r);
  if (r != SHN_NONE32 && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&local_size[r], (uint32_t)__popc(same));
}
__global__ void __launch_bounds__(kBlock)
    final_sizes_kernel(const uint32_t* __restrict__ local_size, uint64_t n_local, uint32_t gid_base,
                       const uint32_t* __restrict__ final_of_super, unsigned long long* __restrict__ sizes) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the same warp aggregation: local components with neighbouring ids mostly share their final one
  uint32_t f = SHN_NONE32, v = 0;
  if (i < n_local) {
    f = final_of_super[gid_base + i];
    v = local_size[i];
  }
  const unsigned same = __match_any_sync(0xFFFFFFFFu, f);
  const int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
  unsigned long long tot = 0;  // sum of v over the lanes of `same`, gathered by the leader
  for (unsigned m = same; m; m &= m - 1) {
    const int src = __ffs(m) - 1;
    const uint32_t x = __shfl_sync(same, v, src);
    tot += x;
  }
  if (f != SHN_NONE32 && lane == leader) atomicAdd(&sizes[f], tot);
}

CcView cc_view(shn_ctx* c, ShardState* s, uint64_t gid_base) {
  SHN_CHECK(s->n_slots == c->n_buckets * SHN_BSLOTS && s->parent.p, "shn_cc_local has not run on the current table");
  return CcView{c->table.as<ShnSlot>(), s->parent.as<uint32_t>(), s->root_id.as<uint32_t>(), (uint32_t)gid_base};
}

}  // namespace

void route_lines(shn_ctx* c, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n,
                 uint64_t first_line, int ds, int k1, uint32_t nranks, uint64_t* h_counts, void* send) {
  SHN_CHECK(k1 >= 1 && k1 <= SHN_MAX_K1 && (SHN_KEY_WORDS == 1 || k1 > 32), "k1 out of range for this key width");
  LinesSrc src{d_keys, d_counts, first_line, ds, k1, nranks};
  route(c, send ? "route_lines_fill" : "route_lines_count", src, n, nranks, h_counts, send);
}

void table_build_records(shn_ctx* c, const void* d_recs, uint64_t n, int k1) {
  const ShnRec* recs = static_cast<const ShnRec*>(d_recs);
  table_begin(c, n, k1, 0);
  c->gline_buf.reserve(std::max<uint64_t>(n, 1) * 8);
  uint64_t* d_gline = c->gline_buf.as<uint64_t>();
  cudaStream_t st = c->stream;
  if (n) {
    DevBuf dups;
    dups.reserve(n * 4);
    unsigned long long* ctr = c->counters.as<unsigned long long>();
    {
      ProfScope ps(c, "table_insert");
      insert_records_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(table_view(c), recs, n, k1, d_gline,
                                                                   dups.as<uint32_t>(), ctr);
      KERNEL_CHECK();
    }
    unsigned long long n_dups = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_dups, ctr + 2, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_dups) {
      ProfScope ps(c, "table_dup_lines");
      dup_resolve_kernel<<<shn_grid(n_dups, kBlock), kBlock, 0, st>>>(table_view(c), recs, dups.as<uint32_t>(),
                                                                     n_dups, d_gline);
      KERNEL_CHECK();
      CUDA_CHECK(cudaStreamSynchronize(st));  // dups returns to the pool
    }
  }
  table_finish(c);
  c->gline_dev = d_gline;  // the seed order of shn_l3_walks reads global lines through it
}

void cc_local(shn_ctx* c, uint64_t* n_local) {
  SHN_CHECK(c->n_buckets > 0, "no K1-mer table built");
  ShardState* s = shard_of(c);
  const uint64_t n_slots = c->n_buckets * SHN_BSLOTS;
  SHN_CHECK(n_slots < 0xFFFFFFFFull, "table shard too large for 32-bit slot indices");
  s->n_slots = n_slots;
  s->parent.reserve(n_slots * 4);
  s->root_id.reserve((n_slots + 1) * 4);
  DevBuf flag;
  flag.reserve((n_slots + 1) * 4);
  cudaStream_t st = c->stream;
  const unsigned sg = (unsigned)std::min<uint64_t>((n_slots + kBlock - 1) / kBlock, (uint64_t)c->sm_count * 32);
  {
    ProfScope ps(c, "cc_init");
    iota_kernel<<<sg, kBlock, 0, st>>>(s->parent.as<uint32_t>(), n_slots);
    KERNEL_CHECK();
  }
  {
    ProfScope ps(c, "cc_local_edges");
    uf_local_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(table_view(c), s->parent.as<uint32_t>(), n_slots,
                                                                 c->k1);
    KERNEL_CHECK();
  }
  {
    ProfScope ps(c, "cc_flatten", 2);
    CUDA_CHECK(cudaMemsetAsync(flag.as<uint32_t>() + n_slots, 0, 4, st));
    flatten_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(s->parent.as<uint32_t>(), n_slots,
                                                                flag.as<uint32_t>());
    KERNEL_CHECK();
    exclusive_sum_u32(c, flag.as<uint32_t>(), s->root_id.as<uint32_t>(), n_slots + 1);
  }
  uint32_t n = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n, s->root_id.as<uint32_t>() + n_slots, 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  s->n_local = n;
  *n_local = n;
}

void cc_cross(shn_ctx* c, uint32_t nranks, uint32_t rank, uint64_t gid_base, uint64_t* h_counts, void* send) {
  ShardState* s = shard_of(c);
  SHN_CHECK(gid_base + s->n_local < 0xFFFFFFFFull, "more than 2^32-1 local components over all ranks");
  CrossSrc src{cc_view(c, s, gid_base), c->k1, nranks, rank};
  route(c, send ? "cc_cross_fill" : "cc_cross_count", src, s->n_slots, nranks, h_counts, send);
}

void cc_resolve(shn_ctx* c, const void* d_recs, uint64_t n, uint64_t gid_base, uint64_t* d_edges,
                uint64_t* n_edges) {
  ShardState* s = shard_of(c);
  *n_edges = 0;
  if (n == 0) return;
  CcView cc = cc_view(c, s, gid_base);
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  cudaStream_t st = c->stream;
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
  {
    ProfScope ps(c, "cc_resolve");
    resolve_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(table_view(c), cc, static_cast<const ShnRec*>(d_recs), n,
                                                          d_edges, ctr);
    KERNEL_CHECK();
  }
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, ctr, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  *n_edges = h;
}

void cc_merge(shn_ctx* c, const uint64_t* d_edges, uint64_t n_edges, uint64_t n_super, uint64_t* n_final) {
  ShardState* s = shard_of(c);
  SHN_CHECK(n_super < 0xFFFFFFFFull, "more than 2^32-1 local components over all ranks");
  s->n_super = n_super;
  s->final_of_super.reserve(std::max<uint64_t>(n_super, 1) * 4);
  *n_final = 0;
  s->n_final = 0;
  if (n_super == 0) return;
  DevBuf parent, flag, rid;
  parent.reserve(n_super * 4);
  flag.reserve((n_super + 1) * 4);
  rid.reserve((n_super + 1) * 4);
  cudaStream_t st = c->stream;
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
  {
    ProfScope ps(c, "cc_merge", 5);
    const unsigned sg = (unsigned)std::min<uint64_t>((n_super + kBlock - 1) / kBlock, (uint64_t)c->sm_count * 32);
    iota_kernel<<<sg, kBlock, 0, st>>>(parent.as<uint32_t>(), n_super);
    KERNEL_CHECK();
    if (n_edges) {
      union_edges_kernel<<<shn_grid(n_edges, kBlock), kBlock, 0, st>>>(d_edges, n_edges, parent.as<uint32_t>(),
                                                                      (uint32_t)n_super, ctr);
      KERNEL_CHECK();
    }
    CUDA_CHECK(cudaMemsetAsync(flag.as<uint32_t>() + n_super, 0, 4, st));
    flatten_kernel<<<shn_grid(n_super, kBlock), kBlock, 0, st>>>(parent.as<uint32_t>(), n_super, flag.as<uint32_t>());
    KERNEL_CHECK();
    exclusive_sum_u32(c, flag.as<uint32_t>(), rid.as<uint32_t>(), n_super + 1);
    final_ids_kernel<<<shn_grid(n_super, kBlock), kBlock, 0, st>>>(parent.as<uint32_t>(), rid.as<uint32_t>(), n_super,
                                                                  s->final_of_super.as<uint32_t>());
    KERNEL_CHECK();
  }
  uint32_t nf = 0;
  unsigned long long bad = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nf, rid.as<uint32_t>() + n_super, 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(&bad, ctr, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  SHN_CHECK(bad == 0, "cc_merge: an edge names a component id >= n_super");
  s->n_final = nf;
  *n_final = nf;
}

void cc_sizes(shn_ctx* c, uint64_t gid_base, uint64_t* d_sizes) {
  ShardState* s = shard_of(c);
  SHN_CHECK(gid_base + s->n_local <= s->n_super, "cc_sizes: gid_base does not match shn_cc_merge's n_super");
  cudaStream_t st = c->stream;
  CUDA_CHECK(cudaMemsetAsync(d_sizes, 0, std::max<uint64_t>(s->n_final, 1) * 8, st));
  if (s->n_local == 0) return;
  CcView cc = cc_view(c, s, gid_base);
  DevBuf local;
  local.reserve(s->n_local * 4);
  CUDA_CHECK(cudaMemsetAsync(local.p, 0, s->n_local * 4, st));
  ProfScope ps(c, "cc_sizes", 2);
  local_sizes_kernel<<<shn_grid(s->n_slots, kBlock), kBlock, 0, st>>>(cc.parent, cc.root_id, s->n_slots,
                                                                     local.as<uint32_t>());
  KERNEL_CHECK();
  final_sizes_kernel<<<shn_grid(s->n_local, kBlock), kBlock, 0, st>>>(
      local.as<uint32_t>(), s->n_local, (uint32_t)gid_base, s->final_of_super.as<uint32_t>(),
      reinterpret_cast<unsigned long long*>(d_sizes));
  KERNEL_CHECK();
  CUDA_CHECK(cudaStreamSynchronize(st));  // `local` returns to the pool
}

void cc_route(shn_ctx* c, const uint32_t* d_owner_of_final, uint64_t gid_base, uint32_t nranks,
              uint64_t* h_counts, void* send) {
  ShardState* s = shard_of(c);
  SHN_CHECK(c->gline_dev != nullptr, "the table was not built with shn_table_build_records");
  ByCompSrc src{cc_view(c, s, gid_base), s->final_of_super.as<uint32_t>(), d_owner_of_final, c->gline_dev};
  route(c, send ? "cc_route_fill" : "cc_route_count", src, s->n_slots, nranks, h_counts, send);
}

void cc_free(shn_ctx* c) {
  if (c->shard && c->shard_free) c->shard_free(c);
  c->shard = nullptr;
}

}  // namespace SHN_NS
