// a3-a9: seeds, greedy walks, shape filter, duplicate filter, contig graph, components
// (run_correction, extension_correction.py:334-450) on the K1-mer table of table.cu.
//
// Exactness argument (DESIGN.md "walks"): a walk only ever probes successors/predecessors of its
// current K1-mer, so it stays inside one connected component of the K1-mer successor graph and
// two walks in different components never read or write the same table slot.  The sequential
// pop order of the reference therefore only matters *within* a component: we label components
// with a lock-free union-find, give every component to one warp, and that warp replays its
// component's seeds in global pop order (weight desc, later input line first).  The union of all
// per-component replays is bit-identical to the reference's single sequential loop.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <thread>

#include "common.cuh"
#include "impls.h"
#include "selfjoin.cuh"
#include "table_dev.cuh"
#include "uf_dev.cuh"

namespace SHN_NS {

struct L3State {
  uint32_t min_weight = 0, min_length = 0;
  shn_l3_sizes sz;
  // per started walk, pop order
  DevBuf w_seed_slot;   // uint32
  DevBuf w_nl, w_nr;    // uint32
  DevBuf w_totwt;       // uint64
  DevBuf w_logstart;    // uint64
  DevBuf walk_log;      // uint8 base codes, one per traversed K1-mer (seed entries unused)
  // candidates (walks passing length + hyperbola, pop order): left by l3_walks for l3_filter / export
  uint64_t n_cand = 0;
  DevBuf cand_walk;      // uint32 walk index per candidate
  DevBuf cand_off;       // uint64 [n_cand+1] offsets into cand_codes
  DevBuf cand_codes;     // uint8 base codes of the candidate contigs
  std::vector<uint64_t> h_cand_off;
  bool foreign = false;  // l3_filter ran on candidates supplied by the caller (sharded path)
  std::vector<uint32_t> h_cand_walk;  // walk index of every candidate (passes length+hyperbola)
  std::vector<uint8_t> h_cand_dup;    // duplicate_check() result per candidate
  std::vector<uint8_t> h_cand_acc;    // accepted per candidate
  // accepted contigs
  DevBuf contig_codes;   // uint8 codes of all candidates, later compacted to accepted
  DevBuf contig_offs;    // uint64 [n_contigs+1]
  std::vector<uint64_t> h_contig_offs;
  DevBuf allowed_keys, allowed_w;
  PairTable edges;       // hi=b, lo=a, count=weight, min_i=first pos in b
  DevBuf labels;         // uint32 [n_contigs+1]
  L3State() { memset(&sz, 0, sizeof(sz)); }
};

namespace {

constexpr int kBlock = 256;

// ---- seeds ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
    seed_count_kernel(const ShnSlot* __restrict__ slots, uint64_t n_slots, uint32_t min_weight,
                      unsigned long long* counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mine = 0;
  uint32_t wmax = 0;
  for (; i < n_slots; i += stride) {
    shn_key_t key;
    uint32_t wz, wi;
    table_load_slot(slots, i, &key, &wz, &wi);
    const bool seed = key != SHN_EMPTY && (wz & SHN_WEIGHT_MASK) >= min_weight;
    mine += seed ? 1 : 0;
    if (seed) wmax = max(wmax, wz & SHN_WEIGHT_MASK);
  }
  typedef cub::BlockReduce<unsigned long long, kBlock> BR;
  __shared__ typename BR::TempStorage tmp;
  unsigned long long tot = BR(tmp).Sum(mine);
  if (threadIdx.x == 0 && tot) atomicAdd(&counters[0], tot);
  // counters[1] = largest seed weight: the seed sort only looks at the bits that vary
  wmax = __reduce_max_sync(0xFFFFFFFFu, wmax);
  if ((threadIdx.x & 31) == 0 && wmax) atomicMax(&counters[1], (unsigned long long)wmax);
}

// (sort key, slot) for every K1-mer with weight >= min_weight; sort key ascending = pop order:
// weight descending, then first-occurrence index descending (stable ascending sort + pop()).
__global__ void __launch_bounds__(kBlock)
    seed_emit_kernel(ShnSlot* slots, uint64_t n_slots, uint32_t min_weight,
                     const uint64_t* __restrict__ gline, uint64_t* __restrict__ sortkey,
                     uint32_t* __restrict__ sslot, unsigned long long* cursor,
                     uint32_t* __restrict__ saved, uint32_t wmask, int ibits) {
  __shared__ unsigned long long block_base;
  __shared__ int warp_off[kBlock / 32];
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  shn_key_t skey = SHN_EMPTY;
  uint32_t wz = 0, first_idx = 0;
  if (i < n_slots) {
    table_load_slot(slots, i, &skey, &wz, &first_idx);
    // this pass also parks the idx word (see idx_park_kernel): one table read less
    saved[i] = first_idx;
    slots[i].idx = 0;
  }
  const uint32_t wt = wz & SHN_WEIGHT_MASK;
  bool is_seed = skey != SHN_EMPTY && wt >= min_weight;
  unsigned b = __ballot_sync(0xFFFFFFFFu, is_seed);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_off[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kBlock / 32; ++w) {
      int cc = warp_off[w];
      warp_off[w] = tot;
      tot += cc;
    }
    block_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (is_seed) {
    uint64_t o = block_base + warp_off[warp] + __popc(b & ((1u << lane) - 1u));
    // sharded tables: first_idx is a receive-buffer position, the global input line (34 bits) is
    // behind gline[]; weights have 30 bits
    // (both fields complemented: ascending key = weight descending, then index descending)
    const uint64_t imask = (1ull << ibits) - 1ull;
    sortkey[o] = ((uint64_t)(~wt & wmask) << ibits) | (~(gline ? gline[first_idx] : (uint64_t)first_idx) & imask);
    sslot[o] = (uint32_t)i;
  }
}

// ---- the idx word of every slot doubles as the walks' scratch word -----------------------------
// (first-occurrence index, needed again by shn_table_dump and the next shn_l3_run, is parked in a
// side array for the duration of the walks)
__global__ void __launch_bounds__(kBlock)
    idx_park_kernel(ShnSlot* slots, uint64_t n_slots, uint32_t* __restrict__ saved) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n_slots; i += stride) {
    saved[i] = slots[i].idx;
    slots[i].idx = 0;
  }
}
__global__ void __launch_bounds__(kBlock)
    idx_unpark_kernel(ShnSlot* slots, uint64_t n_slots, const uint32_t* __restrict__ saved) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n_slots; i += stride) slots[i].idx = saved[i];
}

// ---- raw components: lock-free union-find over table slots ---------------------------------
__global__ void __launch_bounds__(kBlock) uf_init_kernel(uint32_t* parent, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) parent[i] = (uint32_t)i;
}

// One thread per K1-mer: find its successors and link them (1.2 on average).  The four candidates
// share one home bucket (placement by the K-base prefix, common.cuh), so this is ONE probe sequence
// per K1-mer.
// The probes also yield the neighbour masks the walks prune their probes with (layout of the aux
// word: see "greedy walks" below): successor bit b of x = (x[1:] . b) exists, OR-ed by x's own
// thread; predecessor bit f of y = (f . y[:-1]) exists, OR-ed by the thread of that predecessor.
__global__ void __launch_bounds__(kBlock)
    uf_edges_kernel(ShnTableView t, uint32_t* parent, uint64_t n_slots, int k1) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  shn_key_t key = t.slots[i].key;
  if (key == SHN_EMPTY) {
    parent[i] = SHN_NONE32;  // the later passes tell free slots from the parent array alone
    return;
  }
  const uint32_t first = (uint32_t)(key >> (2 * (k1 - 1))) & 3u;
  uint64_t slot[4];
  uint32_t w[4];
  const uint32_t sm = table_find_successors(t, key, k1, slot, w);
  // k-th existing successor of every lane together, whatever its base: a loop over the four bases
  // would run four serialised unions with ~3 of 32 lanes active each (measured: 62 % of the kernel)
  for (uint32_t m = sm; m; m &= m - 1) {
    const int b = __ffs(m) - 1;
    const uint64_t s = b == 0 ? slot[0] : (b == 1 ? slot[1] : (b == 2 ? slot[2] : slot[3]));
    atomicOr(&t.slots[s].idx, 1u << (28 + first));  // the bucket of s was just read: an L2 hit
    if (s != i) uf_union(parent, (uint32_t)i, (uint32_t)s);
  }
  if (sm) atomicOr(&t.slots[i].idx, sm << 24);
}

// parent[i] <- root for occupied slots; roots get flag 1 (free slots: parent = SHN_NONE32)
__global__ void __launch_bounds__(kBlock)
    uf_flatten_kernel(uint32_t* parent, uint64_t n_slots, uint32_t* __restrict__ is_root) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const uint32_t p = __ldcg(&parent[i]);
  const bool occ = p != SHN_NONE32;
  uint32_t root = (uint32_t)i;
  if (occ && p != (uint32_t)i) root = uf_find_ro(parent, p);
  is_root[i] = (occ && root == (uint32_t)i) ? 1u : 0u;
  if (occ && root != p) parent[i] = root;
}

// comp id of every occupied slot + node count per component
__global__ void __launch_bounds__(kBlock)
    comp_count_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ root_id,
                      uint64_t n_slots, uint32_t* __restrict__ comp_nodes) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // one atomic per distinct component of a warp: neighbouring slots mostly share their component
  // (minimizer regions), and one atomic per slot serialises on the large components (config 4:
  // 39.5 ms for this pass)
  uint32_t r = SHN_NONE32;
  if (i < n_slots) {
    const uint32_t p = parent[i];  // the root after flatten
    if (p != SHN_NONE32) r = root_id[p];
  }
  const unsigned same = __match_any_sync(0xFFFFFFFFu, r);
  if (r != SHN_NONE32 && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&comp_nodes[r], (uint32_t)__popc(same));
}

// the same with a block-private histogram in shared memory (n_comps * 4 bytes of dynamic smem):
// 10^8 atomics on a few thousand counters serialise in L2, shared-memory atomics do not
__global__ void __launch_bounds__(kBlock)
    comp_count_smem_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ root_id,
                           uint64_t n_slots, uint32_t n_comps, uint32_t* __restrict__ comp_nodes) {
  extern __shared__ uint32_t hist[];
  for (uint32_t k = threadIdx.x; k < n_comps; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n_slots; i += stride) {
    const uint32_t p = parent[i];
    if (p != SHN_NONE32) atomicAdd(&hist[root_id[p]], 1u);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < n_comps; k += blockDim.x)
    if (hist[k]) atomicAdd(&comp_nodes[k], hist[k]);
}

__global__ void __launch_bounds__(kBlock)
    seed_comp_kernel(const uint32_t* __restrict__ seed_slot, const uint32_t* __restrict__ parent,
                     const uint32_t* __restrict__ root_id, uint64_t n_seeds,
                     uint32_t* __restrict__ seed_comp, uint32_t* __restrict__ rank) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  uint32_t cid = root_id[parent[seed_slot[i]]];
  seed_comp[i] = cid;
  rank[i] = (uint32_t)i;
}

// seeds per component and their offsets from the component-sorted seed list: two binary searches per
// component instead of one atomic per seed (87 M atomics on a few thousand counters)
__global__ void __launch_bounds__(kBlock)
    seed_offsets_kernel(const uint32_t* __restrict__ seed_comp_s, uint64_t n_seeds, uint32_t n_comps,
                        uint64_t* __restrict__ seed_off, uint32_t* __restrict__ comp_seeds) {
  uint32_t cid = blockIdx.x * blockDim.x + threadIdx.x;
  if (cid > n_comps) return;
  uint64_t bound[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint64_t want = (uint64_t)cid + k;  // first seed whose component is >= want
    uint64_t lo = 0, up = n_seeds;
    while (lo < up) {
      const uint64_t mid = (lo + up) >> 1;
      if ((uint64_t)seed_comp_s[mid] < want)
        lo = mid + 1;
      else
        up = mid;
    }
    bound[k] = lo;
  }
  seed_off[cid] = bound[0];
  comp_seeds[cid] = cid < n_comps ? (uint32_t)(bound[1] - bound[0]) : 0u;
}

__global__ void __launch_bounds__(kBlock)
    gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n,
                      uint32_t* __restrict__ dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(kBlock)
    gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n,
                      uint64_t* __restrict__ dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(kBlock)
    comp_work_kernel(const uint32_t* __restrict__ comp_nodes, const uint32_t* __restrict__ comp_seeds,
                     uint32_t n_comps, uint32_t* __restrict__ work, uint32_t* __restrict__ ids,
                     unsigned long long* counters) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = false;
  if (i < n_comps) {
    active = comp_seeds[i] > 0;
    work[i] = active ? comp_nodes[i] : 0u;
    ids[i] = i;
  }
  int tot = __syncthreads_count(active);
  if (threadIdx.x == 0 && tot) atomicAdd(&counters[0], (unsigned long long)tot);
}

__global__ void __launch_bounds__(kBlock)
    gather_walks_kernel(const uint32_t* __restrict__ sel, uint64_t n_walks,
                        const uint32_t* __restrict__ seed_slot, const uint32_t* __restrict__ nl_r,
                        const uint32_t* __restrict__ nr_r, const uint64_t* __restrict__ tot_r,
                        const uint64_t* __restrict__ ls_r, int k1, uint32_t min_length,
                        uint32_t* __restrict__ w_slot, uint32_t* __restrict__ w_nl,
                        uint32_t* __restrict__ w_nr, uint64_t* __restrict__ w_tot,
                        uint64_t* __restrict__ w_ls, uint8_t* __restrict__ is_long) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walks) return;
  uint32_t r = sel[w];
  uint32_t nl = nl_r[r], nr = nr_r[r];
  w_slot[w] = seed_slot[r];
  w_nl[w] = nl;
  w_nr[w] = nr;
  w_tot[w] = tot_r[r];
  w_ls[w] = ls_r[r];
  is_long[w] = ((uint64_t)nl + nr + (uint64_t)k1 >= (uint64_t)min_length) ? 1 : 0;  // len(contig)
}

// a5 on the device: the hyperbola test (extension_correction.py:356-361) in doubles.  The device
// pow() may differ from the host libm in the last bits, so a walk whose two sides are closer than
// tol_rel (1e-9, relative) is not decided here: flag 2 = the host evaluates it with the
// reference's own expression (passes_shape).  flag 1 = passes, 0 = fails.
__global__ void __launch_bounds__(kBlock)
    shape_kernel(const uint32_t* __restrict__ w_nl, const uint32_t* __restrict__ w_nr,
                 const uint64_t* __restrict__ w_tot, uint64_t n_walks, int k1, uint32_t min_weight,
                 uint32_t min_length, double tol_rel, uint8_t* __restrict__ flag,
                 uint32_t* __restrict__ borderline, unsigned long long* counters) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walks) return;
  const uint64_t tot_kmer = (uint64_t)w_nl[w] + w_nr[w] + 1;
  const uint64_t length = tot_kmer + (uint64_t)k1 - 1;
  uint8_t f = 0;
  if (length >= min_length) {
    const double avg_wt = (double)w_tot[w] / (double)(tot_kmer > 1 ? tot_kmer : 1);
    const double lhs = (double)length * pow(avg_wt, 0.25);
    const double rhs = (double)(2ull * min_length) * pow((double)min_weight, 0.25);
    const double tol = tol_rel * rhs;
    if (lhs > rhs + tol) f = 1;
    else if (lhs >= rhs - tol) {
      f = 2;
      borderline[atomicAdd(&counters[0], 1ull)] = (uint32_t)w;  // at most n_walks entries
    }
  }
  flag[w] = f;
}

__global__ void __launch_bounds__(kBlock)
    patch_flags_kernel(const uint32_t* __restrict__ idx, const uint8_t* __restrict__ val, uint64_t n,
                       uint8_t* __restrict__ flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[idx[i]] = val[i];
}

__global__ void __launch_bounds__(kBlock)
    cand_len_kernel(const uint32_t* __restrict__ cand_walk, uint64_t n_cand,
                    const uint32_t* __restrict__ w_nl, const uint32_t* __restrict__ w_nr, int k1,
                    uint64_t* __restrict__ len) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_cand) return;
  uint64_t v = 0;
  if (j < n_cand) {
    const uint32_t w = cand_walk[j];
    v = (uint64_t)w_nl[w] + w_nr[w] + (uint64_t)k1;
  }
  len[j] = v;
}

// ---- greedy walks -----------------------------------------------------------------------
// The walks read the K1-mer table in place (keys and weights never change) and keep everything
// that does change in one 32-bit word per slot, `aux`:
//   bits 31..28  predecessor mask: bit f = K1-mer (f . x[:-1]) exists      } written by
//   bits 27..24  successor mask:   bit b = K1-mer (x[1:] . b) exists        } uf_edges_kernel
//   bit  23      traversed (extension_correction.py:235,347)
//   bits 22..0   claim stamp of the speculative walks (0 = unclaimed)
// The masks prune the probes: of the four successors (predecessors) of the current K1-mer only
// the existing ones are fetched (1.2 on average), and only their children are probed blind for
// the second step of a round.
constexpr int kAuxSuccShift = 24;
constexpr uint32_t kAuxTraversed = 1u << 23;
constexpr uint32_t kAuxStampMask = 0x007FFFFFu;

struct WalkArgs {
  ShnTableView tv;              // the K1-mer table; the walks only write the idx (= aux) words
  uint32_t lookahead;           // walk_kernel: 0 = one step per round (no second-level probes)
  int k1;
  uint32_t n_comps;
  const uint32_t* comp_order;   // components sorted by node count (descending)
  const uint64_t* seed_off;     // [n_comps+1] into ranks_by_comp
  const uint32_t* ranks_by_comp;  // seed ranks grouped by component, ascending inside a group
  const uint32_t* slots_by_comp;  // table slot of the same seeds
  const uint64_t* log_off;      // [n_comps+1] into walk_log
  uint8_t* walk_log;
  // per seed rank outputs
  uint8_t* started;
  uint32_t* nl;
  uint32_t* nr;
  uint64_t* totwt;
  uint64_t* logstart;
  unsigned long long* counters;  // [0]=traversed [1]=max rounds of a warp [2]=log overflow [3]=windows
                                 // [4]=speculative path buffer overflow
  unsigned long long* trace;     // optional: per warp {end time ns, rounds, cycles} (SHN_WALK_TRACE)
};

// Minimizer hash of the candidates' PREFIXES (the 12-mers at offsets 1..n of a candidate, n = k1 - 12;
// offsets count from the last base) without a loop per candidate: the candidates are `cur` shifted by
// one (level 1) or two (level 2) bases and share most 12-mers with it.  Lane p hashes the 12-mer of
// cur at offset p and two warp reductions give the shared minima (cur is warp-uniform, all lanes call
// this):
//   right extension: level 1 = cur's offsets [0, n-1] and nothing else, level 2 = [0, n-2] + one new
//   left extension:  level 1 = cur's offsets [2, n] + one new,          level 2 = [3, n] + two new
// walk_cand_min() adds the new 12-mers of the candidate itself.
__device__ __forceinline__ uint32_t walk_shared_min(const ShnTableView& tv, shn_key_t cur, int k1, int dir,
                                                    int lvl, int lane) {
  if (tv.n_regions == 0) return 0u;
  const int n = k1 - kRegionM;
  const uint32_t hp = lane <= n ? shn_mmer_hash(cur, lane) : 0xFFFFFFFFu;
  const int lo1 = dir == 0 ? 0 : 2, hi1 = dir == 0 ? n - 1 : n;
  const int lo2 = dir == 0 ? 0 : 3, hi2 = dir == 0 ? n - 2 : n;
  const uint32_t m1 = __reduce_min_sync(0xFFFFFFFFu, (lane >= lo1 && lane <= hi1) ? hp : 0xFFFFFFFFu);
  const uint32_t m2 = __reduce_min_sync(0xFFFFFFFFu, (lane >= lo2 && lane <= hi2) ? hp : 0xFFFFFFFFu);
  return lvl == 2 ? m2 : m1;
}
__device__ __forceinline__ uint32_t walk_cand_min(const ShnTableView& tv, shn_key_t cand, int k1, int dir, int lvl,
                                                  uint32_t shared) {
  if (tv.n_regions == 0) return 0u;
  const int n = k1 - kRegionM;  // >= 1 when the table has regions
  uint32_t m = shared;
  if (dir == 0) {  // the appended bases are not part of the prefix of a level-1 candidate
    if (lvl == 2) m = min(m, shn_mmer_hash(cand, 1));
  } else {         // the prepended bases sit at the start of the candidate
    m = min(m, shn_mmer_hash(cand, n));
    if (lvl == 2 && n >= 2) m = min(m, shn_mmer_hash(cand, n - 1));
  }
  return m;
}

// One probe of the walks: the candidate's home bucket is already loaded.
// state: 1 found (slot, raw weight word, aux word), 0 absent, -1 continues in bucket *nextb.
__device__ __forceinline__ int walk_resolve(const ShnTableView& tv, const ShnBucket& bk, shn_key_t cand,
                                            uint64_t hb, uint64_t* cslot, uint32_t* wraw, uint32_t* caux,
                                            uint64_t* nextb) {
  int jj = 0;
  const int state = table_match_bucket2(bk, cand, &jj, wraw, caux);
  *nextb = (hb + 1 == tv.n_buckets) ? 0 : hb + 1;
  if (state == 1) *cslot = SHN_BSLOTS * hb + jj;
  return state;
}
// the rare continuation of a probe sequence (the key was displaced from its home bucket)
#ifdef SHN_WALK_STATS
#define SHN_CHASE_COUNT(n) (n)
__device__ unsigned long long g_chase_iters = 0, g_chase_calls = 0;
#else
#define SHN_CHASE_COUNT(n) 0
#endif
__device__ __forceinline__ int walk_chase(const ShnTableView& tv, shn_key_t cand, uint64_t* cslot,
                                          uint32_t* wraw, uint32_t* caux, uint64_t* nextb) {
#ifdef SHN_WALK_STATS
  atomicAdd(&g_chase_calls, 1ull);
#endif
  for (;;) {
#ifdef SHN_WALK_STATS
    atomicAdd(&g_chase_iters, 1ull);
#endif
    ShnBucket bk;
    table_load_bucket(tv, *nextb, &bk);
    const uint64_t hb = *nextb;
    const int state = walk_resolve(tv, bk, cand, hb, cslot, wraw, caux, nextb);
    if (state >= 0) return state;
  }
}

// One WARP per component.  Seeds are scanned 32 at a time (one coalesced load of the slot
// indices, one gather of the aux words, a ballot).  A walk advances TWO K1-mers per memory round
// trip: lanes 0..3 probe the existing successors (predecessors) c_b of the current node and, in the
// same round, lanes 4..19 probe the second-level nodes succ(c_b, b'); after the first arg-max
// picks c_w the second step is decided from lanes 4+4w..7+4w, whose traversed flags are still
// exact except for c_w itself (marked after the fetch; compared by key).  The arg-max keeps the
// reference's tie order A,G,C,T.  The kernel is bound by the dependent chain of one round
// (~1 DRAM round trip + ~250 instructions), not by issue slots or bandwidth.
constexpr int kWalkBlock = 128;
#ifdef SHN_WIDE
constexpr int kWalkK1 = 33;  // K = 32: the one width that needs 128-bit keys in practice
#else
constexpr int kWalkK1 = 25;  // K = 24, the reference's default (shannon.py): its own kernel instances
#endif
#ifndef SHN_WALK_BLOCKS_PER_SM
#define SHN_WALK_BLOCKS_PER_SM 8
#endif
constexpr int kWalkBlocksPerSM = SHN_WALK_BLOCKS_PER_SM;  // 8 -> 64 registers: co-resident with speculative CTAs

// kK1 != 0: K1 known at compile time (constant masks and shifts; the default K = 24 gets its own
// instance); kTrace: per-round clock reads for SHN_WALK_TRACE.
template <int kK1, bool kTrace>
__global__ void __launch_bounds__(kWalkBlock, kWalkBlocksPerSM) walk_kernel(WalkArgs a) {
  const int k1 = kK1 ? kK1 : a.k1;
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= a.n_comps) return;  // whole warps only
  const uint32_t comp = a.comp_order[warp];
  const uint64_t s_begin = a.seed_off[comp], s_end = a.seed_off[comp + 1];
  uint64_t lp = a.log_off[comp];
  const uint64_t le = a.log_off[comp + 1];
  const shn_key_t mask = shn_key_mask(k1);
  const int top = 2 * (k1 - 1);
  const ShnTableView tv = a.tv;
  unsigned long long rounds = 0, traversed = 0;
  long long mem_cycles = 0, t_begin = clock64();
  const bool tracing = kTrace && a.trace != nullptr;
  bool overflow = false;
  const bool look = a.lookahead != 0;
  // role of this lane inside a round: level 1 (lanes 0..3), level 2 (lanes 4..19), idle
  const int lvl = lane < 4 ? 1 : ((look && lane < 20) ? 2 : 0);
  const shn_key_t b1 = lvl == 1 ? lane : ((lane - 4) >> 2);  // first appended base
  const shn_key_t b2 = (lane - 4) & 3;                       // second appended base (level 2)

  for (uint64_t base = s_begin; base < s_end; base += 32) {
    // ---- scan 32 seeds of this component (pop order) -------------------------------------
    const uint64_t si = base + lane;
    const bool have = si < s_end;
    uint32_t my_rank = 0, my_slot = 0, my_aux = kAuxTraversed;
    if (have) {
      my_rank = a.ranks_by_comp[si];
      my_slot = a.slots_by_comp[si];
      my_aux = __ldcg(&tv.slots[my_slot].idx);
    }
    unsigned pending = __ballot_sync(FULL, have && !(my_aux & kAuxTraversed));
    while (pending) {
      const int j = __ffs(pending) - 1;
      pending &= pending - 1;
      const uint32_t slot = __shfl_sync(FULL, my_slot, j);
      const uint32_t rank = __shfl_sync(FULL, my_rank, j);
      // fresh look: an earlier walk of this batch may have traversed it meanwhile (:346)
      const uint32_t seed_aux = __ldcg(&tv.slots[slot].idx);
      if (seed_aux & kAuxTraversed) continue;  // warp-uniform
      shn_key_t seed_key;
      uint32_t seed_w, seed_i;
      table_load_slot(tv.slots, slot, &seed_key, &seed_w, &seed_i);
      if (lane == 0) {
        tv.slots[slot].idx = seed_aux | kAuxTraversed;      // traversed.add(start_kmer), :347
        if (lp < le) a.walk_log[lp] = 0xFF;        // the seed's own (unused) log entry
      }
      overflow |= lp >= le;
      const uint64_t my_log = lp;
      ++lp;
      ++traversed;
      uint64_t tot = seed_w & SHN_WEIGHT_MASK;
      uint32_t n_dir[2] = {0, 0};
      __syncwarp();
#pragma unroll 1
      for (int dir = 0; dir < 2; ++dir) {  // right extension first, then left (:349-350)
        shn_key_t cur = seed_key;
        uint32_t cur_aux = seed_aux;  // neighbour masks of `cur`
        for (;;) {
          const uint32_t cm = (cur_aux >> (kAuxSuccShift + 4 * dir)) & 0xFu;
          if (cm == 0) break;  // warp-uniform: dead end
          ++rounds;
          const bool act = lvl != 0 && ((cm >> (int)b1) & 1u);
          // candidate keys in the reference's tie order A,G,C,T = codes 0..3 (:10,229)
          shn_key_t cand = 0;
          uint64_t cslot = ~0ull;
          uint32_t wraw = 0, caux = 0;
          int state = 0;       // 1 found, 0 absent, -1 undecided after the home bucket
          uint64_t nextb = 0;  // where an undecided lane would continue
          const uint32_t cand_min = walk_shared_min(tv, cur, k1, dir, lvl, lane);
          if (act) {
            if (dir == 0) {
              cand = ((cur << 2) & mask) | b1;
              if (lvl == 2) cand = ((cand << 2) & mask) | b2;
            } else {
              cand = (cur >> 2) | (b1 << top);
              if (lvl == 2) cand = (cand >> 2) | (b2 << top);
            }
            const uint64_t hb = tv.bucket_with_min(cand, walk_cand_min(tv, cand, k1, dir, lvl, cand_min));
            ShnBucket bk0;
            long long tm0 = 0;
            if (tracing) tm0 = clock64();
            table_load_bucket(tv, hb, &bk0);
            if (tracing) {  // wait for the data here so the cycles are attributed to memory
              volatile uint64_t sink = bk0.w[0] ^ bk0.w[4];
              (void)sink;
              mem_cycles += clock64() - tm0;
            }
            state = walk_resolve(tv, bk0, cand, hb, &cslot, &wraw, &caux, &nextb);
          }
          // only the lanes a decision actually depends on pay for longer probe sequences
          if (lvl == 1 && state < 0) state = walk_chase(tv, cand, &cslot, &wraw, &caux, &nextb);
          bool ok = state == 1 && !(caux & kAuxTraversed);
          // ---- first step: arg-max weight over lanes 0..3, first of equals wins (:159-166) ---
          // score = (weight, 3 - code) + 1 in 32 bits (weights are < 2^30 - 1), 0 = no candidate
          uint32_t score = ok ? ((((wraw & SHN_WEIGHT_MASK) << 2) | (uint32_t)(3 - (lane & 3))) + 1u) : 0u;
          uint32_t m = max(score, __shfl_xor_sync(FULL, score, 1));
          m = max(m, __shfl_xor_sync(FULL, m, 2));  // max of my aligned group of four lanes
          const uint32_t s1 = __shfl_sync(FULL, m, 0);
          if (s1 == 0) break;  // warp-uniform: no extension
          const int w1 = 3 - (int)((s1 - 1u) & 3u);
          const uint32_t bw1 = (s1 - 1u) >> 2;
          const shn_key_t c1 = dir == 0 ? (((cur << 2) & mask) | (shn_key_t)w1)
                                        : ((cur >> 2) | ((shn_key_t)w1 << top));
          if (lane == w1) tv.slots[cslot].idx = caux | kAuxTraversed;  // traversed.add(last), :235
          if (lane == 0 && lp < le) a.walk_log[lp] = (uint8_t)w1;
          overflow |= lp >= le;
          ++lp;
          ++traversed;
          tot += bw1;
          ++n_dir[dir];
          if (!look) {  // warp-uniform: no prefetched level, continue from c1
            cur = c1;
            cur_aux = __shfl_sync(FULL, caux, w1);
            __syncwarp();
            continue;
          }
          // ---- second step from the prefetched level: group 4+4*w1; c1 is traversed by now ---
          const int g2 = 4 + 4 * w1;
          // c1's own neighbour mask (in the aux word lane w1 just fetched) says which of its four
          // children exist: a blind probe that is undecided after its home bucket continues only
          // if its child exists (most undecided probes are absent keys behind a full bucket)
          const uint32_t c1mask = (__shfl_sync(FULL, caux, w1) >> (kAuxSuccShift + 4 * dir)) & 0xFu;
          if (lane >= g2 && lane < g2 + 4 && state < 0) {
            if ((c1mask >> (lane - g2)) & 1u) {
              state = walk_chase(tv, cand, &cslot, &wraw, &caux, &nextb);
              ok = state == 1 && !(caux & kAuxTraversed);
            } else {
              state = 0;
              ok = false;
            }
          }
          ok = ok && cand != c1;
          score = ok ? ((((wraw & SHN_WEIGHT_MASK) << 2) | (uint32_t)(3 - (lane & 3))) + 1u) : 0u;
          m = max(score, __shfl_xor_sync(FULL, score, 1));
          m = max(m, __shfl_xor_sync(FULL, m, 2));
          const uint32_t s2 = __shfl_sync(FULL, m, g2);
          if (s2 == 0) {
            __syncwarp();
            break;  // warp-uniform: the walk ends at c1 in this direction
          }
          const int w2 = 3 - (int)((s2 - 1u) & 3u);
          const uint32_t bw2 = (s2 - 1u) >> 2;
          cur = dir == 0 ? (((c1 << 2) & mask) | (shn_key_t)w2) : ((c1 >> 2) | ((shn_key_t)w2 << top));
          cur_aux = __shfl_sync(FULL, caux, g2 + w2);
          if (lane == g2 + w2) tv.slots[cslot].idx = caux | kAuxTraversed;
          if (lane == 0 && lp < le) a.walk_log[lp] = (uint8_t)w2;
          overflow |= lp >= le;
          ++lp;
          ++traversed;
          tot += bw2;
          ++n_dir[dir];
          __syncwarp();  // orders the flag stores before the next round of probes
        }
      }
      if (lane == 0) {
        a.started[rank] = 1;
        a.nr[rank] = n_dir[0];
        a.nl[rank] = n_dir[1];
        a.totwt[rank] = tot;
        a.logstart[rank] = my_log;
      }
    }
  }
  if (lane == 0) {
    atomicAdd(&a.counters[0], traversed);
    atomicMax(&a.counters[1], rounds);
    if (overflow) atomicAdd(&a.counters[2], 1ull);
    if (a.trace) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.trace[3 * (uint64_t)warp] = tns;
      a.trace[3 * (uint64_t)warp + 1] = rounds;
      a.trace[3 * (uint64_t)warp + 2] = ((unsigned long long)(clock64() - t_begin) << 32) |
                                        (unsigned long long)(mem_cycles >> 8);
    }
  }
}

// ---- speculative windowed walks for the large components ------------------------------------
// The serial replay of one component is the critical path of the whole stage (latency of ~10^5
// dependent probe rounds).  Large components are therefore given a CTA of kSpecWarps warps that
// run the next kSpecWindow untraversed seeds of the pop order CONCURRENTLY and commit them IN ORDER
// ("deterministic reservations"):
//   * the right extension of the walk at window position p carries the stamp 2 (kSpecWindow - p),
//     its left extension that stamp minus one (earlier seed = larger stamps; the two halves are
//     separate units of work, see phase 2), and both claim a K1-mer with atomicMax on the slot's aux
//     word (stamp in the low bits; the bits above it are the same for every claimant of an
//     untraversed slot);
//   * a candidate is blocked for a half walk iff it is committed-traversed or carries a stamp >= its
//     own: an EARLIER seed, the half itself, or -- for a left half -- the right half of its own
//     seed; smaller stamps are ignored and overwritten (stolen);
//   * after the window every path is re-read: a walk is intact iff every slot of its right path
//     still carries the right stamp and every slot of its left path the left stamp.  The longest prefix of intact walks is exactly what the sequential loop would have
//     produced (an intact walk only ever yielded to walks before it, all of which are intact and
//     final; nothing it examined-and-rejected can matter): those are
//     committed (traversed bits, log, metas); the others clear their stamps and are retried in
//     the next window, which starts at the first uncommitted seed.  The first walk of a window is
//     always intact, so every window makes progress.  After a window every untraversed slot is
//     back to stamp 0, so stamps only ever order the walks of one window;
//   * most seeds of a window lie on a chain that an earlier seed of the same window walks over.
//     Such a walk is not a misspeculation: if its SEED ends up stamped by a committed earlier walk
//     the sequential loop would have found the seed traversed and skipped it (:346), so the walk is
//     dropped (SKIP) and the prefix goes on.  Its other, phantom claims may have blocked later
//     walks, so every walk records which window positions ever blocked it and commits only if all
//     of them committed.
// Warps per CTA: 16 for the largest components (their serial chain is the critical path), 8 for
// the next tier (half the resident-warp budget per component, so three times as many components
// can walk speculatively at once); see l3_run.
#ifndef SHN_SPEC_WINDOW
#define SHN_SPEC_WINDOW 32
#endif
// Seeds per window (<= 64: window positions are bits of a 64-bit blocker mask).  The warps of the
// CTA pull the window's half walks from a shared counter, so one long extension occupies one warp
// while the others work through the many short ones (98 % of the seeds of a large component are
// found traversed, or are taken by an earlier walk of the same window).
constexpr int kSpecWindow = SHN_SPEC_WINDOW;
#ifndef SHN_SPEC_CTAS_PER_SM
#define SHN_SPEC_CTAS_PER_SM 2
#endif
constexpr int kSpecCtasPerSM = SHN_SPEC_CTAS_PER_SM;  // x 16 warps: 64 registers per thread
#ifndef SHN_SPEC16_PER_SM
#define SHN_SPEC16_PER_SM 2
#endif
constexpr int kSpec16PerSM = SHN_SPEC16_PER_SM;  // register budget of the 16-warp tier (1 -> 128 registers)

struct SpecArgs {
  WalkArgs w;
  uint32_t* path_slot;       // per component: kSpecWindow buffers of (node count) entries
  uint8_t* path_base;
  const uint64_t* path_off;  // [n_spec + 1] first entry of every component's buffers
  unsigned long long* phase_ns;  // optional (SHN_WALK_TRACE): per CTA ns in {collect, walk, check, resolve, commit}
};

template <int kSpecWarps, int kK1>
__global__ void __launch_bounds__(kSpecWarps * 32, kSpecWarps == 16 ? kSpec16PerSM : kSpecCtasPerSM * 16 / kSpecWarps)
    walk_spec_kernel(SpecArgs sa) {
  const WalkArgs& a = sa.w;
  const int k1 = kK1 ? kK1 : a.k1;
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t comp = a.comp_order[blockIdx.x];
  const uint64_t s_begin = a.seed_off[comp], s_end = a.seed_off[comp + 1];
  const uint64_t le = a.log_off[comp + 1];
  const shn_key_t mask = shn_key_mask(k1);
  const int top = 2 * (k1 - 1);
  const ShnTableView tv = a.tv;
  const uint64_t path_cap = (sa.path_off[blockIdx.x + 1] - sa.path_off[blockIdx.x]) / kSpecWindow;
  uint32_t* const cta_path_slot = sa.path_slot + sa.path_off[blockIdx.x];
  uint8_t* const cta_path_base = sa.path_base + sa.path_off[blockIdx.x];
  const int lvl = lane < 4 ? 1 : (lane < 20 ? 2 : 0);
  const shn_key_t b1 = lvl == 1 ? lane : ((lane - 4) >> 2);
  const shn_key_t b2 = (lane - 4) & 3;

  __shared__ uint64_t sh_cursor, sh_cursor_after, sh_lp;
  __shared__ uint32_t sh_win_pos[kSpecWindow], sh_win_slot[kSpecWindow];
  __shared__ uint32_t sh_len_r[kSpecWindow], sh_len_l[kSpecWindow];  // path entries of the two halves (right: + the seed)
  __shared__ uint32_t sh_intact[kSpecWindow], sh_poison[kSpecWindow];
  __shared__ unsigned long long sh_tot_r[kSpecWindow], sh_tot_l[kSpecWindow];
  __shared__ unsigned long long sh_block[kSpecWindow];  // window positions whose stamps blocked this walk
  __shared__ uint32_t sh_thief[kSpecWindow];            // position holding this walk's seed (or none)
  __shared__ uint32_t sh_status[kSpecWindow];           // 1 commit, 2 skip
  __shared__ uint64_t sh_off[kSpecWindow];
  __shared__ uint32_t sh_win_n, sh_next, sh_P;
  __shared__ unsigned sh_mask[kSpecWarps];
  __shared__ uint64_t sh_scan;
  if (threadIdx.x == 0) {
    sh_cursor = s_begin;
    sh_lp = a.log_off[comp];
  }
  unsigned long long rounds = 0, traversed = 0, windows = 0, n_commit = 0, n_retry = 0;
  bool overflow = false, path_overflow = false;
  unsigned long long ph[6] = {0, 0, 0, 0, 0, 0}, t_ph = 0;
  const bool ph_on = sa.phase_ns != nullptr && threadIdx.x == 0;
#define SHN_PHASE(k)                                                  \
  if (ph_on) {                                                        \
    unsigned long long t_now;                                         \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));         \
    ph[k] += t_now - t_ph;                                            \
    t_ph = t_now;                                                     \
  }
  if (ph_on) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph));
  __syncthreads();

  for (;;) {
    // ---- 1. the next untraversed seeds in pop order: every warp scans 32 seeds of a stretch of
    // kSpecWarps x 32, the fresh ones are numbered across the warps (98 % of the seeds of a large
    // component are already traversed: one warp scanning alone took 9 % of the component's time)
    if (threadIdx.x == 0) {
      sh_win_n = 0;
      sh_next = 0;
      sh_scan = sh_cursor;
    }
    if (threadIdx.x < kSpecWindow) {
      sh_len_r[threadIdx.x] = sh_len_l[threadIdx.x] = 0u;
      sh_tot_r[threadIdx.x] = sh_tot_l[threadIdx.x] = 0ull;
      sh_poison[threadIdx.x] = 0u;
      sh_block[threadIdx.x] = 0ull;
    }
    __syncthreads();
    for (;;) {
      const uint64_t pos = sh_scan;
      const uint32_t n = sh_win_n;
      if (n >= (uint32_t)kSpecWindow || pos >= s_end) break;  // CTA-uniform
      const uint64_t si = pos + 32u * warp + lane;
      uint32_t slot = 0, av = kAuxTraversed;
      if (si < s_end) {
        slot = a.slots_by_comp[si];
        av = __ldcg(&tv.slots[slot].idx);
      }
      const unsigned fresh = __ballot_sync(FULL, si < s_end && !(av & kAuxTraversed));
      if (lane == 0) sh_mask[warp] = fresh;
      __syncthreads();
      uint32_t k = n;
      for (int v = 0; v < warp; ++v) k += __popc(sh_mask[v]);
      for (unsigned f = fresh; f && k < (uint32_t)kSpecWindow; f &= f - 1, ++k) {
        const int j = __ffs(f) - 1;
        const uint32_t sj = __shfl_sync(FULL, slot, j);
        if (lane == 0) {
          sh_win_pos[k] = (uint32_t)(pos + 32u * warp + j - s_begin);
          sh_win_slot[k] = sj;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t tot = n;
        uint64_t consumed = min((uint64_t)(32 * kSpecWarps), s_end - pos);
        for (int v = 0; v < kSpecWarps; ++v) {
          unsigned m = sh_mask[v];
          const uint32_t cnt = __popc(m);
          if (tot + cnt >= (uint32_t)kSpecWindow) {  // the window fills up inside warp v's 32 seeds
            for (uint32_t need = (uint32_t)kSpecWindow - tot; need > 1; --need) m &= m - 1;
            if (cnt) consumed = 32u * v + (__ffs(m) - 1) + 1;
            tot = kSpecWindow;
            break;
          }
          tot += cnt;
        }
        sh_win_n = tot;
        sh_scan = pos + consumed;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) sh_cursor_after = sh_scan;
    __syncthreads();
    SHN_PHASE(0)
    const uint32_t win_n = sh_win_n;
    if (win_n == 0) break;
    ++windows;

    // ---- 2. speculative walks: every warp pulls HALF walks until none is left ---------------------
    // The right and the left extension of a seed run as two units of work (usually on two warps): the
    // longest walk of a window is its critical path, and its two directions are independent except
    // that the left extension must treat what the right extension took as traversed.  The right half
    // carries the stamp 2(W - position), the left half that stamp minus one, so the left half yields
    // to its own right half like to any earlier seed; if the right half reaches a K1-mer the left half
    // claimed first (a cycle through the seed), it steals it, the left path is no longer intact and
    // the walk is retried.  The first seed of a window runs both directions on one warp, one after the
    // other, exactly like the sequential loop: it is always intact, so every window makes progress.
    const uint32_t n_units = 2 * win_n - 1;
    for (;;) {
      uint32_t u = 0;
      if (lane == 0) u = atomicAdd(&sh_next, 1u);
      u = __shfl_sync(FULL, u, 0);
      if (u >= n_units) break;
      const uint32_t ws = (u + 1) >> 1;
      const int d_begin = u == 0 ? 0 : (int)((u + 1) & 1u), d_end = u == 0 ? 2 : d_begin + 1;
      uint32_t* my_path_slot = cta_path_slot + (uint64_t)ws * path_cap;
      uint8_t* my_path_base = cta_path_base + (uint64_t)ws * path_cap;
      unsigned long long bl = 0;  // per lane: positions of earlier walks whose stamps blocked a candidate
      const uint32_t stamp_r = 2u * ((uint32_t)kSpecWindow - ws);  // earlier seed = larger stamp
      const uint32_t seed_slot = sh_win_slot[ws];
      const uint32_t seed_aux = __ldcg(&tv.slots[seed_slot].idx);
      bool go = true, poisoned = false;
      if (d_begin == 0) {
        // the right half claims the seed; an earlier walk of this window may already hold it
        uint32_t old = 0;
        if (lane == 0) old = atomicMax(&tv.slots[seed_slot].idx, (seed_aux & ~kAuxStampMask) | stamp_r);
        old = __shfl_sync(FULL, old, 0);
        go = (old & kAuxStampMask) < stamp_r;
        if (lane == 0) sh_len_r[ws] = go ? 1u : 0u;
      } else {
        go = (seed_aux & kAuxStampMask) <= stamp_r;  // (a later claim by an earlier seed: the right path fails its check)
      }
      if (go) {
        shn_key_t seed_key;
        uint32_t seed_w, seed_i;
        table_load_slot(tv.slots, seed_slot, &seed_key, &seed_w, &seed_i);
        if (d_begin == 0 && lane == 0) {
          my_path_slot[0] = seed_slot;
          my_path_base[0] = 0xFF;
        }
        // Claims are OPTIMISTIC: the atomicMax of a step is issued and the walk moves on; its
        // result is looked at one round later, under the latency of the next probes.  A claim
        // that lost against an earlier seed leaves that seed's stamp in the slot, so the walk can
        // never pass the intact check below; it is abandoned as soon as the loss is seen.
        uint32_t pend1 = 0, pend2 = 0;
#pragma unroll 1
        for (int dir = d_begin; dir < d_end && !poisoned; ++dir) {
          const uint32_t stamp = stamp_r - (uint32_t)dir;
          uint32_t len = dir == 0 ? 1u : 0u;  // entries of this half's path (the seed opens the right one)
          uint64_t tot = dir == 0 ? (uint64_t)(seed_w & SHN_WEIGHT_MASK) : 0ull;
          shn_key_t cur = seed_key;
          uint32_t cur_aux = seed_aux;
          for (;;) {
            const uint32_t cm = (cur_aux >> (kAuxSuccShift + 4 * dir)) & 0xFu;
            if (cm == 0) break;  // dead end
            ++rounds;
            const bool act = lvl != 0 && ((cm >> (int)b1) & 1u);
            shn_key_t cand = 0;
            uint64_t cslot = ~0ull, hb = 0;
            uint32_t wraw = 0, caux = 0;
            int state = 0;
            uint64_t nextb = 0;
            ShnBucket bk0;
            const uint32_t cand_min = walk_shared_min(tv, cur, k1, dir, lvl, lane);
            if (act) {
              if (dir == 0) {
                cand = ((cur << 2) & mask) | b1;
                if (lvl == 2) cand = ((cand << 2) & mask) | b2;
              } else {
                cand = (cur >> 2) | (b1 << top);
                if (lvl == 2) cand = (cand >> 2) | (b2 << top);
              }
              hb = tv.bucket_with_min(cand, walk_cand_min(tv, cand, k1, dir, lvl, cand_min));
              table_load_bucket(tv, hb, &bk0);
            }
            {  // the claims of the previous round, by now usually back from L2
              const bool lost = (pend1 & kAuxStampMask) >= stamp || (pend2 & kAuxStampMask) >= stamp;
              pend1 = pend2 = 0;
              if (__any_sync(FULL, lost)) {
                poisoned = true;
                break;
              }
            }
            if (act) state = walk_resolve(tv, bk0, cand, hb, &cslot, &wraw, &caux, &nextb);
            if (lvl == 1 && state < 0) state = walk_chase(tv, cand, &cslot, &wraw, &caux, &nextb);
            // blocked: committed-traversed, or stamped by an earlier seed, by this half or (left half)
            // by the right half of the same seed
            bool ok = state == 1 && !(caux & kAuxTraversed) && (caux & kAuxStampMask) < stamp;
            if (lvl == 1 && state == 1 && !(caux & kAuxTraversed) && (caux & kAuxStampMask) > stamp_r)
              bl |= 1ull << ((uint32_t)kSpecWindow - (((caux & kAuxStampMask) + 1u) >> 1));
            // ---- first step ------------------------------------------------------------------
            uint32_t score = ok ? ((((wraw & SHN_WEIGHT_MASK) << 2) | (uint32_t)(3 - (lane & 3))) + 1u) : 0u;
            uint32_t m = max(score, __shfl_xor_sync(FULL, score, 1));
            m = max(m, __shfl_xor_sync(FULL, m, 2));
            const uint32_t s1 = __shfl_sync(FULL, m, 0);
            if (s1 == 0) break;  // no extension in this direction
            const int w1 = 3 - (int)((s1 - 1u) & 3u);
            const uint32_t bw1 = (s1 - 1u) >> 2;
            if (lane == w1) pend1 = atomicMax(&tv.slots[cslot].idx, (caux & ~kAuxStampMask) | stamp);
            const shn_key_t c1 = dir == 0 ? (((cur << 2) & mask) | (shn_key_t)w1)
                                          : ((cur >> 2) | ((shn_key_t)w1 << top));
            const uint32_t c1slot = (uint32_t)__shfl_sync(FULL, (uint32_t)cslot, w1);
            if (lane == 0 && len < path_cap) {  // right half from the front, left half from the back
              const uint64_t e = dir == 0 ? (uint64_t)len : path_cap - 1 - len;
              my_path_slot[e] = c1slot;
              my_path_base[e] = (uint8_t)w1;
            }
            path_overflow |= len >= path_cap;
            ++len;
            tot += bw1;
            // ---- second step from the prefetched level ---------------------------------------
            const int g2 = 4 + 4 * w1;
            // (as in walk_kernel: c1's neighbour mask decides which undecided blind probes go on)
            const uint32_t c1mask = (__shfl_sync(FULL, caux, w1) >> (kAuxSuccShift + 4 * dir)) & 0xFu;
            if (lane >= g2 && lane < g2 + 4 && state < 0) {
              if ((c1mask >> (lane - g2)) & 1u) {
                state = walk_chase(tv, cand, &cslot, &wraw, &caux, &nextb);
                ok = state == 1 && !(caux & kAuxTraversed) && (caux & kAuxStampMask) < stamp;
              } else {
                state = 0;
                ok = false;
              }
            }
            ok = ok && cand != c1;
            if (lane >= g2 && lane < g2 + 4 && state == 1 && !(caux & kAuxTraversed) &&
                (caux & kAuxStampMask) > stamp_r)
              bl |= 1ull << ((uint32_t)kSpecWindow - (((caux & kAuxStampMask) + 1u) >> 1));
            score = ok ? ((((wraw & SHN_WEIGHT_MASK) << 2) | (uint32_t)(3 - (lane & 3))) + 1u) : 0u;
            m = max(score, __shfl_xor_sync(FULL, score, 1));
            m = max(m, __shfl_xor_sync(FULL, m, 2));
            const uint32_t s2 = __shfl_sync(FULL, m, g2);
            if (s2 == 0) {  // the walk ends at c1 in this direction
              __syncwarp();  // the claim on c1 is ordered before the first probes of the other direction
              break;
            }
            const int w2 = 3 - (int)((s2 - 1u) & 3u);
            const uint32_t bw2 = (s2 - 1u) >> 2;
            if (lane == g2 + w2) pend2 = atomicMax(&tv.slots[cslot].idx, (caux & ~kAuxStampMask) | stamp);
            cur = dir == 0 ? (((c1 << 2) & mask) | (shn_key_t)w2) : ((c1 >> 2) | ((shn_key_t)w2 << top));
            cur_aux = __shfl_sync(FULL, caux, g2 + w2);
            const uint32_t c2slot = (uint32_t)__shfl_sync(FULL, (uint32_t)cslot, g2 + w2);
            if (lane == 0 && len < path_cap) {
              const uint64_t e = dir == 0 ? (uint64_t)len : path_cap - 1 - len;
              my_path_slot[e] = c2slot;
              my_path_base[e] = (uint8_t)w2;
            }
            path_overflow |= len >= path_cap;
            ++len;
            tot += bw2;
            __syncwarp();
          }
          // claims still pending when this half ended, and halves abandoned after a lost claim: never
          // committed from this window, whatever the stamps on the path say
          poisoned |= __any_sync(FULL, (pend1 & kAuxStampMask) >= stamp || (pend2 & kAuxStampMask) >= stamp) != 0;
          pend1 = pend2 = 0;
          if (lane == 0) {
            if (dir == 0) {
              sh_len_r[ws] = len;
              sh_tot_r[ws] = tot;
            } else {
              sh_len_l[ws] = len;
              sh_tot_l[ws] = tot;
            }
          }
        }
      }
      const unsigned blo = __reduce_or_sync(FULL, (unsigned)bl);
      const unsigned bhi = __reduce_or_sync(FULL, (unsigned)(bl >> 32));
      if (lane == 0) {
        if (poisoned) atomicOr(&sh_poison[ws], 1u);
        const unsigned long long bb = (((unsigned long long)bhi << 32) | blo) & ~(1ull << ws);
        if (bb) atomicOr(&sh_block[ws], bb);
      }
    }
    __syncthreads();  // all claims of the window are in L2
    SHN_PHASE(1)

    // ---- 3. which paths are intact, and who holds the seeds? -----------------------------------
    // The whole CTA re-reads every path of the window (a long walk checked by one warp alone took
    // as long as the walk itself: measured 38 of 90 ms for the largest component).
    if (threadIdx.x < win_n) {
      const uint32_t ws = threadIdx.x, stamp_r = 2u * ((uint32_t)kSpecWindow - ws);
      // (both halves together longer than the buffer: they may have overwritten each other)
      if ((uint64_t)sh_len_r[ws] + sh_len_l[ws] > path_cap) atomicAdd(&a.counters[4], 1ull);
      sh_intact[ws] = (sh_len_r[ws] && !sh_poison[ws]) ? 1u : 0u;
      const uint32_t sst = __ldcg(&tv.slots[sh_win_slot[ws]].idx) & kAuxStampMask;
      sh_thief[ws] = (sst > stamp_r && sst <= 2u * (uint32_t)kSpecWindow)
                         ? (uint32_t)kSpecWindow - ((sst + 1u) >> 1)
                         : 0xFFFFFFFFu;
    }
    __syncthreads();
    SHN_PHASE(5)
    for (uint32_t ws = 0; ws < win_n; ++ws) {
      // (a half longer than the buffer is only stamped/cleared up to the buffer: the stage is rerun)
      const uint32_t len_r = (uint32_t)min((uint64_t)sh_len_r[ws], path_cap);
      const uint32_t len_l = (uint32_t)min((uint64_t)sh_len_l[ws], path_cap - len_r);
      const uint32_t stamp_r = 2u * ((uint32_t)kSpecWindow - ws);
      const uint32_t* ps = cta_path_slot + (uint64_t)ws * path_cap;
      bool mine = true;
#pragma unroll 4
      for (uint32_t e = threadIdx.x; e < len_r + len_l; e += kSpecWarps * 32) {
        const bool right = e < len_r;
        const uint64_t at = right ? (uint64_t)e : path_cap - 1 - (e - len_r);
        mine &= (__ldcg(&tv.slots[ps[at]].idx) & kAuxStampMask) == (right ? stamp_r : stamp_r - 1u);
      }
      if (!mine) sh_intact[ws] = 0u;   // benign race: every writer stores 0
    }
    __syncthreads();
    SHN_PHASE(2)
    if (threadIdx.x == 0) {
      // in pop order: COMMIT = intact and never blocked by anything but committed walks;
      // SKIP = the seed belongs to a committed earlier walk (the sequential loop finds it traversed);
      // anything else ends the prefix and is retried
      uint32_t P = 0;
      uint64_t off = sh_lp;
      unsigned long long committed = 0;
      for (; P < win_n; ++P) {
        uint32_t st = 0;
        if (sh_intact[P] && (sh_block[P] & ~committed) == 0) st = 1;
        else if (sh_thief[P] != 0xFFFFFFFFu && ((committed >> sh_thief[P]) & 1ull)) st = 2;
        if (st == 0) break;
        sh_status[P] = st;
        if (st == 1) {
          committed |= 1ull << P;
          sh_off[P] = off;
          off += (uint64_t)sh_len_r[P] + sh_len_l[P];
        }
      }
      sh_P = P;
      sh_lp = off;
      sh_cursor = P < win_n ? s_begin + sh_win_pos[P] : sh_cursor_after;
    }
    __syncthreads();
    SHN_PHASE(3)
    const uint32_t P = sh_P;
    if (warp == 0) {
      n_commit += P;
      n_retry += win_n - P;
    }

    // ---- 4. commit the intact prefix in order, roll the rest back ---------------------------
    for (uint32_t ws = warp; ws < win_n; ws += kSpecWarps) {
      const uint32_t len_r = (uint32_t)min((uint64_t)sh_len_r[ws], path_cap);
      const uint32_t len_l = (uint32_t)min((uint64_t)sh_len_l[ws], path_cap - len_r);
      const uint32_t len = len_r + len_l, stamp_r = 2u * ((uint32_t)kSpecWindow - ws);
      const uint32_t* ps = cta_path_slot + (uint64_t)ws * path_cap;
      const uint8_t* pb = cta_path_base + (uint64_t)ws * path_cap;
      if (ws < P && sh_status[ws] == 1) {
        // walk log: the seed, the right extension, then the left extension, each in walking order
        const uint64_t off = sh_off[ws];
        for (uint32_t e = lane; e < len; e += 32) {
          const uint64_t at = e < len_r ? (uint64_t)e : path_cap - 1 - (e - len_r);
          atomicOr(&tv.slots[ps[at]].idx, kAuxTraversed);
          if (off + e < le) a.walk_log[off + e] = pb[at];
        }
        overflow |= off + len > le;
        if (lane == 0 && len) {
          const uint32_t rank = a.ranks_by_comp[s_begin + sh_win_pos[ws]];
          a.started[rank] = 1;
          a.nr[rank] = sh_len_r[ws] - 1u;
          a.nl[rank] = sh_len_l[ws];
          a.totwt[rank] = sh_tot_r[ws] + sh_tot_l[ws];
          a.logstart[rank] = off;
        }
        traversed += len;
      } else {
        for (uint32_t e = lane; e < len; e += 32) {
          const bool right = e < len_r;
          const uint64_t at = right ? (uint64_t)e : path_cap - 1 - (e - len_r);
          const uint32_t stamp = right ? stamp_r : stamp_r - 1u;
          uint32_t* p = &tv.slots[ps[at]].idx;
          const uint32_t v = __ldcg(p);  // only the stamp bits change: the CAS fails iff stolen
          if ((v & kAuxStampMask) == stamp) atomicCAS(p, v, v & ~kAuxStampMask);
        }
      }
    }
    __syncthreads();
    SHN_PHASE(4)
  }
#undef SHN_PHASE
  if (ph_on)
    for (int k = 0; k < 6; ++k) sa.phase_ns[6 * (uint64_t)blockIdx.x + k] = ph[k];
  if (lane == 0) {
    atomicAdd(&a.counters[0], traversed);
    atomicMax(&a.counters[1], rounds);
    if (overflow) atomicAdd(&a.counters[2], 1ull);
    if (path_overflow) atomicAdd(&a.counters[4], 1ull);  // a walk longer than its path buffer: the host reruns serially
    if (warp == 0) atomicAdd(&a.counters[3], windows);
    if (warp == 0 && a.trace) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      a.trace[3 * (uint64_t)blockIdx.x] = tns;
      a.trace[3 * (uint64_t)blockIdx.x + 1] = (windows << 32) | (rounds & 0xFFFFFFFFull);
      a.trace[3 * (uint64_t)blockIdx.x + 2] = (n_commit << 32) | n_retry;
    }
  }
}

// ---- contig assembly from the walk log -----------------------------------------------------
// one thread per output base of the candidate contigs
__global__ void __launch_bounds__(kBlock)
    assemble_kernel(const ShnSlot* __restrict__ slots, const uint8_t* __restrict__ walk_log,
                    const uint32_t* __restrict__ cand_walk, const uint64_t* __restrict__ cand_off,
                    uint64_t n_cand, uint64_t total, const uint32_t* __restrict__ w_seed_slot,
                    const uint32_t* __restrict__ w_nl, const uint32_t* __restrict__ w_nr,
                    const uint64_t* __restrict__ w_logstart, int k1, uint8_t* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t lo = 0, hi = n_cand;
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&cand_off[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  uint32_t w = cand_walk[lo];
  uint64_t q = g - cand_off[lo];
  uint32_t nl = w_nl[w], nr = w_nr[w];
  uint64_t ls = w_logstart[w];
  uint8_t code;
  if (q < nl) {
    code = walk_log[ls + 1 + nr + (nl - 1 - q)];            // reversed(left_extension)
  } else if (q < (uint64_t)nl + k1) {
    shn_key_t key = slots[w_seed_slot[w]].key;                // start_kmer
    code = (uint8_t)((key >> (2 * (k1 - 1 - (int)(q - nl)))) & 3u);
  } else {
    code = walk_log[ls + 1 + (q - nl - k1)];                  // right_extension
  }
  out[g] = code;
}

// (key, owner, pos) entries of all length-L windows of the given contigs (2-bit codes)
__global__ void __launch_bounds__(kBlock)
    window_entries_kernel(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offs,
                          const uint64_t* __restrict__ ent_off, uint64_t n_contigs, uint64_t total,
                          int L, uint32_t owner_base, uint64_t* __restrict__ keys,
                          uint32_t* __restrict__ owner, uint32_t* __restrict__ pos) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t lo = 0, hi = n_contigs;
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  uint64_t start = offs[lo], end = offs[lo + 1];
  if (g + L > end) return;
  uint64_t x = 0;
  for (int j = 0; j < L; ++j) x = (x << 2) | (uint64_t)(codes[g + j] & 3u);
  uint64_t o = ent_off[lo] + (g - start);
  keys[o] = x;
  owner[o] = owner_base + (uint32_t)lo;
  pos[o] = (uint32_t)(g - start);
}

// K1-mer windows of the accepted contigs as table keys (SHN_KEY_WORDS words each), contig order
__global__ void __launch_bounds__(kBlock)
    k1mer_windows_kernel(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offs,
                         const uint64_t* __restrict__ ent_off, uint64_t n_contigs, uint64_t total,
                         int k1, uint64_t* __restrict__ keys) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t lo = 0, hi = n_contigs;
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  uint64_t start = offs[lo], end = offs[lo + 1];
  if (g + k1 > end) return;
  shn_key_t x = 0;
  for (int j = 0; j < k1; ++j) x = (x << 2) | (shn_key_t)(codes[g + j] & 3u);
  shn_store_key(keys, ent_off[lo] + (g - start), x);
}

__global__ void __launch_bounds__(kBlock)
    window_counts_kernel(const uint64_t* __restrict__ offs, uint64_t n, int L, uint64_t* __restrict__ cnt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    cnt[i] = 0;
    return;
  }
  uint64_t len = offs[i + 1] - offs[i];
  cnt[i] = len >= (uint64_t)L ? len - L + 1 : 0;
}

// ---- duplicate filter: frontier rounds over the pair table ---------------------------------
// status: 0 = unresolved, 1 = accepted, 2 = rejected.  Candidate j resolves once every partner
// d < j that shares an r-mer with it is resolved; the lowest unresolved candidate always is.
__global__ void __launch_bounds__(kBlock)
    dup_round_kernel(const uint64_t* __restrict__ seg_off, const uint32_t* __restrict__ lo,
                     const uint32_t* __restrict__ count, const uint32_t* __restrict__ max_i,
                     const uint32_t* __restrict__ covered, const uint64_t* __restrict__ cand_off,
                     uint64_t lo_cand, uint64_t hi_cand, const uint8_t* __restrict__ status_in,
                     uint8_t* __restrict__ status_out, uint8_t* __restrict__ dup_flag,
                     unsigned long long* counters) {
  uint64_t j = lo_cand + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= hi_cand) return;
  uint8_t st = status_in[j];
  if (st != 0) {
    status_out[j] = st;
    return;
  }
  // best = accepted partner with max (count, last hit position, id)  -- duplicate_check's
  // `>=` running maximum, extension_correction.py:251-259 (closed form in SURVEY 8a a6).
  // Partners are sorted by id ascending, so '>=' on (count,last) lets the larger id win ties.
  // An unresolved partner only matters if it could still become that maximum.
  uint32_t b_cnt = 0, b_last = 0, b_cov = 0;
  uint32_t u_cnt = 0, u_last = 0;
  bool have = false, have_u = false, u_after_best = false;
  for (uint64_t p = seg_off[j]; p < seg_off[j + 1]; ++p) {
    uint8_t sd = status_in[lo[p]];
    if (sd == 2) continue;
    uint32_t cnt = count[p], last = max_i[p];
    if (sd == 1) {
      if (!have || cnt > b_cnt || (cnt == b_cnt && last >= b_last)) {
        have = true;
        b_cnt = cnt;
        b_last = last;
        b_cov = covered[p];
        u_after_best = false;
      }
    } else {  // unresolved: remember the strongest one, and whether it comes after `best`
      if (!have_u || cnt > u_cnt || (cnt == u_cnt && last >= u_last)) {
        have_u = true;
        u_cnt = cnt;
        u_last = last;
      }
      if (have && cnt == b_cnt && last == b_last) u_after_best = true;  // larger id, equal key
    }
  }
  bool ready = !have_u;
  if (have_u && have) {
    // the strongest unresolved partner loses against `best` even if it gets accepted
    bool u_wins = u_cnt > b_cnt || (u_cnt == b_cnt && u_last > b_last) ||
                  (u_cnt == b_cnt && u_last == b_last && u_after_best);
    ready = !u_wins;
  }
  if (!ready) {
    status_out[j] = 0;
    atomicAdd(&counters[0], 1ull);
    return;
  }
  uint64_t len = cand_off[j + 1] - cand_off[j];
  bool dup = have && (2ull * b_cov > len);  // sum(a) > 0.5*len, :267
  dup_flag[j] = dup ? 1 : 0;
  status_out[j] = dup ? 2 : 1;
}

__global__ void __launch_bounds__(kBlock)
    seg_offsets_kernel(const uint32_t* __restrict__ hi, uint64_t n_pairs, uint64_t n_owner,
                       uint64_t* __restrict__ seg_off) {
  // seg_off[j] = first pair index with hi >= j, for j in 0..n_owner: one binary search per owner
  // (a loop over the gap behind every pair left one thread filling the whole tail: 1.4 ms per launch)
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_owner) return;
  uint64_t lo = 0, up = n_pairs;
  while (lo < up) {
    const uint64_t mid = (lo + up) >> 1;
    if ((uint64_t)hi[mid] < j)
      lo = mid + 1;
    else
      up = mid;
  }
  seg_off[j] = lo;
}

// ---- compaction of accepted candidates -----------------------------------------------------
__global__ void __launch_bounds__(kBlock)
    copy_contigs_kernel(const uint8_t* __restrict__ src, const uint64_t* __restrict__ src_off,
                        const uint32_t* __restrict__ acc_cand, const uint64_t* __restrict__ dst_off,
                        uint64_t n_acc, uint64_t total, uint8_t* __restrict__ dst) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t lo = 0, hi = n_acc;
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&dst_off[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  dst[g] = src[src_off[acc_cand[lo]] + (g - dst_off[lo])];
}

__global__ void __launch_bounds__(kBlock)
    allowed_weights_kernel(ShnTableView t, const uint64_t* __restrict__ keys, uint64_t n,
                           uint32_t* __restrict__ w_out, unsigned long long* counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w = 0;
  uint64_t s = table_find(t, shn_load_key(keys, i), &w);
  if (s == ~0ull) atomicAdd(&counters[0], 1ull);
  w_out[i] = w & SHN_WEIGHT_MASK;
}

// ---- contig components: min-label propagation over the (small) edge list --------------------
__global__ void __launch_bounds__(kBlock) label_init_kernel(uint32_t* label, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) label[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    label_hook_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n_edges,
                      uint32_t* label, unsigned long long* changed) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  uint32_t la = label[a[e]], lb = label[b[e]];
  if (la == lb) return;
  uint32_t lo = la < lb ? la : lb, hi = la < lb ? lb : la;
  atomicAdd(changed, 1ull);
  // hook the larger label's representative under the smaller label
  atomicMin(&label[hi], lo);
  atomicMin(&label[a[e]], lo);
  atomicMin(&label[b[e]], lo);
}

__global__ void __launch_bounds__(kBlock)
    label_jump_kernel(uint32_t* label, uint64_t n, unsigned long long* changed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t l = label[i];
  uint32_t ll = label[l];
  if (ll < l) {
    label[i] = ll;
    atomicAdd(changed, 1ull);
  }
}

template <typename T>
void d2h(shn_ctx* c, std::vector<T>& dst, const void* src, uint64_t n) {
  dst.resize(n);
  if (n) CUDA_CHECK(cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

template <typename T>
void h2d(shn_ctx* c, DevBuf& dst, const std::vector<T>& src) {
  dst.reserve(std::max<uint64_t>(src.size(), 1) * sizeof(T));
  if (!src.empty())
    CUDA_CHECK(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice,
                               c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

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

template <typename InT, typename OutT>
void exclusive_sum(shn_ctx* c, const InT* in, OutT* out, uint64_t n) {
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, c->stream));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, in, out, (int64_t)n, c->stream));
}

struct CastU64 {
  __device__ __forceinline__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};
void exclusive_sum_u32(shn_ctx* c, const uint32_t* in, uint64_t* out, uint64_t n) {
  cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> it(in, CastU64());
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, out, (int64_t)n, c->stream));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, it, out, (int64_t)n, c->stream));
}

// The length and hyperbola terms of extension_correction.py:353,361, evaluated on the host with
// the same double expression order and the same libm pow() CPython's math.pow calls.
bool passes_shape(uint64_t length, uint64_t tot_wt, uint64_t tot_kmer, uint32_t min_weight,
                  uint32_t min_length) {
  volatile double avg_wt = (double)tot_wt / (double)(tot_kmer > 1 ? tot_kmer : 1);
  volatile double lhs = (double)length * pow(avg_wt, 1 / 4.0);
  volatile double rhs = (double)(2ull * min_length) * pow((double)min_weight, 1 / 4.0);
  return length >= min_length && lhs >= rhs;
}

}  // namespace

static void l3_state_free(shn_ctx* c) {
  delete static_cast<L3State*>(c->l3);
  c->l3 = nullptr;
}

struct HostTrace {  // SHN_HOST_TRACE=1: wall-clock marks of l3_run's host side on stderr
  bool on = getenv("SHN_HOST_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[l3 host] %-28s +%.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// a3-a5: seeds, raw components, greedy walks, shape filter, candidate contigs (the part of the
// seed loop that only needs the K1-mer table; every K1-mer graph component is self-contained)
static bool l3_walks_impl(shn_ctx* c, uint32_t min_weight, uint32_t min_length, bool no_spec);
void l3_walks(shn_ctx* c, uint32_t min_weight, uint32_t min_length) {
  // second attempt without the speculative tiers if a walk outgrew its (capped) path buffer
  if (!l3_walks_impl(c, min_weight, min_length, false)) {
    const bool ok = l3_walks_impl(c, min_weight, min_length, true);
    SHN_CHECK(ok, "internal error: path overflow without speculative walks");
  }
}
static bool l3_walks_impl(shn_ctx* c, uint32_t min_weight, uint32_t min_length, bool no_spec) {
  HostTrace ht;
  SHN_CHECK(c->n_buckets > 0, "no K1-mer table built (call shn_table_build first)");
  shn_l3_free(c);
  L3State* s = new L3State();
  c->l3 = s;
  c->l3_free = &l3_state_free;
  s->min_weight = min_weight;
  s->min_length = min_length;
  const int k1 = c->k1;
  const uint64_t n_slots = c->n_buckets * SHN_BSLOTS;
  SHN_CHECK(n_slots < 0xFFFFFFFFull, "table too large for 32-bit slot indices");
  ShnTableView tv = table_view(c);
  cudaStream_t st = c->stream;
  unsigned long long h[8];
  const unsigned stream_grid =
      (unsigned)std::min<uint64_t>((n_slots + kBlock - 1) / kBlock, (uint64_t)c->sm_count * 32);

  // from here until the walks are done the idx words of the table hold the walks' aux words
  struct IdxParking {
    shn_ctx* c;
    ShnSlot* slots;
    uint64_t n_slots;
    unsigned grid;
    DevBuf saved;
    bool parked = false, pending = false;
    cudaEvent_t ev_done = nullptr;
    // The restore pass only writes the idx words, which nothing between the walks and the end of
    // l3_walks reads: it runs on a side stream under the compaction / assembly of the candidates.
    void unpark_async() {
      if (!parked) return;
      parked = false;
      if (!c->stream3) cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking);
      cudaEvent_t ev_go = c->prof_event();
      ev_done = c->prof_event();
      cudaEventRecord(ev_go, c->stream);
      cudaStreamWaitEvent(c->stream3, ev_go, 0);
      idx_unpark_kernel<<<grid, kBlock, 0, c->stream3>>>(slots, n_slots, saved.as<uint32_t>());
      cudaEventRecord(ev_done, c->stream3);
      c->prof_pool.push_back(ev_go);
      pending = true;
    }
    void finish() {
      if (!pending) return;
      pending = false;
      cudaStreamWaitEvent(c->stream, ev_done, 0);
      cudaEventSynchronize(ev_done);
      c->prof_pool.push_back(ev_done);
      saved.release();
    }
    void unpark() {
      unpark_async();
      finish();
    }
    ~IdxParking() { unpark(); }  // also on the error paths: the table must stay usable
  } parking{c, tv.slots, n_slots, stream_grid};
  parking.saved.reserve(n_slots * 4);

  // ---- a3: seeds in pop order -------------------------------------------------------------
  unsigned long long* ctr = zero_counters(c);
  {
    ProfScope ps(c, "seed_count");
    seed_count_kernel<<<stream_grid, kBlock, 0, st>>>(tv.slots, n_slots, min_weight, ctr);
    KERNEL_CHECK();
  }
  read_counters(c, h, 2);
  const uint64_t n_seeds = h[0];
  int wbits = 1;  // bits of the largest seed weight
  while (wbits < 30 && (h[1] >> wbits)) ++wbits;
  const uint32_t wmask = (1u << wbits) - 1u;
  // index field: input lines of this table, or (sharded tables) 34-bit global lines behind gline[]
  int ibits = 34;
  if (c->explicit_idx && !c->gline_dev) {
    ibits = 32;
  } else if (!c->gline_dev) {
    ibits = 1;
    while (ibits < 32 && (c->n_items >> ibits)) ++ibits;
  }
  s->sz.n_seeds = n_seeds;
  DevBuf seed_slot;  // by rank
  seed_slot.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  if (n_seeds) {
    DevBuf skey, sslot, skey2;
    skey.reserve(n_seeds * 8);
    skey2.reserve(n_seeds * 8);
    sslot.reserve(n_seeds * 4);
    ctr = zero_counters(c);
    {
      ProfScope ps(c, "seed_emit");
      seed_emit_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(
          tv.slots, n_slots, min_weight, c->gline_dev, skey.as<uint64_t>(), sslot.as<uint32_t>(), ctr,
          parking.saved.as<uint32_t>(), wmask, ibits);
      KERNEL_CHECK();
      parking.parked = true;
    }
    ProfScope ps(c, "seed_sort");
    const int key_bits = ibits + wbits;  // only the bits that vary are sorted
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, skey.as<uint64_t>(), skey2.as<uint64_t>(),
                                               sslot.as<uint32_t>(), seed_slot.as<uint32_t>(),
                                               (int64_t)n_seeds, 0, key_bits, st));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, skey.as<uint64_t>(),
                                               skey2.as<uint64_t>(), sslot.as<uint32_t>(),
                                               seed_slot.as<uint32_t>(), (int64_t)n_seeds, 0, key_bits, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }

  if (!parking.parked) {
    ProfScope ps(c, "idx_park");
    idx_park_kernel<<<stream_grid, kBlock, 0, st>>>(tv.slots, n_slots, parking.saved.as<uint32_t>());
    KERNEL_CHECK();
    parking.parked = true;
  }
  ht.mark("seeds sorted");
  // ---- raw components of the successor graph ------------------------------------------------
  DevBuf parent, root_flag, root_id;
  parent.reserve(n_slots * 4);
  root_flag.reserve((n_slots + 1) * 4);
  root_id.reserve((n_slots + 1) * 4);
  {
    ProfScope ps(c, "uf_init");
    uf_init_kernel<<<stream_grid, kBlock, 0, st>>>(parent.as<uint32_t>(), n_slots);
    KERNEL_CHECK();
  }
  {
    ProfScope ps(c, "uf_edges");
    uf_edges_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(tv, parent.as<uint32_t>(), n_slots, k1);
    KERNEL_CHECK();
  }
  {
    ProfScope ps(c, "uf_flatten");
    CUDA_CHECK(cudaMemsetAsync(root_flag.as<uint32_t>() + n_slots, 0, 4, st));
    uf_flatten_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(parent.as<uint32_t>(), n_slots,
                                                                   root_flag.as<uint32_t>());
    KERNEL_CHECK();
  }
  uint32_t n_comps = 0;
  {
    ProfScope ps(c, "comp_ids");
    exclusive_sum(c, root_flag.as<uint32_t>(), root_id.as<uint32_t>(), n_slots + 1);
    CUDA_CHECK(cudaMemcpyAsync(&n_comps, root_id.as<uint32_t>() + n_slots, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  root_flag.release();
  s->sz.n_raw_comps = n_comps;

  DevBuf comp_nodes, comp_seeds, log_off, seed_off;
  comp_nodes.reserve(((uint64_t)n_comps + 1) * 4);
  comp_seeds.reserve(((uint64_t)n_comps + 1) * 4);
  log_off.reserve(((uint64_t)n_comps + 1) * 8);
  seed_off.reserve(((uint64_t)n_comps + 1) * 8);
  CUDA_CHECK(cudaMemsetAsync(comp_nodes.p, 0, ((uint64_t)n_comps + 1) * 4, st));
  CUDA_CHECK(cudaMemsetAsync(comp_seeds.p, 0, ((uint64_t)n_comps + 1) * 4, st));
  {
    ProfScope ps(c, "comp_count");
    const char* envh = getenv("SHN_COMP_HIST_MAX_BYTES");  // tests: 0 forces the global-atomics variant
    const uint64_t hist_max = envh ? strtoull(envh, nullptr, 10) : 40ull * 1024;
    if (n_comps && (uint64_t)n_comps * 4 <= hist_max) {
      comp_count_smem_kernel<<<c->sm_count * 8, kBlock, (size_t)n_comps * 4, st>>>(
          parent.as<uint32_t>(), root_id.as<uint32_t>(), n_slots, n_comps, comp_nodes.as<uint32_t>());
    } else {
      comp_count_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(
          parent.as<uint32_t>(), root_id.as<uint32_t>(), n_slots, comp_nodes.as<uint32_t>());
    }
    KERNEL_CHECK();
  }
  DevBuf seed_comp, rank_in, seed_comp_s, ranks_by_comp;
  seed_comp.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  rank_in.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  seed_comp_s.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  ranks_by_comp.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  if (n_seeds) {
    ProfScope ps(c, "seed_group", 3);
    seed_comp_kernel<<<shn_grid(n_seeds, kBlock), kBlock, 0, st>>>(
        seed_slot.as<uint32_t>(), parent.as<uint32_t>(), root_id.as<uint32_t>(), n_seeds,
        seed_comp.as<uint32_t>(), rank_in.as<uint32_t>());
    KERNEL_CHECK();
    int bits = 1;
    while (bits < 32 && (n_comps >> bits)) ++bits;
    size_t tb = 0;  // stable: ranks stay ascending inside a component
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, seed_comp.as<uint32_t>(),
                                               seed_comp_s.as<uint32_t>(), rank_in.as<uint32_t>(),
                                               ranks_by_comp.as<uint32_t>(), (int64_t)n_seeds, 0, bits, st));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, seed_comp.as<uint32_t>(),
                                               seed_comp_s.as<uint32_t>(), rank_in.as<uint32_t>(),
                                               ranks_by_comp.as<uint32_t>(), (int64_t)n_seeds, 0, bits, st));
  }
  // table slots of the seeds, grouped like ranks_by_comp
  DevBuf slots_by_comp;
  slots_by_comp.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  if (n_seeds) {
    ProfScope ps(c, "seed_group");
    gather_u32_kernel<<<shn_grid(n_seeds, kBlock), kBlock, 0, st>>>(
        seed_slot.as<uint32_t>(), ranks_by_comp.as<uint32_t>(), n_seeds, slots_by_comp.as<uint32_t>());
    KERNEL_CHECK();
  }
  parent.release();
  root_id.release();
  exclusive_sum_u32(c, comp_nodes.as<uint32_t>(), log_off.as<uint64_t>(), (uint64_t)n_comps + 1);
  if (n_seeds) {
    seed_offsets_kernel<<<shn_grid((uint64_t)n_comps + 1, kBlock), kBlock, 0, st>>>(
        seed_comp_s.as<uint32_t>(), n_seeds, n_comps, seed_off.as<uint64_t>(), comp_seeds.as<uint32_t>());
    KERNEL_CHECK();
  } else {
    CUDA_CHECK(cudaMemsetAsync(seed_off.p, 0, ((uint64_t)n_comps + 1) * 8, st));
  }
  // components that own at least one seed, in descending node count: similar-sized components
  // share a warp and the big ones start first
  DevBuf comp_order;
  comp_order.reserve(std::max<uint32_t>(n_comps, 1) * 4);
  uint32_t n_active = 0;
  DevBuf work_s;  // node counts of the active components, descending (same order as comp_order)
  if (n_comps) {
    DevBuf ids, work;
    ids.reserve((uint64_t)n_comps * 4);
    work.reserve((uint64_t)n_comps * 4);
    work_s.reserve((uint64_t)n_comps * 4);
    ctr = zero_counters(c);
    ProfScope ps(c, "comp_order", 2);
    comp_work_kernel<<<shn_grid(n_comps, kBlock), kBlock, 0, st>>>(
        comp_nodes.as<uint32_t>(), comp_seeds.as<uint32_t>(), n_comps, work.as<uint32_t>(),
        ids.as<uint32_t>(), ctr);
    KERNEL_CHECK();
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(
        nullptr, tb, work.as<uint32_t>(), work_s.as<uint32_t>(), ids.as<uint32_t>(),
        comp_order.as<uint32_t>(), (int64_t)n_comps, 0, 32, st));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(
        c->tmp(tb), tb, work.as<uint32_t>(), work_s.as<uint32_t>(), ids.as<uint32_t>(),
        comp_order.as<uint32_t>(), (int64_t)n_comps, 0, 32, st));
    read_counters(c, h, 1);
    n_active = (uint32_t)h[0];
  }

  ht.mark("components + seed groups");
  // ---- a4: greedy walks -----------------------------------------------------------------------
  const uint64_t n_nodes = c->n_distinct;
  s->walk_log.reserve(std::max<uint64_t>(n_nodes, 1));
  DevBuf started, nl_r, nr_r, tot_r, ls_r;
  started.reserve(std::max<uint64_t>(n_seeds, 1));
  nl_r.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  nr_r.reserve(std::max<uint64_t>(n_seeds, 1) * 4);
  tot_r.reserve(std::max<uint64_t>(n_seeds, 1) * 8);
  ls_r.reserve(std::max<uint64_t>(n_seeds, 1) * 8);
  CUDA_CHECK(cudaMemsetAsync(started.p, 0, std::max<uint64_t>(n_seeds, 1), st));
  ctr = zero_counters(c);
  if (n_active) {
    WalkArgs a;
    a.tv = tv;
    {
      const char* envl = getenv("SHN_NO_LOOKAHEAD");
      a.lookahead = envl ? 0u : 1u;
    }
    a.k1 = k1;
    a.n_comps = n_active;
    a.comp_order = comp_order.as<uint32_t>();
    a.seed_off = seed_off.as<uint64_t>();
    a.ranks_by_comp = ranks_by_comp.as<uint32_t>();
    a.slots_by_comp = slots_by_comp.as<uint32_t>();
    a.log_off = log_off.as<uint64_t>();
    a.walk_log = s->walk_log.as<uint8_t>();
    a.started = started.as<uint8_t>();
    a.nl = nl_r.as<uint32_t>();
    a.nr = nr_r.as<uint32_t>();
    a.totwt = tot_r.as<uint64_t>();
    a.logstart = ls_r.as<uint64_t>();
    a.counters = ctr;
    a.trace = nullptr;
    DevBuf trace;
    const bool want_trace = getenv("SHN_WALK_TRACE") != nullptr;
    if (want_trace) {
      trace.reserve((uint64_t)n_active * 24);
      a.trace = trace.as<unsigned long long>();
    }
    // large components: speculative windows (one CTA each); the rest: one warp each, on a second
    // stream so that both kernels share the GPU
    uint32_t n_spec = 0;
    std::vector<uint64_t> h_path_off(1, 0);
    {
      const uint32_t peek = std::min<uint32_t>(n_active, 1u << 16);
      std::vector<uint32_t> top;
      d2h(c, top, work_s.p, peek);
      const char* env = getenv("SHN_SPEC_MIN_NODES");
      const uint64_t min_nodes = env ? strtoull(env, nullptr, 10) : 60000ull;
      const char* envb = getenv("SHN_SPEC_SCRATCH_GB");
      const uint64_t budget = (envb ? strtoull(envb, nullptr, 10) : 32ull) << 30;  // scratch for the paths
      const char* envc = getenv("SHN_SPEC_MAX_COMPS");
      const uint32_t max_spec = envc ? (uint32_t)strtoul(envc, nullptr, 10) : 4u * (uint32_t)c->sm_count;
      // path buffers: one per window position, as long as the longest walk can be.  A walk can in
      // principle visit its whole component; buffers are capped (SHN_SPEC_PATH_CAP entries) and the
      // rare walk that does not fit makes the whole stage fall back to the serial kernel (below).
      const char* envp = getenv("SHN_SPEC_PATH_CAP");
      const uint64_t path_cap = no_spec ? 0 : (envp ? strtoull(envp, nullptr, 10) : (4ull << 20));
      while (path_cap && n_spec < peek && n_spec < max_spec && top[n_spec] >= min_nodes) {
        const uint64_t cap = std::min<uint64_t>(top[n_spec], path_cap);
        if ((h_path_off.back() + (uint64_t)kSpecWindow * cap) * 5 > budget) break;
        h_path_off.push_back(h_path_off.back() + (uint64_t)kSpecWindow * cap);
        ++n_spec;
      }
    }
    DevBuf path_slot, path_base, path_off, phase_ns;
    cudaEvent_t ev_fork = c->prof_event(), ev_join = c->prof_event(), ev_join3 = c->prof_event(),
                ev_join4 = c->prof_event();
    {
      ProfScope ps(c, "walk", (n_spec ? 2 : 0) + (n_spec < n_active ? 1 : 0));  // upper bound: two tiers
      CUDA_CHECK(cudaEventRecord(ev_fork, st));
      if (n_spec < n_active) {
        if (!c->stream3) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamWaitEvent(c->stream3, ev_fork, 0));
      }
      bool serial_launched = false;
      auto launch_serial = [&]() {
        if (serial_launched || n_spec >= n_active) return;
        serial_launched = true;
        WalkArgs b = a;
        b.comp_order = a.comp_order + n_spec;
        b.n_comps = n_active - n_spec;
        if (b.trace) b.trace += 3 * (uint64_t)n_spec;
        const unsigned grid = shn_grid((uint64_t)b.n_comps * 32, kWalkBlock);
        if (b.trace)
          walk_kernel<0, true><<<grid, kWalkBlock, 0, c->stream3>>>(b);
        else if (k1 == kWalkK1)
          walk_kernel<kWalkK1, false><<<grid, kWalkBlock, 0, c->stream3>>>(b);
        else
          walk_kernel<0, false><<<grid, kWalkBlock, 0, c->stream3>>>(b);
        KERNEL_CHECK();
        CUDA_CHECK(cudaEventRecord(ev_join, c->stream3));
        CUDA_CHECK(cudaStreamWaitEvent(st, ev_join, 0));
      };
      if (n_spec) {
        SpecArgs sa;
        sa.w = a;
        sa.w.n_comps = n_spec;
        path_slot.reserve(h_path_off.back() * 4);
        path_base.reserve(h_path_off.back());
        h2d(c, path_off, h_path_off);
        sa.path_slot = path_slot.as<uint32_t>();
        sa.path_base = path_base.as<uint8_t>();
        sa.path_off = path_off.as<uint64_t>();
        sa.phase_ns = nullptr;
        if (want_trace) {
          phase_ns.reserve((uint64_t)n_spec * 6 * 8);
          CUDA_CHECK(cudaMemsetAsync(phase_ns.p, 0, (uint64_t)n_spec * 6 * 8, st));
          sa.phase_ns = phase_ns.as<unsigned long long>();
        }
        const char* envt = getenv("SHN_SPEC_TIER16");
        const uint32_t n16 =
            std::min<uint32_t>(n_spec, envt ? (uint32_t)strtoul(envt, nullptr, 10) : (uint32_t)c->sm_count / 4u);
        if (n16) {
          SpecArgs t = sa;
          t.w.n_comps = n16;
          if (k1 == kWalkK1)
            walk_spec_kernel<16, kWalkK1><<<n16, 16 * 32, 0, st>>>(t);
          else
            walk_spec_kernel<16, 0><<<n16, 16 * 32, 0, st>>>(t);
          KERNEL_CHECK();
        }
        // Launch order = block scheduling order: the 16-warp CTAs and the one-warp components hold
        // the long serial chains and must start at once; the short 8-warp CTAs fill what is left.
        if (!getenv("SHN_WALK_SERIAL_LAST")) launch_serial();
        // further tiers on their own streams: all kernels share the GPU.  SHN_SPEC_TIER8 = how many
        // components after the 16-warp tier get 8-warp CTAs (default: all); the rest get 4 warps.
        const char* envt8 = getenv("SHN_SPEC_TIER8");
        const uint32_t n8 = std::min<uint32_t>(n_spec - n16, envt8 ? (uint32_t)strtoul(envt8, nullptr, 10) : 0xFFFFFFFFu);
        const uint32_t n4 = n_spec - n16 - n8;
        for (int tier = 0; tier < 2; ++tier) {
          const uint32_t first = tier == 0 ? n16 : n16 + n8, count = tier == 0 ? n8 : n4;
          if (!count) continue;
          cudaStream_t& ts = tier == 0 ? c->stream4 : c->stream5;
          if (!ts) CUDA_CHECK(cudaStreamCreateWithFlags(&ts, cudaStreamNonBlocking));
          CUDA_CHECK(cudaStreamWaitEvent(ts, ev_fork, 0));
          SpecArgs t = sa;
          t.w.comp_order = a.comp_order + first;
          t.w.n_comps = count;
          t.path_off = sa.path_off + first;
          if (t.phase_ns) t.phase_ns += 6 * (uint64_t)first;
          if (t.w.trace) t.w.trace += 3 * (uint64_t)first;
          if (tier == 0) {
            if (k1 == kWalkK1)
              walk_spec_kernel<8, kWalkK1><<<count, 8 * 32, 0, ts>>>(t);
            else
              walk_spec_kernel<8, 0><<<count, 8 * 32, 0, ts>>>(t);
          } else {
            if (k1 == kWalkK1)
              walk_spec_kernel<4, kWalkK1><<<count, 4 * 32, 0, ts>>>(t);
            else
              walk_spec_kernel<4, 0><<<count, 4 * 32, 0, ts>>>(t);
          }
          KERNEL_CHECK();
          cudaEvent_t ev = tier == 0 ? ev_join3 : ev_join4;
          CUDA_CHECK(cudaEventRecord(ev, ts));
          CUDA_CHECK(cudaStreamWaitEvent(st, ev, 0));
        }
      }
      launch_serial();
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    c->prof_pool.push_back(ev_fork);
    c->prof_pool.push_back(ev_join);
    c->prof_pool.push_back(ev_join3);
    c->prof_pool.push_back(ev_join4);
    s->sz.n_spec_comps = n_spec;
    if (want_trace) {  // when does each component's warp finish? (tail = critical path)
      std::vector<unsigned long long> tr;
      d2h(c, tr, trace.p, (uint64_t)n_active * 3);
      unsigned long long t_min = ~0ull, t_max = 0, r_sum = 0;
      for (uint32_t w = 0; w < n_spec; ++w) t_min = std::min(t_min, tr[3 * w]);
      for (uint32_t w = n_spec; w < n_active; ++w) {
        t_min = std::min(t_min, tr[3 * w]);
        t_max = std::max(t_max, tr[3 * w]);
        r_sum += tr[3 * w + 1];
      }
      std::vector<unsigned long long> ends;
      for (uint32_t w = n_spec; w < n_active; ++w) ends.push_back(tr[3 * w] - t_min);
      if (ends.empty()) ends.push_back(0);
      std::sort(ends.begin(), ends.end());
      auto pct = [&](double p) { return ends[(size_t)(p * (ends.size() - 1))] / 1e6; };
      fprintf(stderr,
              "[walk trace] warps=%u total_rounds=%llu  finish-time spread (ms after the first warp "
              "finished): p50=%.1f p90=%.1f p99=%.1f max=%.1f\n",
              n_active, r_sum, pct(0.5), pct(0.9), pct(0.99), pct(1.0));
      {
        std::vector<uint32_t> top;
        d2h(c, top, work_s.p, std::min<uint32_t>(n_active, 4096));
        for (uint32_t w = 0; w < n_spec; ++w)
          if (w < 24 || w % 16 == 0 || w + 1 == n_spec)
            fprintf(stderr,
                    "[walk trace] spec comp #%u: nodes=%u windows=%llu rounds(warp0)=%llu committed=%llu "
                    "retried=%llu end=+%.1f ms\n",
                    w, top[w], tr[3 * w + 1] >> 32, tr[3 * w + 1] & 0xFFFFFFFFull, tr[3 * w + 2] >> 32,
                    tr[3 * w + 2] & 0xFFFFFFFFull, (tr[3 * w] - t_min) / 1e6);
      }
      if (n_spec) {
        std::vector<unsigned long long> phs;
        d2h(c, phs, phase_ns.p, (uint64_t)n_spec * 6);
        for (uint32_t w = 0; w < std::min<uint32_t>(n_spec, 4); ++w)
          fprintf(stderr,
                  "[walk trace] spec comp #%u phases (ms of thread 0): collect=%.1f walk=%.1f seed-check=%.1f "
                  "path-check=%.1f resolve=%.1f commit=%.1f\n",
                  w, phs[6 * w] / 1e6, phs[6 * w + 1] / 1e6, phs[6 * w + 5] / 1e6, phs[6 * w + 2] / 1e6,
                  phs[6 * w + 3] / 1e6, phs[6 * w + 4] / 1e6);
        const uint32_t w8 = std::min<uint32_t>(n_spec - 1, (uint32_t)c->sm_count / 4u);
        fprintf(stderr,
                "[walk trace] spec comp #%u (first 8-warp CTA) phases: collect=%.1f walk=%.1f seed-check=%.1f "
                "path-check=%.1f resolve=%.1f commit=%.1f\n",
                w8, phs[6 * w8] / 1e6, phs[6 * w8 + 1] / 1e6, phs[6 * w8 + 5] / 1e6, phs[6 * w8 + 2] / 1e6,
                phs[6 * w8 + 3] / 1e6, phs[6 * w8 + 4] / 1e6);
      }
      for (uint32_t w = n_spec; w < std::min<uint32_t>(n_active, n_spec + 6); ++w)
        fprintf(stderr,
                "[walk trace] largest one-warp component #%u: rounds=%llu cycles=%llu (%.0f per round) of which "
                "waiting for the home buckets=%llu (%.0f per round) end=+%.1f ms\n",
                w, tr[3 * w + 1], tr[3 * w + 2] >> 32, (double)(tr[3 * w + 2] >> 32) / std::max(1ull, tr[3 * w + 1]),
                (tr[3 * w + 2] & 0xFFFFFFFFull) << 8,
                (double)((tr[3 * w + 2] & 0xFFFFFFFFull) << 8) / std::max(1ull, tr[3 * w + 1]),
                (tr[3 * w] - t_min) / 1e6);
    }
  }
  {
    c->launches += 1;
    parking.unpark_async();
    CUDA_CHECK(cudaGetLastError());
  }
  read_counters(c, h, 5);
#ifdef SHN_WALK_STATS
  {
    unsigned long long it = 0, ca = 0;
    cudaMemcpyFromSymbol(&it, g_chase_iters, 8);
    cudaMemcpyFromSymbol(&ca, g_chase_calls, 8);
    fprintf(stderr, "[walk stats] chase calls=%llu iterations=%llu (cumulative) rounds(max warp)=%llu traversed=%llu\n",
            ca, it, h[1], h[0]);
  }
#endif
  if (h[4] != 0) {  // the idx words are restored; nothing else was changed
    parking.finish();
    return false;
  }
  SHN_CHECK(h[2] == 0, "internal error: walk log overflow (component node count mismatch)");
  s->sz.n_traversed = h[0];
  s->sz.walk_rounds = h[1];
  s->sz.spec_windows = h[3];

  ht.mark("walks done");
  // started walks in pop order (device compaction), then a5: the walks that pass the length +
  // hyperbola filter, in pop order ("candidates")
  uint64_t n_walks = 0, n_cand = 0;
  std::vector<uint32_t>& cand_walk = s->h_cand_walk;
  cand_walk.clear();
  std::vector<uint64_t>& cand_off = s->h_cand_off;
  cand_off.assign(1, 0);
  DevBuf& d_cand_walk = s->cand_walk;
  DevBuf& d_cand_off = s->cand_off;
  DevBuf& cand_codes = s->cand_codes;
  if (n_seeds) {
    DevBuf sel, nsel;
    sel.reserve(n_seeds * 4);
    nsel.reserve(8);
    ProfScope ps(c, "walk_compact", 8);
    cub::CountingInputIterator<uint32_t> it(0);
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, tb, it, started.as<uint8_t>(), sel.as<uint32_t>(),
                                          nsel.as<uint64_t>(), (int64_t)n_seeds, st));
    CUDA_CHECK(cub::DeviceSelect::Flagged(c->tmp(tb), tb, it, started.as<uint8_t>(), sel.as<uint32_t>(),
                                          nsel.as<uint64_t>(), (int64_t)n_seeds, st));
    CUDA_CHECK(cudaMemcpyAsync(&n_walks, nsel.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const uint64_t nw1 = std::max<uint64_t>(n_walks, 1);
    s->w_seed_slot.reserve(nw1 * 4);
    s->w_nl.reserve(nw1 * 4);
    s->w_nr.reserve(nw1 * 4);
    s->w_totwt.reserve(nw1 * 8);
    s->w_logstart.reserve(nw1 * 8);
    ht.mark("select started");
    if (n_walks) {
      DevBuf is_long, shape;
      is_long.reserve(n_walks);
      shape.reserve(n_walks);
      d_cand_walk.reserve(n_walks * 4);
      gather_walks_kernel<<<shn_grid(n_walks, kBlock), kBlock, 0, st>>>(
          sel.as<uint32_t>(), n_walks, seed_slot.as<uint32_t>(), nl_r.as<uint32_t>(),
          nr_r.as<uint32_t>(), tot_r.as<uint64_t>(), ls_r.as<uint64_t>(), k1, min_length,
          s->w_seed_slot.as<uint32_t>(), s->w_nl.as<uint32_t>(), s->w_nr.as<uint32_t>(),
          s->w_totwt.as<uint64_t>(), s->w_logstart.as<uint64_t>(), is_long.as<uint8_t>());
      KERNEL_CHECK();
      ht.mark("gather_walks launched");
      const char* envs = getenv("SHN_SHAPE_TOL");  // tests: a huge tolerance sends every walk to the host
      const double shape_tol = envs ? strtod(envs, nullptr) : 1e-9;
      DevBuf borderline;
      borderline.reserve(n_walks * 4);
      ctr = zero_counters(c);
      shape_kernel<<<shn_grid(n_walks, kBlock), kBlock, 0, st>>>(
          s->w_nl.as<uint32_t>(), s->w_nr.as<uint32_t>(), s->w_totwt.as<uint64_t>(), n_walks, k1,
          min_weight, min_length, shape_tol, shape.as<uint8_t>(), borderline.as<uint32_t>(), ctr);
      KERNEL_CHECK();
      read_counters(c, h, 1);
      if (h[0]) {
        // borderline walks (a handful): decided on the host with the reference's expression and
        // libm's pow, then patched into the flags
        const uint64_t nb = h[0];
        DevBuf b_nl, b_nr, b_tot, b_val;
        b_nl.reserve(nb * 4);
        b_nr.reserve(nb * 4);
        b_tot.reserve(nb * 8);
        b_val.reserve(nb);
        gather_u32_kernel<<<shn_grid(nb, kBlock), kBlock, 0, st>>>(s->w_nl.as<uint32_t>(),
                                                                  borderline.as<uint32_t>(), nb, b_nl.as<uint32_t>());
        gather_u32_kernel<<<shn_grid(nb, kBlock), kBlock, 0, st>>>(s->w_nr.as<uint32_t>(),
                                                                  borderline.as<uint32_t>(), nb, b_nr.as<uint32_t>());
        gather_u64_kernel<<<shn_grid(nb, kBlock), kBlock, 0, st>>>(s->w_totwt.as<uint64_t>(),
                                                                  borderline.as<uint32_t>(), nb, b_tot.as<uint64_t>());
        KERNEL_CHECK();
        std::vector<uint32_t> h_nl, h_nr;
        std::vector<uint64_t> h_tot;
        d2h(c, h_nl, b_nl.p, nb);
        d2h(c, h_nr, b_nr.p, nb);
        d2h(c, h_tot, b_tot.p, nb);
        std::vector<uint8_t> h_val(nb);
        for (uint64_t i = 0; i < nb; ++i) {
          const uint64_t tot_kmer = (uint64_t)h_nl[i] + h_nr[i] + 1;
          h_val[i] = passes_shape(tot_kmer + k1 - 1, h_tot[i], tot_kmer, min_weight, min_length) ? 1 : 0;
        }
        CUDA_CHECK(cudaMemcpyAsync(b_val.p, h_val.data(), nb, cudaMemcpyHostToDevice, st));
        patch_flags_kernel<<<shn_grid(nb, kBlock), kBlock, 0, st>>>(borderline.as<uint32_t>(),
                                                                   b_val.as<uint8_t>(), nb, shape.as<uint8_t>());
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(st));  // h_val goes out of scope
      }
      ht.mark("shape filter");
      tb = 0;
      CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, tb, it, shape.as<uint8_t>(), d_cand_walk.as<uint32_t>(),
                                            nsel.as<uint64_t>(), (int64_t)n_walks, st));
      CUDA_CHECK(cub::DeviceSelect::Flagged(c->tmp(tb), tb, it, shape.as<uint8_t>(),
                                            d_cand_walk.as<uint32_t>(), nsel.as<uint64_t>(),
                                            (int64_t)n_walks, st));
      CUDA_CHECK(cudaMemcpyAsync(&n_cand, nsel.p, 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      ht.mark("select candidates");
      DevBuf cand_len;
      cand_len.reserve((n_cand + 1) * 8);
      d_cand_off.reserve((n_cand + 1) * 8);
      cand_len_kernel<<<shn_grid(n_cand + 1, kBlock), kBlock, 0, st>>>(
          d_cand_walk.as<uint32_t>(), n_cand, s->w_nl.as<uint32_t>(), s->w_nr.as<uint32_t>(), k1,
          cand_len.as<uint64_t>());
      KERNEL_CHECK();
      exclusive_sum(c, cand_len.as<uint64_t>(), d_cand_off.as<uint64_t>(), n_cand + 1);
      d2h(c, cand_walk, d_cand_walk.p, n_cand);
      d2h(c, cand_off, d_cand_off.p, n_cand + 1);
    }
  }
  ht.mark("cand offsets + d2h");
  if (d_cand_walk.p == nullptr) d_cand_walk.reserve(4);
  if (d_cand_off.p == nullptr) {
    d_cand_off.reserve(8);
    CUDA_CHECK(cudaMemsetAsync(d_cand_off.p, 0, 8, st));
  }
  s->sz.n_walks = n_walks;
  const uint64_t cand_bases = cand_off.back();
  s->sz.n_candidates = n_cand;

  cand_codes.reserve(std::max<uint64_t>(cand_bases, 1));
  ht.mark("before assemble");
  if (cand_bases) {
    ProfScope ps(c, "assemble");
    assemble_kernel<<<shn_grid(cand_bases, kBlock), kBlock, 0, st>>>(
        tv.slots, s->walk_log.as<uint8_t>(), d_cand_walk.as<uint32_t>(), d_cand_off.as<uint64_t>(),
        n_cand, cand_bases, s->w_seed_slot.as<uint32_t>(), s->w_nl.as<uint32_t>(),
        s->w_nr.as<uint32_t>(), s->w_logstart.as<uint64_t>(), k1, cand_codes.as<uint8_t>());
    KERNEL_CHECK();
  }

  s->n_cand = n_cand;
  parking.finish();
  CUDA_CHECK(cudaStreamSynchronize(st));
  CUDA_CHECK(cudaGetLastError());
  ht.mark("assembled candidates");
  return true;
}

namespace {
L3State* need_l3(shn_ctx* c) {
  SHN_CHECK(c->l3 != nullptr, "shn_l3_run has not been called on this context");
  return static_cast<L3State*>(c->l3);
}
}  // namespace

// a6-a9: duplicate filter, allowed set, contig C-mer graph, contig components -- on the candidates
// l3_walks left behind (ext == 0), or on a candidate list supplied by the caller (ext != 0: the
// sharded path merges the candidates of all ranks in global pop order; d_codes / d_offs are device
// pointers).  allow_missing: allowed K1-mers that are not in THIS context's table get weight 0
// instead of an error (their owner rank knows the weight).
void l3_filter(shn_ctx* c, const uint8_t* ext_codes, const uint64_t* ext_offs, uint64_t ext_n, int ext,
               int allow_missing) {
  HostTrace ht;
  L3State* s = need_l3(c);
  const int k1 = c->k1;
  ShnTableView tv = table_view(c);
  cudaStream_t st = c->stream;
  unsigned long long h[8];
  unsigned long long* ctr = nullptr;
  const uint32_t min_weight = s->min_weight;
  (void)min_weight;
  if (ext) {
    s->foreign = true;
    s->n_cand = ext_n;
    s->cand_off.reserve((ext_n + 1) * 8);
    if (ext_n) {
      CUDA_CHECK(cudaMemcpyAsync(s->cand_off.p, ext_offs, (ext_n + 1) * 8, cudaMemcpyDeviceToDevice, st));
      d2h(c, s->h_cand_off, s->cand_off.p, ext_n + 1);
    } else {
      CUDA_CHECK(cudaMemsetAsync(s->cand_off.p, 0, 8, st));
      s->h_cand_off.assign(1, 0);
    }
    const uint64_t nb = s->h_cand_off.back();
    s->cand_codes.reserve(std::max<uint64_t>(nb, 1));
    if (nb) CUDA_CHECK(cudaMemcpyAsync(s->cand_codes.p, ext_codes, nb, cudaMemcpyDeviceToDevice, st));
  }
  const uint64_t n_cand = s->n_cand;
  std::vector<uint64_t>& cand_off = s->h_cand_off;
  const uint64_t cand_bases = cand_off.back();
  DevBuf& d_cand_off = s->cand_off;
  DevBuf& cand_codes = s->cand_codes;
  s->sz.n_candidates = n_cand;
  // ---- a6: duplicate filter ------------------------------------------------------------------
  std::vector<uint8_t> h_status(n_cand, 1);
  std::vector<uint8_t> h_dup(n_cand, 0);
  const int R = 15;  // extension_correction.py:357
  if (n_cand) {
    DevBuf ent_cnt, ent_off;
    ent_cnt.reserve((n_cand + 1) * 8);
    ent_off.reserve((n_cand + 1) * 8);
    window_counts_kernel<<<shn_grid(n_cand + 1, kBlock), kBlock, 0, st>>>(d_cand_off.as<uint64_t>(),
                                                                          n_cand, R, ent_cnt.as<uint64_t>());
    KERNEL_CHECK();
    exclusive_sum(c, ent_cnt.as<uint64_t>(), ent_off.as<uint64_t>(), n_cand + 1);
    uint64_t n_ent = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_ent, ent_off.as<uint64_t>() + n_cand, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    SelfJoin sj;
    if (n_ent) {
      DevBuf keys, owner, pos;
      keys.reserve(n_ent * 8);
      owner.reserve(n_ent * 4);
      pos.reserve(n_ent * 4);
      {
        ProfScope ps(c, "rmer_entries");
        window_entries_kernel<<<shn_grid(cand_bases, kBlock), kBlock, 0, st>>>(
            cand_codes.as<uint8_t>(), d_cand_off.as<uint64_t>(), ent_off.as<uint64_t>(), n_cand,
            cand_bases, R, 0u, keys.as<uint64_t>(), owner.as<uint32_t>(), pos.as<uint32_t>());
        KERNEL_CHECK();
      }
      sj.prepare(c, "rmer", keys, owner, pos, ent_off.as<uint64_t>(), n_cand, 0u, n_ent, 2 * R, R);
    }
    // Candidates are resolved in blocks of ascending rank: when a block is joined, every earlier
    // candidate is already accepted or rejected, and rejected ones (the bulk: near-duplicates of a
    // few accepted contigs) never produce match events again.
    DevBuf seg_off, st_a, st_b, dupf;
    seg_off.reserve((n_cand + 2) * 8);
    st_a.reserve(n_cand);
    st_b.reserve(n_cand);
    dupf.reserve(n_cand);
    CUDA_CHECK(cudaMemsetAsync(st_a.p, 0, n_cand, st));
    CUDA_CHECK(cudaMemsetAsync(dupf.p, 0, n_cand, st));
    const char* envb = getenv("SHN_DUP_BLOCKS");
    const uint64_t n_blocks = std::max<uint64_t>(1, envb ? strtoull(envb, nullptr, 10) : 4);
    const uint64_t block = std::max<uint64_t>(4096, (n_cand + n_blocks - 1) / n_blocks);
    uint64_t rounds = 0;
    PairTable pt;
    for (uint64_t lo = 0; lo < n_cand; lo += block) {
      const uint64_t hi = std::min(n_cand, lo + block);
      sj.join((uint32_t)lo, (uint32_t)hi, st_a.as<uint8_t>(), &pt);
      seg_offsets_kernel<<<shn_grid(n_cand + 1, kBlock), kBlock, 0, st>>>(pt.hi.as<uint32_t>(), pt.n,
                                                                          n_cand, seg_off.as<uint64_t>());
      KERNEL_CHECK();
      CUDA_CHECK(cudaMemcpyAsync(st_b.p, st_a.p, n_cand, cudaMemcpyDeviceToDevice, st));
      // kDupBatch rounds per host round trip; rounds after convergence only copy the statuses
      constexpr int kDupBatch = 8;
      for (uint64_t r_in_block = 0;; r_in_block += kDupBatch) {
        ctr = zero_counters(c);
        {
          ProfScope ps(c, "dup_round", kDupBatch);
          for (int g = 0; g < kDupBatch; ++g) {
            dup_round_kernel<<<shn_grid(hi - lo, kBlock), kBlock, 0, st>>>(
                seg_off.as<uint64_t>(), pt.lo.as<uint32_t>(), pt.count.as<uint32_t>(),
                pt.max_i.as<uint32_t>(), pt.covered.as<uint32_t>(), d_cand_off.as<uint64_t>(), lo, hi,
                st_a.as<uint8_t>(), st_b.as<uint8_t>(), dupf.as<uint8_t>(), ctr + g);
            KERNEL_CHECK();
            std::swap(st_a.p, st_b.p);
            std::swap(st_a.bytes, st_b.bytes);
          }
        }
        read_counters(c, h, kDupBatch);
        bool done = false;
        for (int g = 0; g < kDupBatch; ++g) {
          ++rounds;
          if (h[g] == 0) {
            done = true;
            break;
          }
        }
        if (done) break;
        SHN_CHECK(r_in_block <= hi - lo + 1, "internal error: duplicate filter does not converge");
      }
      // both buffers agree outside [lo, hi); make them agree inside as well
      CUDA_CHECK(cudaMemcpyAsync(st_b.as<uint8_t>() + lo, st_a.as<uint8_t>() + lo, hi - lo,
                                 cudaMemcpyDeviceToDevice, st));
    }
    s->sz.dup_rounds = rounds;
    d2h(c, h_status, st_a.p, n_cand);
    d2h(c, h_dup, dupf.p, n_cand);
  }
  // accepted contigs, acceptance order = pop order
  std::vector<uint32_t> acc_cand;
  std::vector<uint64_t> acc_off(1, 0);
  s->h_cand_dup = h_dup;
  s->h_cand_acc.assign(n_cand, 0);
  for (uint64_t j = 0; j < n_cand; ++j) {
    if (h_status[j] == 1) {
      s->h_cand_acc[j] = 1;
      acc_cand.push_back((uint32_t)j);
      acc_off.push_back(acc_off.back() + (cand_off[j + 1] - cand_off[j]));
    }
  }
  const uint64_t n_contigs = acc_cand.size();
  const uint64_t contig_bases = acc_off.back();
  s->sz.n_contigs = n_contigs;
  s->sz.contig_bases = contig_bases;
  s->h_contig_offs = acc_off;
  h2d(c, s->contig_offs, acc_off);
  s->contig_codes.reserve(std::max<uint64_t>(contig_bases, 1));
  if (contig_bases) {
    DevBuf d_acc;
    h2d(c, d_acc, acc_cand);
    ProfScope ps(c, "copy_contigs");
    copy_contigs_kernel<<<shn_grid(contig_bases, kBlock), kBlock, 0, st>>>(
        cand_codes.as<uint8_t>(), d_cand_off.as<uint64_t>(), d_acc.as<uint32_t>(),
        s->contig_offs.as<uint64_t>(), n_contigs, contig_bases, s->contig_codes.as<uint8_t>());
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  cand_codes.release();

  ht.mark("duplicate filter + accepted contigs");
  // ---- a7: allowed K1-mers with weights, in contig order ------------------------------------
  uint64_t n_allowed = 0;
  if (n_contigs) {
    DevBuf cnt, off, owner, pos;
    cnt.reserve((n_contigs + 1) * 8);
    off.reserve((n_contigs + 1) * 8);
    window_counts_kernel<<<shn_grid(n_contigs + 1, kBlock), kBlock, 0, st>>>(
        s->contig_offs.as<uint64_t>(), n_contigs, k1, cnt.as<uint64_t>());
    KERNEL_CHECK();
    exclusive_sum(c, cnt.as<uint64_t>(), off.as<uint64_t>(), n_contigs + 1);
    CUDA_CHECK(cudaMemcpyAsync(&n_allowed, off.as<uint64_t>() + n_contigs, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    s->allowed_keys.reserve(std::max<uint64_t>(n_allowed, 1) * 8 * SHN_KEY_WORDS);
    s->allowed_w.reserve(std::max<uint64_t>(n_allowed, 1) * 4);
    if (n_allowed) {
      owner.reserve(n_allowed * 4);
      pos.reserve(n_allowed * 4);
      ctr = zero_counters(c);
      ProfScope ps(c, "allowed", 2);
      k1mer_windows_kernel<<<shn_grid(contig_bases, kBlock), kBlock, 0, st>>>(
          s->contig_codes.as<uint8_t>(), s->contig_offs.as<uint64_t>(), off.as<uint64_t>(), n_contigs,
          contig_bases, k1, s->allowed_keys.as<uint64_t>());
      KERNEL_CHECK();
      allowed_weights_kernel<<<shn_grid(n_allowed, kBlock), kBlock, 0, st>>>(
          tv, s->allowed_keys.as<uint64_t>(), n_allowed, s->allowed_w.as<uint32_t>(), ctr);
      KERNEL_CHECK();
      read_counters(c, h, 1);
      SHN_CHECK(h[0] == 0 || allow_missing, "internal error: a contig K1-mer is missing from the table");
    }
  }
  s->sz.n_allowed = n_allowed;

  ht.mark("allowed set");
  // ---- a8: contig C-mer graph (C = K1-1) --------------------------------------------------------
  const int C = k1 - 1;
  s->labels.reserve((n_contigs + 1) * 4);
  label_init_kernel<<<shn_grid(n_contigs + 1, kBlock), kBlock, 0, st>>>(s->labels.as<uint32_t>(),
                                                                        n_contigs + 1);
  KERNEL_CHECK();
  if (n_contigs && C >= 1) {
    DevBuf cnt, off;
    cnt.reserve((n_contigs + 1) * 8);
    off.reserve((n_contigs + 1) * 8);
    window_counts_kernel<<<shn_grid(n_contigs + 1, kBlock), kBlock, 0, st>>>(
        s->contig_offs.as<uint64_t>(), n_contigs, C, cnt.as<uint64_t>());
    KERNEL_CHECK();
    exclusive_sum(c, cnt.as<uint64_t>(), off.as<uint64_t>(), n_contigs + 1);
    uint64_t n_ent = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_ent, off.as<uint64_t>() + n_contigs, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_ent) {
      DevBuf keys, owner, pos;
      keys.reserve(n_ent * 8);
      owner.reserve(n_ent * 4);
      pos.reserve(n_ent * 4);
      {
        ProfScope ps(c, "cmer_entries");
        window_entries_kernel<<<shn_grid(contig_bases, kBlock), kBlock, 0, st>>>(
            s->contig_codes.as<uint8_t>(), s->contig_offs.as<uint64_t>(), off.as<uint64_t>(), n_contigs,
            contig_bases, C, 1u, keys.as<uint64_t>(), owner.as<uint32_t>(), pos.as<uint32_t>());
        KERNEL_CHECK();
      }
      SelfJoin cj;
      cj.prepare(c, "cmer", keys, owner, pos, off.as<uint64_t>(), n_contigs, 1u, n_ent, 2 * C, 1);
      cj.join(0u, 0xFFFFFFFFu, nullptr, &s->edges);
    }
  }
  s->sz.n_edges = s->edges.n;

  ht.mark("contig graph");
  // ---- a9: components of the contig graph (label = minimum contig index) ----------------------
  if (s->edges.n) {
    ProfScope ps(c, "contig_components", 2);
    for (int it = 0;; ++it) {
      ctr = zero_counters(c);
      label_hook_kernel<<<shn_grid(s->edges.n, kBlock), kBlock, 0, st>>>(
          s->edges.lo.as<uint32_t>(), s->edges.hi.as<uint32_t>(), s->edges.n, s->labels.as<uint32_t>(), ctr);
      KERNEL_CHECK();
      label_jump_kernel<<<shn_grid(n_contigs + 1, kBlock), kBlock, 0, st>>>(s->labels.as<uint32_t>(),
                                                                            n_contigs + 1, ctr);
      KERNEL_CHECK();
      read_counters(c, h, 1);
      if (h[0] == 0) break;
      SHN_CHECK(it < 100000, "internal error: label propagation does not converge");
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  ht.mark("contig components (end)");
}

void l3_run(shn_ctx* c, uint32_t min_weight, uint32_t min_length) {
  l3_walks(c, min_weight, min_length);
  l3_filter(c, nullptr, nullptr, 0, 0, 0);
}

// ---- getters -----------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(kBlock)
    seed_keys_kernel(const ShnSlot* __restrict__ slots, const uint32_t* __restrict__ w_slot, uint64_t n,
                     uint64_t* __restrict__ keys) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) shn_store_key(keys, i, slots[w_slot[i]].key);
}

__global__ void __launch_bounds__(kBlock)
    codes_to_ascii_kernel(const uint8_t* __restrict__ codes, uint64_t n, char* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = shn_base_of(codes[i]);
}
}  // namespace

void l3_get_sizes(shn_ctx* c, shn_l3_sizes* out) { *out = need_l3(c)->sz; }

// ---- candidates of this context's walks, for the merge across ranks (device pointers) -----------
namespace {
__global__ void __launch_bounds__(kBlock)
    cand_seed_kernel(const ShnSlot* __restrict__ slots, const uint32_t* __restrict__ cand_walk,
                     const uint32_t* __restrict__ w_slot, uint64_t n, const uint64_t* __restrict__ gline,
                     uint32_t* __restrict__ weight, uint64_t* __restrict__ line) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  shn_key_t key;
  uint32_t wz, wi;
  table_load_slot(slots, w_slot[cand_walk[j]], &key, &wz, &wi);
  weight[j] = wz & SHN_WEIGHT_MASK;
  line[j] = gline ? gline[wi] : (uint64_t)wi;
}
}  // namespace

void l3_cand_sizes(shn_ctx* c, uint64_t* n_cand, uint64_t* n_bases) {
  L3State* s = need_l3(c);
  SHN_CHECK(!s->foreign, "the candidates of this context were replaced by shn_l3_filter");
  *n_cand = s->n_cand;
  *n_bases = s->h_cand_off.back();
}

// pop-order key of every candidate = (seed weight desc, input line of the seed desc)
void l3_cand_export(shn_ctx* c, uint32_t* d_weight, uint64_t* d_line, uint64_t* d_offs,
                    uint8_t* d_codes) {
  L3State* s = need_l3(c);
  SHN_CHECK(!s->foreign, "the candidates of this context were replaced by shn_l3_filter");
  const uint64_t n = s->n_cand, nb = s->h_cand_off.back();
  cudaStream_t st = c->stream;
  if (n) {
    cand_seed_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(table_view(c).slots, s->cand_walk.as<uint32_t>(),
                                                            s->w_seed_slot.as<uint32_t>(), n, c->gline_dev,
                                                            d_weight, d_line);
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(d_offs, s->cand_off.p, (n + 1) * 8, cudaMemcpyDeviceToDevice, st));
  if (nb) CUDA_CHECK(cudaMemcpyAsync(d_codes, s->cand_codes.p, nb, cudaMemcpyDeviceToDevice, st));
}

void l3_set_allowed_weights(shn_ctx* c, const uint32_t* d_w) {
  L3State* s = need_l3(c);
  if (s->sz.n_allowed)
    CUDA_CHECK(cudaMemcpyAsync(s->allowed_w.p, d_w, s->sz.n_allowed * 4, cudaMemcpyDeviceToDevice, c->stream));
}

// device-resident accepted contigs (2-bit codes, offsets) for the L4 map
void l3_contigs_dev(shn_ctx* c, const uint8_t** codes, const uint64_t** offs, uint64_t* n,
                        uint64_t* n_allowed) {
  L3State* s = need_l3(c);
  *codes = s->contig_codes.as<uint8_t>();
  *offs = s->contig_offs.as<uint64_t>();
  *n = s->sz.n_contigs;
  *n_allowed = s->sz.n_allowed;
}

// device-resident allowed set (keys, weights) for the L4 map when caller and callee share the ctx
void l3_allowed_dev(shn_ctx* c, const uint64_t** keys, const uint32_t** weights, uint64_t* n) {
  L3State* s = need_l3(c);
  *keys = s->allowed_keys.as<uint64_t>();
  *weights = s->allowed_w.as<uint32_t>();
  *n = s->sz.n_allowed;
}

void l3_get_walks(shn_ctx* c, uint64_t* seed_keys, uint32_t* n_left, uint32_t* n_right,
                           uint64_t* tot_wt, uint8_t* flags) {
  L3State* s = need_l3(c);
  uint64_t n = s->sz.n_walks;
  if (n == 0) return;
  cudaStream_t st = c->stream;
  if (seed_keys) {
    DevBuf keys;
    keys.reserve(n * 8 * SHN_KEY_WORDS);
    seed_keys_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(table_view(c).slots, s->w_seed_slot.as<uint32_t>(),
                                                            n, keys.as<uint64_t>());
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(seed_keys, keys.p, n * 8 * SHN_KEY_WORDS, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (n_left) CUDA_CHECK(cudaMemcpyAsync(n_left, s->w_nl.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (n_right) CUDA_CHECK(cudaMemcpyAsync(n_right, s->w_nr.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (tot_wt) CUDA_CHECK(cudaMemcpyAsync(tot_wt, s->w_totwt.p, n * 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (flags) {
    memset(flags, 0, n);
    for (size_t j = 0; j < s->h_cand_walk.size(); ++j) {
      uint8_t f = 1;   // bits 1-2 describe this context's own candidates only
      if (!s->foreign && s->h_cand_dup[j]) f |= 2;
      if (!s->foreign && s->h_cand_acc[j]) f |= 4;
      flags[s->h_cand_walk[j]] = f;
    }
  }
}

void l3_get_contigs(shn_ctx* c, char* bases, uint64_t* offsets) {
  L3State* s = need_l3(c);
  uint64_t nb = s->sz.contig_bases;
  memcpy(offsets, s->h_contig_offs.data(), s->h_contig_offs.size() * 8);
  if (nb == 0) return;
  DevBuf ascii;
  ascii.reserve(nb);
  codes_to_ascii_kernel<<<shn_grid(nb, kBlock), kBlock, 0, c->stream>>>(s->contig_codes.as<uint8_t>(), nb,
                                                                       ascii.as<char>());
  KERNEL_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(bases, ascii.p, nb, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void l3_get_allowed(shn_ctx* c, uint64_t* keys, uint32_t* weights) {
  L3State* s = need_l3(c);
  uint64_t n = s->sz.n_allowed;
  if (n == 0) return;
  if (keys)
    CUDA_CHECK(cudaMemcpyAsync(keys, s->allowed_keys.p, n * 8 * SHN_KEY_WORDS, cudaMemcpyDeviceToHost,
                               c->stream));
  if (weights)
    CUDA_CHECK(cudaMemcpyAsync(weights, s->allowed_w.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void l3_get_edges(shn_ctx* c, uint32_t* a, uint32_t* b, uint32_t* weight, uint32_t* fp) {
  L3State* s = need_l3(c);
  uint64_t n = s->edges.n;
  if (n == 0) return;
  cudaStream_t st = c->stream;
  if (a) CUDA_CHECK(cudaMemcpyAsync(a, s->edges.lo.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (b) CUDA_CHECK(cudaMemcpyAsync(b, s->edges.hi.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (weight) CUDA_CHECK(cudaMemcpyAsync(weight, s->edges.count.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (fp) CUDA_CHECK(cudaMemcpyAsync(fp, s->edges.min_i.p, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
}

void l3_get_labels(shn_ctx* c, uint32_t* label) {
  L3State* s = need_l3(c);
  CUDA_CHECK(cudaMemcpyAsync(label, s->labels.p, (s->sz.n_contigs + 1) * 4, cudaMemcpyDeviceToHost,
                             c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace SHN_NS
