// f3: the first step of the consumer of the hot path's output --
// multibridging.load_single_jellyfish (multibridging.py:145-172) + Node.condense_all
// (mbgraph.py:479-507, Edge.condense :184-258): de Bruijn nodes = the K-mer prefix and suffix of
// every K1-mer line, edges = the lines; every unambiguous edge (source with one out-edge,
// destination with one in-edge, source != destination) is condensed, i.e. maximal chains become
// unitig nodes.  The reference condenses edge by edge in list order; the fixed point is the same
// for every order except on pure cycles of unambiguous edges (left uncondensed here and counted).
//   node table (hash, first-appearance order = Node.nodes order)  ->  degrees per node  ->
//   prev pointer along unambiguous edges  ->  pointer jumping (chain head + distance)  ->
//   unitig strings, count / prevalence / norm sums, edges between unitigs
// Edge copy counts follow the reference: an edge re-created by a condensation starts at 0.0, so only
// edges between two single-K-mer nodes keep round(prevalence).
// K <= 32 (one 64-bit word per K-mer); key-width independent of the K1-mer tables: compiled once.
#include <cub/cub.cuh>

#include "common.cuh"
#include "table_dev.cuh"

using namespace narrow;

namespace {
constexpr int kBlock = 256;
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct CondenseState {
  int K = 0;
  uint64_t n_nodes = 0, n_unitigs = 0, n_bases = 0, n_edges = 0, n_cycle_nodes = 0;
  DevBuf bases, offs, count, prevalence, e_src, e_dst, e_cc;
};

void condense_state_free(shn_ctx* c) {
  delete static_cast<CondenseState*>(c->condense);
  c->condense = nullptr;
}

__global__ void __launch_bounds__(kBlock) clear_kernel(ShnSlot* slots, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) table_store_empty(slots, i, 0xFFFFFFFFu);
}

// line i: prefix appears at 2i, suffix at 2i+1 (k1 then k2, multibridging.py:162-166)
__global__ void __launch_bounds__(kBlock)
    insert_nodes_kernel(ShnTableView t, const uint64_t* __restrict__ pre, const uint64_t* __restrict__ suf, uint64_t n,
                        uint32_t* __restrict__ slot_pre, uint32_t* __restrict__ slot_suf, unsigned long long* ctr) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int n_new = 0;
  uint32_t old;
  const uint64_t a = pre[i], b = suf[i];
  if (a == SHN_EMPTY_KEY || b == SHN_EMPTY_KEY) {
    atomicAdd(&ctr[1], 1ull);
    return;
  }
  const uint64_t sa = table_insert_add(t, a, 1u, (uint32_t)(2 * i), &n_new, &old);
  const uint64_t sb = table_insert_add(t, b, 1u, (uint32_t)(2 * i + 1), &n_new, &old);
  if (sa == ~0ull || sb == ~0ull) {
    atomicAdd(&ctr[1], 1ull);
    return;
  }
  slot_pre[i] = (uint32_t)sa;
  slot_suf[i] = (uint32_t)sb;
  if (n_new) atomicAdd(&ctr[0], (unsigned long long)n_new);
}

struct Occupied {
  const ShnSlot* slots;
  __device__ bool operator()(uint32_t i) const { return slots[i].key != SHN_EMPTY_KEY; }
};

__global__ void __launch_bounds__(kBlock)
    node_order_kernel(const ShnSlot* __restrict__ slots, const uint32_t* __restrict__ occ, uint64_t n_nodes,
                      uint32_t* __restrict__ first_idx) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_nodes) first_idx[j] = slots[occ[j]].idx;
}

__global__ void __launch_bounds__(kBlock)
    node_ids_kernel(const ShnSlot* __restrict__ slots, const uint32_t* __restrict__ occ_sorted, uint64_t n_nodes,
                    uint32_t* __restrict__ node_of_slot, uint64_t* __restrict__ node_key) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_nodes) return;
  const uint32_t s = occ_sorted[j];
  node_of_slot[s] = (uint32_t)j;
  node_key[j] = slots[s].key;
}

__global__ void __launch_bounds__(kBlock)
    degrees_kernel(const uint32_t* __restrict__ slot_pre, const uint32_t* __restrict__ slot_suf,
                   const uint32_t* __restrict__ node_of_slot, uint64_t n, uint32_t* __restrict__ src,
                   uint32_t* __restrict__ dst, uint32_t* __restrict__ out_deg, uint32_t* __restrict__ in_deg) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = node_of_slot[slot_pre[i]], d = node_of_slot[slot_suf[i]];
  src[i] = s;
  dst[i] = d;
  atomicAdd(&out_deg[s], 1u);
  atomicAdd(&in_deg[d], 1u);
}

__device__ __forceinline__ bool unambiguous(uint32_t s, uint32_t d, const uint32_t* out_deg, const uint32_t* in_deg) {
  return out_deg[s] == 1u && in_deg[d] == 1u && s != d;  // mbgraph.py:488-492
}

// anc[v] = predecessor along the unambiguous in-edge (v itself for a chain head), dist = 1 / 0
__global__ void __launch_bounds__(kBlock)
    chain_init_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst,
                      const uint32_t* __restrict__ out_deg, const uint32_t* __restrict__ in_deg, uint64_t n,
                      uint32_t* __restrict__ prev) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = src[i], d = dst[i];
  if (unambiguous(s, d, out_deg, in_deg)) prev[d] = s;  // at most one such edge per destination
}

__global__ void __launch_bounds__(kBlock)
    chain_start_kernel(const uint32_t* __restrict__ prev, uint64_t n_nodes, uint32_t* __restrict__ anc,
                       uint32_t* __restrict__ dist) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const uint32_t p = prev[v];
  anc[v] = p == kNone ? (uint32_t)v : p;
  dist[v] = p == kNone ? 0u : 1u;
}

__global__ void __launch_bounds__(kBlock)
    chain_jump_kernel(const uint32_t* __restrict__ anc, const uint32_t* __restrict__ dist, uint64_t n_nodes,
                      uint32_t* __restrict__ anc2, uint32_t* __restrict__ dist2, unsigned long long* changed) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const uint32_t a = anc[v];
  const uint32_t aa = anc[a];
  anc2[v] = aa;
  dist2[v] = dist[v] + dist[a];  // steps to a, plus a's steps to its own ancestor (0 for a head)
  if (aa != a) atomicAdd(changed, 1ull);
}

// pure cycles never reach a head: their nodes stay single (prev cleared), heads get flag 1
__global__ void __launch_bounds__(kBlock)
    chain_finish_kernel(uint32_t* __restrict__ prev, uint32_t* __restrict__ anc, uint32_t* __restrict__ dist,
                        uint64_t n_nodes, int converged, uint32_t* __restrict__ is_head, unsigned long long* n_cycle) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  bool cyc = false;
  if (!converged) {
    const uint32_t a = anc[v];
    cyc = prev[a] != kNone;  // after > log2(n) doublings the ancestor of a chain node is its head
  }
  if (cyc) {
    anc[v] = (uint32_t)v;
    dist[v] = 0;
    atomicAdd(n_cycle, 1ull);
  }
  is_head[v] = (cyc || prev[v] == kNone) ? 1u : 0u;
}
__global__ void __launch_bounds__(kBlock)
    cycle_unlink_kernel(uint32_t* __restrict__ prev, const uint32_t* __restrict__ anc, uint64_t n_nodes) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n_nodes && anc[v] == (uint32_t)v) prev[v] = kNone;
}

__global__ void __launch_bounds__(kBlock)
    unitig_sums_kernel(const uint32_t* __restrict__ anc, const uint32_t* __restrict__ uid_of_head,
                       const uint32_t* __restrict__ out_deg, uint64_t n_nodes, int K, uint32_t* __restrict__ cnt,
                       unsigned long long* __restrict__ prevalence) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const uint32_t u = uid_of_head[anc[v]];
  atomicAdd(&cnt[u], 1u);
  // node.prevalence = sum(e.weight for e in node.out_edges) with weight = K-1 (multibridging.py:168,171)
  atomicAdd(&prevalence[u], (unsigned long long)(K - 1) * out_deg[v]);
}

__global__ void __launch_bounds__(kBlock)
    unitig_len_kernel(const uint32_t* __restrict__ cnt, uint64_t n_unitigs, int K, uint64_t* __restrict__ len) {
  uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u > n_unitigs) return;
  len[u] = u < n_unitigs ? (uint64_t)K + cnt[u] - 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
    unitig_bases_kernel(const uint64_t* __restrict__ node_key, const uint32_t* __restrict__ anc,
                        const uint32_t* __restrict__ dist, const uint32_t* __restrict__ uid_of_head,
                        const uint64_t* __restrict__ offs, uint64_t n_nodes, int K, char* __restrict__ bases) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const uint64_t key = node_key[v];
  const uint64_t o = offs[uid_of_head[anc[v]]];
  const uint32_t r = dist[v];
  if (r == 0) {
    for (int j = 0; j < K; ++j) bases[o + j] = shn_base_of((uint32_t)(key >> (2 * (K - 1 - j))) & 3u);
  } else {
    bases[o + K - 1 + r] = shn_base_of((uint32_t)key & 3u);  // new_bases = source.bases + destination.bases[K-1:]
  }
}

__global__ void __launch_bounds__(kBlock)
    unitig_edges_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst,
                        const uint32_t* __restrict__ prev, const uint32_t* __restrict__ anc,
                        const uint32_t* __restrict__ uid_of_head, const uint32_t* __restrict__ cnt,
                        const uint32_t* __restrict__ prevalence_in, uint64_t n, uint32_t* __restrict__ e_src,
                        uint32_t* __restrict__ e_dst, uint32_t* __restrict__ e_cc, unsigned long long* cursor) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = src[i], d = dst[i];
  // condensed lines: the one unambiguous in-edge of d (a repeated line makes out_deg 2: never condensed)
  if (prev[d] == s && s != d) return;
  const uint32_t us = uid_of_head[anc[s]], ud = uid_of_head[anc[d]];
  const unsigned long long o = atomicAdd(cursor, 1ull);
  e_src[o] = us;
  e_dst[o] = ud;
  e_cc[o] = (cnt[us] == 1u && cnt[ud] == 1u) ? prevalence_in[i] : 0u;  // re-created edges start at 0.0
}

void excl_sum_u32(shn_ctx* c, const uint32_t* in, uint32_t* out, uint64_t n) {
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, c->stream));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, in, out, (int64_t)n, c->stream));
}
}  // namespace

void shn_condense_run_impl(shn_ctx* c, const uint64_t* h_pre, const uint64_t* h_suf, const uint32_t* h_prev,
                           uint64_t n, int K, uint64_t* n_unitigs, uint64_t* n_bases, uint64_t* n_edges,
                           uint64_t* n_cycle_nodes) {
  SHN_CHECK(K >= 2 && K <= 32, "K must be in 2..32");
  SHN_CHECK(n < 0x7FFFFFFFull, "at most 2^31-1 K1-mer lines per component");
  if (c->condense && c->condense_free) c->condense_free(c);
  CondenseState* s = new CondenseState();
  c->condense = s;
  c->condense_free = &condense_state_free;
  s->K = K;
  *n_unitigs = *n_bases = *n_edges = *n_cycle_nodes = 0;
  if (n == 0) return;
  cudaStream_t st = c->stream;
  DevBuf pre, suf, pin, table, slot_pre, slot_suf;
  pre.reserve(n * 8);
  suf.reserve(n * 8);
  pin.reserve(n * 4);
  CUDA_CHECK(cudaMemcpyAsync(pre.p, h_pre, n * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(suf.p, h_suf, n * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(pin.p, h_prev, n * 4, cudaMemcpyHostToDevice, st));
  const uint64_t nb = std::max<uint64_t>(256, (4 * n + SHN_BSLOTS - 1) / SHN_BSLOTS);   // 2n K-mers, load <= 0.5
  const uint64_t n_slots = nb * SHN_BSLOTS;
  SHN_CHECK(n_slots < 0xFFFFFFFFull, "node table too large");
  table.reserve(n_slots * sizeof(ShnSlot));
  slot_pre.reserve(n * 4);
  slot_suf.reserve(n * 4);
  ShnTableView tv{table.as<ShnSlot>(), nb};
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), st));
  unsigned long long h[4];
  {
    ProfScope ps(c, "condense_nodes", 2);
    clear_kernel<<<shn_grid(n_slots, kBlock), kBlock, 0, st>>>(tv.slots, n_slots);
    insert_nodes_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(tv, pre.as<uint64_t>(), suf.as<uint64_t>(), n,
                                                               slot_pre.as<uint32_t>(), slot_suf.as<uint32_t>(), ctr);
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, 16, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  SHN_CHECK(h[1] == 0, "K-mer equal to the free-slot marker (poly-T of 32 bases) or node table full");
  const uint64_t n_nodes = h[0];
  s->n_nodes = n_nodes;
  // node ids in first-appearance order (= the order of Node.nodes after loading)
  DevBuf occ, nsel, fidx, fidx_s, occ_s, node_of_slot, node_key;
  occ.reserve(n_nodes * 4);
  nsel.reserve(8);
  fidx.reserve(n_nodes * 4);
  fidx_s.reserve(n_nodes * 4);
  occ_s.reserve(n_nodes * 4);
  node_of_slot.reserve(n_slots * 4);
  node_key.reserve(n_nodes * 8);
  {
    ProfScope ps(c, "condense_order", 4);
    cub::CountingInputIterator<uint32_t> it(0);
    Occupied pred{tv.slots};
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceSelect::If(nullptr, tb, it, occ.as<uint32_t>(), nsel.as<uint64_t>(), (int64_t)n_slots, pred, st));
    CUDA_CHECK(cub::DeviceSelect::If(c->tmp(tb), tb, it, occ.as<uint32_t>(), nsel.as<uint64_t>(), (int64_t)n_slots, pred, st));
    node_order_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(tv.slots, occ.as<uint32_t>(), n_nodes,
                                                                   fidx.as<uint32_t>());
    KERNEL_CHECK();
    tb = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, fidx.as<uint32_t>(), fidx_s.as<uint32_t>(),
                                               occ.as<uint32_t>(), occ_s.as<uint32_t>(), (int64_t)n_nodes, 0, 32, st));
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, fidx.as<uint32_t>(), fidx_s.as<uint32_t>(),
                                               occ.as<uint32_t>(), occ_s.as<uint32_t>(), (int64_t)n_nodes, 0, 32, st));
    node_ids_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(tv.slots, occ_s.as<uint32_t>(), n_nodes,
                                                                 node_of_slot.as<uint32_t>(), node_key.as<uint64_t>());
    KERNEL_CHECK();
  }
  // degrees, unambiguous edges, chains
  DevBuf src, dst, out_deg, in_deg, prev, anc, dist, anc2, dist2, is_head, uid;
  src.reserve(n * 4);
  dst.reserve(n * 4);
  for (DevBuf* b : {&out_deg, &in_deg, &prev, &anc, &dist, &anc2, &dist2}) b->reserve(n_nodes * 4);
  is_head.reserve((n_nodes + 1) * 4);
  uid.reserve((n_nodes + 1) * 4);
  CUDA_CHECK(cudaMemsetAsync(out_deg.p, 0, n_nodes * 4, st));
  CUDA_CHECK(cudaMemsetAsync(in_deg.p, 0, n_nodes * 4, st));
  CUDA_CHECK(cudaMemsetAsync(prev.p, 0xFF, n_nodes * 4, st));
  {
    ProfScope ps(c, "condense_degrees", 3);
    degrees_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(slot_pre.as<uint32_t>(), slot_suf.as<uint32_t>(),
                                                          node_of_slot.as<uint32_t>(), n, src.as<uint32_t>(),
                                                          dst.as<uint32_t>(), out_deg.as<uint32_t>(), in_deg.as<uint32_t>());
    chain_init_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(src.as<uint32_t>(), dst.as<uint32_t>(),
                                                             out_deg.as<uint32_t>(), in_deg.as<uint32_t>(), n,
                                                             prev.as<uint32_t>());
    chain_start_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(prev.as<uint32_t>(), n_nodes, anc.as<uint32_t>(),
                                                                    dist.as<uint32_t>());
    KERNEL_CHECK();
  }
  int converged = 0;
  {
    ProfScope ps(c, "condense_jump", 40);
    int max_it = 2;
    while ((1ull << max_it) < n_nodes + 2) ++max_it;
    for (int it = 0; it <= max_it; ++it) {
      CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
      chain_jump_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(anc.as<uint32_t>(), dist.as<uint32_t>(), n_nodes,
                                                                     anc2.as<uint32_t>(), dist2.as<uint32_t>(), ctr);
      KERNEL_CHECK();
      std::swap(anc.p, anc2.p);
      std::swap(dist.p, dist2.p);
      CUDA_CHECK(cudaMemcpyAsync(h, ctr, 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (h[0] == 0) {
        converged = 1;
        break;
      }
    }
  }
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 16, st));
  CUDA_CHECK(cudaMemsetAsync(is_head.as<uint32_t>() + n_nodes, 0, 4, st));
  chain_finish_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(prev.as<uint32_t>(), anc.as<uint32_t>(),
                                                                   dist.as<uint32_t>(), n_nodes, converged,
                                                                   is_head.as<uint32_t>(), ctr);
  KERNEL_CHECK();
  if (!converged) {
    cycle_unlink_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(prev.as<uint32_t>(), anc.as<uint32_t>(), n_nodes);
    KERNEL_CHECK();
  }
  excl_sum_u32(c, is_head.as<uint32_t>(), uid.as<uint32_t>(), n_nodes + 1);
  uint32_t nu = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nu, uid.as<uint32_t>() + n_nodes, 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  s->n_unitigs = nu;
  s->n_cycle_nodes = h[0];
  // unitig attributes and strings
  DevBuf len;
  s->count.reserve(std::max<uint64_t>(nu, 1) * 4);
  s->prevalence.reserve(std::max<uint64_t>(nu, 1) * 8);
  s->offs.reserve((uint64_t)(nu + 1) * 8);
  len.reserve((uint64_t)(nu + 1) * 8);
  CUDA_CHECK(cudaMemsetAsync(s->count.p, 0, std::max<uint64_t>(nu, 1) * 4, st));
  CUDA_CHECK(cudaMemsetAsync(s->prevalence.p, 0, std::max<uint64_t>(nu, 1) * 8, st));
  {
    ProfScope ps(c, "condense_unitigs", 4);
    unitig_sums_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(
        anc.as<uint32_t>(), uid.as<uint32_t>(), out_deg.as<uint32_t>(), n_nodes, K, s->count.as<uint32_t>(),
        s->prevalence.as<unsigned long long>());
    unitig_len_kernel<<<shn_grid((uint64_t)nu + 1, kBlock), kBlock, 0, st>>>(s->count.as<uint32_t>(), nu, K,
                                                                            len.as<uint64_t>());
    KERNEL_CHECK();
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.as<uint64_t>(), s->offs.as<uint64_t>(), (int64_t)nu + 1, st));
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, len.as<uint64_t>(), s->offs.as<uint64_t>(), (int64_t)nu + 1, st));
    CUDA_CHECK(cudaMemcpyAsync(&s->n_bases, s->offs.as<uint64_t>() + nu, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    s->bases.reserve(std::max<uint64_t>(s->n_bases, 1));
    unitig_bases_kernel<<<shn_grid(n_nodes, kBlock), kBlock, 0, st>>>(
        node_key.as<uint64_t>(), anc.as<uint32_t>(), dist.as<uint32_t>(), uid.as<uint32_t>(), s->offs.as<uint64_t>(),
        n_nodes, K, s->bases.as<char>());
    KERNEL_CHECK();
  }
  s->e_src.reserve(n * 4);
  s->e_dst.reserve(n * 4);
  s->e_cc.reserve(n * 4);
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
  {
    ProfScope ps(c, "condense_edges");
    unitig_edges_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(
        src.as<uint32_t>(), dst.as<uint32_t>(), prev.as<uint32_t>(), anc.as<uint32_t>(), uid.as<uint32_t>(),
        s->count.as<uint32_t>(), pin.as<uint32_t>(), n, s->e_src.as<uint32_t>(), s->e_dst.as<uint32_t>(),
        s->e_cc.as<uint32_t>(), ctr);
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(h, ctr, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  s->n_edges = h[0];
  *n_unitigs = s->n_unitigs;
  *n_bases = s->n_bases;
  *n_edges = s->n_edges;
  *n_cycle_nodes = s->n_cycle_nodes;
}

void shn_condense_get_impl(shn_ctx* c, char* bases, uint64_t* offsets, uint32_t* count, uint64_t* prevalence,
                           uint32_t* e_src, uint32_t* e_dst, uint32_t* e_cc) {
  SHN_CHECK(c->condense != nullptr, "shn_condense_run has not been called on this context");
  CondenseState* s = static_cast<CondenseState*>(c->condense);
  cudaStream_t st = c->stream;
  const uint64_t nu = s->n_unitigs, ne = s->n_edges;
  if (nu) {
    if (bases && s->n_bases) CUDA_CHECK(cudaMemcpyAsync(bases, s->bases.p, s->n_bases, cudaMemcpyDeviceToHost, st));
    if (offsets) CUDA_CHECK(cudaMemcpyAsync(offsets, s->offs.p, (nu + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (count) CUDA_CHECK(cudaMemcpyAsync(count, s->count.p, nu * 4, cudaMemcpyDeviceToHost, st));
    if (prevalence) CUDA_CHECK(cudaMemcpyAsync(prevalence, s->prevalence.p, nu * 8, cudaMemcpyDeviceToHost, st));
  } else if (offsets) {
    offsets[0] = 0;
  }
  if (ne) {
    if (e_src) CUDA_CHECK(cudaMemcpyAsync(e_src, s->e_src.p, ne * 4, cudaMemcpyDeviceToHost, st));
    if (e_dst) CUDA_CHECK(cudaMemcpyAsync(e_dst, s->e_dst.p, ne * 4, cudaMemcpyDeviceToHost, st));
    if (e_cc) CUDA_CHECK(cudaMemcpyAsync(e_cc, s->e_cc.p, ne * 4, cudaMemcpyDeviceToHost, st));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
}
