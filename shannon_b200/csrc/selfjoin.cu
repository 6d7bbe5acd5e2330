// See selfjoin.cuh.  Sort-based, race-free, deterministic.
//
// prepare(): the entries are sorted by key once (stable, so inside a run of equal keys they stay in
// generation order = owner ascending, pos ascending).  Every entry g then knows, in GENERATION order,
// the range [rs, ge) of sorted positions that holds its partners: the part of its run in front of its
// own owner's group.
//
// join(lo, hi): one thread per entry of the owners [lo, hi) -- a contiguous range of generation
// order -- emits one match event (d, j, i) per live partner.  Partners are sorted by owner, so the
// ones inside the block are the tail of [rs, ge) (found by a binary search for `lo`); of the ones in
// front of the block only the accepted count, and those are read from a compacted list (prefix counts
// per sorted position), so rejected candidates -- the bulk -- are never looked at again.  The events
// come out ordered by (j, i); a STABLE radix sort on d alone (bits_owner bits instead of the whole
// 64-bit (j, d, i) key) groups them by (d, j) with i still ascending, one warp per group reduces it,
// and the (few) pairs are put into (j, d) order at the end.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <memory>

#include "selfjoin.cuh"

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) iota_kernel(uint32_t* a, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    gather_heads_kernel(const uint64_t* __restrict__ keys_s, const uint32_t* __restrict__ idx_s,
                        const uint32_t* __restrict__ owner, uint64_t n, uint32_t* __restrict__ owner_s,
                        uint32_t* __restrict__ run_head, uint32_t* __restrict__ grp_head) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // one random gather per entry: the previous entry's owner comes from the neighbouring lane
  const uint32_t o = p < n ? owner[idx_s[p]] : 0u;
  uint32_t o_prev = __shfl_up_sync(0xFFFFFFFFu, o, 1);
  if (p >= n) return;
  if (lane == 0 && p > 0) o_prev = owner[idx_s[p - 1]];
  owner_s[p] = o;
  bool hk = p == 0 || keys_s[p] != keys_s[p - 1];
  bool hg = hk || o_prev != o;
  run_head[p] = hk ? (uint32_t)p : 0u;
  grp_head[p] = hg ? (uint32_t)p : 0u;
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// sorted position p -> generation index g = idx_s[p]; one 8-byte store per entry
__global__ void __launch_bounds__(kBlock)
    scatter_ranges_kernel(const uint32_t* __restrict__ idx_s, const uint32_t* __restrict__ run_start,
                          const uint32_t* __restrict__ grp_start, uint64_t n, uint2* __restrict__ range_g) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  range_g[idx_s[p]] = make_uint2(run_start[p], grp_start[p]);
}

// accepted entries in front of the block, per sorted position
__global__ void __launch_bounds__(kBlock)
    acc_flag_kernel(const uint32_t* __restrict__ owner_s, uint64_t n, uint32_t lo,
                    const uint8_t* __restrict__ status, uint32_t* __restrict__ flag) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q > n) return;
  uint32_t f = 0;
  if (q < n) {
    const uint32_t o = owner_s[q];
    f = (o < lo && status[o] == 1) ? 1u : 0u;
  }
  flag[q] = f;
}
__global__ void __launch_bounds__(kBlock)
    acc_list_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ acc_cnt, uint64_t n,
                    uint32_t* __restrict__ acc_q) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n && flag[q]) acc_q[acc_cnt[q]] = (uint32_t)q;
}

// first sorted position in [a, b) whose owner is >= lo (owners ascend inside a run)
__device__ __forceinline__ uint32_t lower_bound_owner(const uint32_t* __restrict__ owner_s, uint32_t a,
                                                      uint32_t b, uint32_t lo) {
  while (a < b) {
    const uint32_t m = a + ((b - a) >> 1);
    if (owner_s[m] < lo)
      a = m + 1;
    else
      b = m;
  }
  return a;
}

__global__ void __launch_bounds__(kBlock)
    count_events_kernel(const uint2* __restrict__ range_g, const uint32_t* __restrict__ owner_s, const uint32_t* __restrict__ acc_cnt,
                        uint64_t g_lo, uint64_t m, uint32_t lo, uint32_t* __restrict__ cnt,
                        uint32_t* __restrict__ lb_out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > m) return;
  uint32_t c = 0, lb = 0;
  if (t < m) {
    const uint2 rg = range_g[g_lo + t];
    const uint32_t rs = rg.x, ge = rg.y;
    lb = ge;
    if (rs < ge) {
      lb = lo ? lower_bound_owner(owner_s, rs, ge, lo) : rs;
      c = ge - lb;
      if (acc_cnt) c += acc_cnt[lb] - acc_cnt[rs];
    }
  }
  cnt[t] = c;
  lb_out[t] = lb;
}

template <typename V>
__global__ void __launch_bounds__(kBlock)
    emit_events_kernel(const uint64_t* __restrict__ ev_off, const uint2* __restrict__ range_g,
                       const uint32_t* __restrict__ lb_in,
                       const uint32_t* __restrict__ owner_s, const uint32_t* __restrict__ acc_cnt,
                       const uint32_t* __restrict__ acc_q, const uint32_t* __restrict__ owner_g,
                       const uint32_t* __restrict__ pos_g, uint64_t g_lo, uint64_t m, uint32_t lo,
                       int bits_pos, uint32_t* __restrict__ ev_d, V* __restrict__ ev_v) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  uint64_t e = ev_off[t];
  if (ev_off[t + 1] == e) return;
  const uint64_t g = g_lo + t;
  const V val = (V)(((V)(owner_g[g] - lo) << bits_pos) | (V)pos_g[g]);
  const uint2 rg = range_g[g];
  const uint32_t rs = rg.x, ge = rg.y, lb = lb_in[t];
  if (acc_cnt) {
    for (uint32_t a = acc_cnt[rs], b = acc_cnt[lb]; a < b; ++a) {
      ev_d[e] = owner_s[acc_q[a]];
      ev_v[e++] = val;
    }
  }
  for (uint32_t q = lb; q < ge; ++q) {
    ev_d[e] = owner_s[q];
    ev_v[e++] = val;
  }
}

// events sorted by (d, j, i): 1 at the first event of every (d, j) group
template <typename V>
__global__ void __launch_bounds__(kBlock)
    pair_heads_kernel(const uint32_t* __restrict__ d_s, const V* __restrict__ v_s, uint64_t n_events,
                      int bits_pos, uint32_t* __restrict__ head) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > n_events) return;
  uint32_t h = 0;
  if (e < n_events)
    h = (e == 0 || d_s[e] != d_s[e - 1] || (v_s[e] >> bits_pos) != (v_s[e - 1] >> bits_pos)) ? 1u : 0u;
  head[e] = h;
}
__global__ void __launch_bounds__(kBlock)
    pair_starts_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ pid, uint64_t n_events,
                       uint32_t* __restrict__ start) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > n_events) return;
  if (e == n_events)
    start[pid[e]] = (uint32_t)n_events;  // pid[n_events] = number of pairs
  else if (head[e])
    start[pid[e]] = (uint32_t)e;
}

// one warp per (d, j) group: count, first / last position, covered = |union of [i, i+r)|
template <typename V>
__global__ void __launch_bounds__(kBlock)
    reduce_pairs_kernel(const uint32_t* __restrict__ d_s, const V* __restrict__ v_s,
                        const uint32_t* __restrict__ start, uint64_t n_pairs, int bits_pos, uint32_t r,
                        uint32_t* __restrict__ jrel, uint32_t* __restrict__ p_lo,
                        uint32_t* __restrict__ count, uint32_t* __restrict__ min_i,
                        uint32_t* __restrict__ max_i, uint32_t* __restrict__ covered) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_pairs) return;  // whole warps only
  const uint32_t s = start[w], t = start[w + 1];
  const V pm = (V)(((V)1 << bits_pos) - 1);
  uint32_t cov = 0;
  for (uint32_t e = s + lane; e + 1 < t; e += 32) {
    const uint32_t gap = (uint32_t)(v_s[e + 1] & pm) - (uint32_t)(v_s[e] & pm);
    cov += gap < r ? gap : r;
  }
  cov = __reduce_add_sync(0xFFFFFFFFu, cov);
  if (lane == 0) {
    const V first = v_s[s], last = v_s[t - 1];
    jrel[w] = (uint32_t)(first >> bits_pos);
    p_lo[w] = d_s[s];
    count[w] = t - s;
    min_i[w] = (uint32_t)(first & pm);
    max_i[w] = (uint32_t)(last & pm);
    covered[w] = cov + r;  // the last position of the group contributes a full interval
  }
}

__global__ void __launch_bounds__(kBlock)
    order_pairs_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ jrel_s, uint64_t n,
                       uint32_t lo, const uint32_t* __restrict__ t_lo, const uint32_t* __restrict__ t_count,
                       const uint32_t* __restrict__ t_min, const uint32_t* __restrict__ t_max,
                       const uint32_t* __restrict__ t_cov, uint32_t* hi, uint32_t* olo, uint32_t* count,
                       uint32_t* min_i, uint32_t* max_i, uint32_t* covered) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t w = perm[t];
  hi[t] = jrel_s[t] + lo;
  olo[t] = t_lo[w];
  count[t] = t_count[w];
  min_i[t] = t_min[w];
  max_i[t] = t_max[w];
  covered[t] = t_cov[w];
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) ++b;
  return b;
}

struct CastU64 {
  __device__ __forceinline__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};

void take(DevBuf& dst, DevBuf& src) {
  dst.release();
  std::swap(dst.p, src.p);
  std::swap(dst.bytes, src.bytes);
  std::swap(dst.pool, src.pool);
}

template <typename V>
void sort_reduce(SelfJoin* sj, uint64_t n_events, uint32_t lo, int jrel_bits, DevBuf& ev_d, DevBuf& ev_v,
                 PairTable* out) {
  shn_ctx* c = sj->c;
  cudaStream_t st = c->stream;
  const std::string n_sort2 = sj->tag + "_sort_events", n_reduce = sj->tag + "_reduce_pairs";
  const int bp = sj->bits_pos;
  DevBuf d_s, v_s;
  d_s.reserve(n_events * 4);
  v_s.reserve(n_events * sizeof(V));
  std::unique_ptr<ProfScope> ps(new ProfScope(c, n_sort2.c_str(), 1));
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, ev_d.as<uint32_t>(), d_s.as<uint32_t>(), ev_v.as<V>(),
                                             v_s.as<V>(), (int64_t)n_events, 0, sj->bits_owner, st));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, ev_d.as<uint32_t>(), d_s.as<uint32_t>(),
                                             ev_v.as<V>(), v_s.as<V>(), (int64_t)n_events, 0, sj->bits_owner,
                                             st));
  ps.reset();
  ev_d.release();
  ev_v.release();
  ps.reset(new ProfScope(c, n_reduce.c_str(), 7));
  DevBuf head, pid;
  head.reserve((n_events + 1) * 4);
  pid.reserve((n_events + 1) * 4);
  pair_heads_kernel<V><<<shn_grid(n_events + 1, kBlock), kBlock, 0, st>>>(d_s.as<uint32_t>(), v_s.as<V>(),
                                                                          n_events, bp, head.as<uint32_t>());
  KERNEL_CHECK();
  tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, head.as<uint32_t>(), pid.as<uint32_t>(),
                                           (int64_t)(n_events + 1), st));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, head.as<uint32_t>(), pid.as<uint32_t>(),
                                           (int64_t)(n_events + 1), st));
  uint32_t n_pairs32 = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_pairs32, pid.as<uint32_t>() + n_events, 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  const uint64_t n_pairs = n_pairs32;
  DevBuf start, t_j, t_lo, t_count, t_min, t_max, t_cov;
  start.reserve((n_pairs + 1) * 4);
  for (DevBuf* b : {&t_j, &t_lo, &t_count, &t_min, &t_max, &t_cov}) b->reserve(n_pairs * 4);
  pair_starts_kernel<<<shn_grid(n_events + 1, kBlock), kBlock, 0, st>>>(head.as<uint32_t>(), pid.as<uint32_t>(),
                                                                       n_events, start.as<uint32_t>());
  KERNEL_CHECK();
  reduce_pairs_kernel<V><<<shn_grid(n_pairs * 32, kBlock), kBlock, 0, st>>>(
      d_s.as<uint32_t>(), v_s.as<V>(), start.as<uint32_t>(), n_pairs, bp, sj->r, t_j.as<uint32_t>(),
      t_lo.as<uint32_t>(), t_count.as<uint32_t>(), t_min.as<uint32_t>(), t_max.as<uint32_t>(),
      t_cov.as<uint32_t>());
  KERNEL_CHECK();
  // (d, j) order -> (j, d) order: stable sort of the pairs on j
  DevBuf iota, perm, j_s;
  iota.reserve(n_pairs * 4);
  perm.reserve(n_pairs * 4);
  j_s.reserve(n_pairs * 4);
  iota_kernel<<<shn_grid(n_pairs, kBlock), kBlock, 0, st>>>(iota.as<uint32_t>(), n_pairs);
  KERNEL_CHECK();
  tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, t_j.as<uint32_t>(), j_s.as<uint32_t>(),
                                             iota.as<uint32_t>(), perm.as<uint32_t>(), (int64_t)n_pairs, 0,
                                             jrel_bits, st));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, t_j.as<uint32_t>(), j_s.as<uint32_t>(),
                                             iota.as<uint32_t>(), perm.as<uint32_t>(), (int64_t)n_pairs, 0,
                                             jrel_bits, st));
  for (DevBuf* b : {&out->hi, &out->lo, &out->count, &out->min_i, &out->max_i, &out->covered})
    b->reserve(n_pairs * 4);
  order_pairs_kernel<<<shn_grid(n_pairs, kBlock), kBlock, 0, st>>>(
      perm.as<uint32_t>(), j_s.as<uint32_t>(), n_pairs, lo, t_lo.as<uint32_t>(), t_count.as<uint32_t>(),
      t_min.as<uint32_t>(), t_max.as<uint32_t>(), t_cov.as<uint32_t>(), out->hi.as<uint32_t>(),
      out->lo.as<uint32_t>(), out->count.as<uint32_t>(), out->min_i.as<uint32_t>(), out->max_i.as<uint32_t>(),
      out->covered.as<uint32_t>());
  KERNEL_CHECK();
  ps.reset();
  CUDA_CHECK(cudaStreamSynchronize(st));
  out->n = n_pairs;
}

}  // namespace

void SelfJoin::prepare(shn_ctx* ctx, const char* tag_, DevBuf& keys, DevBuf& owner, DevBuf& pos,
                       const uint64_t* d_ent_off, uint64_t n_owner_, uint32_t owner_base_, uint64_t n_,
                       int key_bits, uint32_t r_) {
  c = ctx;
  tag = tag_;
  n = n_;
  r = r_;
  n_owner = n_owner_;
  owner_base = owner_base_;
  if (n == 0) return;
  SHN_CHECK(n < 0xFFFFFFFFull, "self-join: more than 2^32-1 entries");
  cudaStream_t st = c->stream;
  h_ent_off.resize(n_owner + 1);
  CUDA_CHECK(cudaMemcpyAsync(h_ent_off.data(), d_ent_off, (n_owner + 1) * 8, cudaMemcpyDeviceToHost, st));
  take(owner_g, owner);
  take(pos_g, pos);
  const std::string n_sort = tag + "_sort_keys", n_heads = tag + "_runs";
  DevBuf keys_s, idx, idx_s, run_head, grp_head, run_start, grp_start;
  keys_s.reserve(n * 8);
  idx.reserve(n * 4);
  idx_s.reserve(n * 4);
  std::unique_ptr<ProfScope> ps(new ProfScope(c, n_sort.c_str(), 2));
  iota_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(idx.as<uint32_t>(), n);
  KERNEL_CHECK();
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint64_t>(), keys_s.as<uint64_t>(),
                                             idx.as<uint32_t>(), idx_s.as<uint32_t>(), (int64_t)n, 0,
                                             key_bits, st));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, keys.as<uint64_t>(), keys_s.as<uint64_t>(),
                                             idx.as<uint32_t>(), idx_s.as<uint32_t>(), (int64_t)n, 0,
                                             key_bits, st));
  ps.reset();
  keys.release();
  idx.release();
  ps.reset(new ProfScope(c, n_heads.c_str(), 6));
  owner_s.reserve(n * 4);
  run_head.reserve(n * 4);
  grp_head.reserve(n * 4);
  run_start.reserve(n * 4);
  grp_start.reserve(n * 4);
  range_g.reserve(n * 8);
  gather_heads_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(
      keys_s.as<uint64_t>(), idx_s.as<uint32_t>(), owner_g.as<uint32_t>(), n, owner_s.as<uint32_t>(),
      run_head.as<uint32_t>(), grp_head.as<uint32_t>());
  KERNEL_CHECK();
  tb = 0;
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, tb, run_head.as<uint32_t>(),
                                            run_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(c->tmp(tb), tb, run_head.as<uint32_t>(),
                                            run_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(c->tmp(tb), tb, grp_head.as<uint32_t>(),
                                            grp_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  scatter_ranges_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(idx_s.as<uint32_t>(), run_start.as<uint32_t>(),
                                                               grp_start.as<uint32_t>(), n, range_g.as<uint2>());
  KERNEL_CHECK();
  // bit budget of the event words: d | (j - lo, i)
  DevBuf mx;
  mx.reserve(8);
  tb = 0;
  CUDA_CHECK(cub::DeviceReduce::Max(nullptr, tb, pos_g.as<uint32_t>(), mx.as<uint32_t>(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceReduce::Max(c->tmp(tb), tb, pos_g.as<uint32_t>(), mx.as<uint32_t>(), (int64_t)n, st));
  uint32_t h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, mx.p, 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  ps.reset();
  bits_owner = bits_for((uint64_t)owner_base + n_owner);  // owners are < owner_base + n_owner
  bits_pos = bits_for(h);
  SHN_CHECK(bits_owner + bits_pos <= 64, "self-join: contig count / length exceed the 64-bit event budget");
}

void SelfJoin::join(uint32_t lo, uint32_t hi, const uint8_t* d_status, PairTable* out) {
  out->n = 0;
  if (n == 0) return;
  cudaStream_t st = c->stream;
  auto ent_index = [&](uint32_t owner) -> uint64_t {
    if (owner <= owner_base) return 0;
    return std::min<uint64_t>((uint64_t)owner - owner_base, n_owner);
  };
  const uint64_t g_lo = h_ent_off[ent_index(lo)], g_hi = h_ent_off[ent_index(hi)];
  if (g_hi <= g_lo) return;
  const uint64_t m = g_hi - g_lo;
  const std::string n_emit = tag + "_emit_events";
  // live partners in front of the block: accepted ones only (none without a status array: then the
  // caller joins everything at once, lo = 0)
  SHN_CHECK(d_status != nullptr || lo <= owner_base, "self-join: a block join needs the status array");
  const bool use_acc = d_status != nullptr && lo > owner_base;
  std::unique_ptr<ProfScope> ps(new ProfScope(c, n_emit.c_str(), use_acc ? 7 : 4));
  DevBuf acc_flag, acc_cnt, acc_q;
  size_t tb = 0;
  if (use_acc) {
    acc_flag.reserve((n + 1) * 4);
    acc_cnt.reserve((n + 1) * 4);
    acc_flag_kernel<<<shn_grid(n + 1, kBlock), kBlock, 0, st>>>(owner_s.as<uint32_t>(), n, lo, d_status,
                                                               acc_flag.as<uint32_t>());
    KERNEL_CHECK();
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, acc_flag.as<uint32_t>(), acc_cnt.as<uint32_t>(),
                                             (int64_t)(n + 1), st));
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, acc_flag.as<uint32_t>(), acc_cnt.as<uint32_t>(),
                                             (int64_t)(n + 1), st));
    acc_q.reserve(n * 4);  // upper bound without a host round trip
    acc_list_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(acc_flag.as<uint32_t>(), acc_cnt.as<uint32_t>(), n,
                                                           acc_q.as<uint32_t>());
    KERNEL_CHECK();
  }
  DevBuf cnt, lb, ev_off;
  cnt.reserve((m + 1) * 4);
  lb.reserve((m + 1) * 4);
  ev_off.reserve((m + 1) * 8);
  count_events_kernel<<<shn_grid(m + 1, kBlock), kBlock, 0, st>>>(
      range_g.as<uint2>(), owner_s.as<uint32_t>(), use_acc ? acc_cnt.as<uint32_t>() : nullptr, g_lo, m, lo > owner_base ? lo : 0u, cnt.as<uint32_t>(),
      lb.as<uint32_t>());
  KERNEL_CHECK();
  {
    cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> it(cnt.as<uint32_t>(), CastU64());
    tb = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, ev_off.as<uint64_t>(), (int64_t)(m + 1), st));
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, it, ev_off.as<uint64_t>(), (int64_t)(m + 1), st));
  }
  uint64_t n_events = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_events, ev_off.as<uint64_t>() + m, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (n_events == 0) return;
  SHN_CHECK(n_events < 0x7FFFFFFFull, "self-join: more than 2^31-1 match events in one block");
  const uint64_t owner_end = (uint64_t)owner_base + n_owner;  // exclusive
  const uint64_t j_span = std::min<uint64_t>(hi, owner_end) - std::min<uint64_t>(lo, owner_end);
  const int jrel_bits = bits_for(j_span ? j_span - 1 : 0);
  const bool narrow = jrel_bits + bits_pos <= 32;
  DevBuf ev_d, ev_v;
  ev_d.reserve(n_events * 4);
  ev_v.reserve(n_events * (narrow ? 4 : 8));
  if (narrow) {
    emit_events_kernel<uint32_t><<<shn_grid(m, kBlock), kBlock, 0, st>>>(
        ev_off.as<uint64_t>(), range_g.as<uint2>(), lb.as<uint32_t>(),
        owner_s.as<uint32_t>(), use_acc ? acc_cnt.as<uint32_t>() : nullptr, acc_q.as<uint32_t>(),
        owner_g.as<uint32_t>(), pos_g.as<uint32_t>(), g_lo, m, lo, bits_pos, ev_d.as<uint32_t>(),
        ev_v.as<uint32_t>());
  } else {
    emit_events_kernel<uint64_t><<<shn_grid(m, kBlock), kBlock, 0, st>>>(
        ev_off.as<uint64_t>(), range_g.as<uint2>(), lb.as<uint32_t>(),
        owner_s.as<uint32_t>(), use_acc ? acc_cnt.as<uint32_t>() : nullptr, acc_q.as<uint32_t>(),
        owner_g.as<uint32_t>(), pos_g.as<uint32_t>(), g_lo, m, lo, bits_pos, ev_d.as<uint32_t>(),
        ev_v.as<uint64_t>());
  }
  KERNEL_CHECK();
  ps.reset();
  cnt.release();
  lb.release();
  ev_off.release();
  acc_flag.release();
  acc_cnt.release();
  acc_q.release();
  if (narrow)
    sort_reduce<uint32_t>(this, n_events, lo, jrel_bits, ev_d, ev_v, out);
  else
    sort_reduce<uint64_t>(this, n_events, lo, jrel_bits, ev_d, ev_v, out);
  if (getenv("SHN_HOST_TRACE"))
    fprintf(stderr, "[self-join %s] owners [%u,%u): %llu of %llu entries, %llu events (%d-bit values), %llu pairs\n",
            tag.c_str(), lo, hi, (unsigned long long)m, (unsigned long long)n, (unsigned long long)n_events,
            narrow ? 32 : 64, (unsigned long long)out->n);
}
