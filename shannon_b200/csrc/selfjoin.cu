// See selfjoin.cuh.  Sort-based, race-free, deterministic.
#include <cub/cub.cuh>

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
                        const uint32_t* __restrict__ owner, const uint32_t* __restrict__ pos,
                        uint64_t n, uint32_t* __restrict__ owner_s, uint32_t* __restrict__ pos_s,
                        uint32_t* __restrict__ run_head, uint32_t* __restrict__ grp_head) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t g = idx_s[p];
  uint32_t o = owner[g];
  owner_s[p] = o;
  pos_s[p] = pos[g];
  bool hk = p == 0 || keys_s[p] != keys_s[p - 1];
  bool hg = hk || owner[idx_s[p - 1]] != o;
  run_head[p] = hk ? (uint32_t)p : 0u;
  grp_head[p] = hg ? (uint32_t)p : 0u;
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// number of qualifying partners of sorted entry p (0 for entries whose owner is outside [lo,hi))
__global__ void __launch_bounds__(kBlock)
    count_partners_kernel(const uint32_t* __restrict__ owner_s, const uint32_t* __restrict__ run_start,
                          const uint32_t* __restrict__ grp_start, uint64_t n, uint32_t lo, uint32_t hi,
                          const uint8_t* __restrict__ status, uint64_t* __restrict__ cnt) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  uint64_t c = 0;
  if (p < n) {
    uint32_t o = owner_s[p];
    if (o >= lo && o < hi) {
      uint32_t a = run_start[p], b = grp_start[p];
      if (!status) {
        c = b - a;
      } else {
        for (uint32_t q = a; q < b; ++q) {
          uint32_t d = owner_s[q];
          c += (d >= lo || status[d] == 1) ? 1 : 0;
        }
      }
    }
  }
  cnt[p] = c;
}

// one thread per sorted entry p: its events (j = owner(p), d = owner(q), i = pos(p))
__global__ void __launch_bounds__(kBlock)
    emit_events_kernel(const uint64_t* __restrict__ ev_off, const uint32_t* __restrict__ run_start,
                       const uint32_t* __restrict__ grp_start, const uint32_t* __restrict__ owner_s,
                       const uint32_t* __restrict__ pos_s, uint64_t n, uint32_t lo, uint32_t hi,
                       const uint8_t* __restrict__ status, int bits_owner, int bits_pos,
                       uint64_t* __restrict__ ev) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint64_t e = ev_off[p];
  if (ev_off[p + 1] == e) return;
  const uint64_t j = owner_s[p], i = pos_s[p];
  const uint64_t head = (j << (bits_owner + bits_pos)) | i;
  for (uint32_t q = run_start[p], b = grp_start[p]; q < b; ++q) {
    uint32_t d = owner_s[q];
    if (!status || d >= lo || status[d] == 1) ev[e++] = head | ((uint64_t)d << bits_pos);
  }
}

struct EvVal {
  uint32_t count, min_i, max_i, covered;
};
struct EvReduce {
  __device__ __forceinline__ EvVal operator()(const EvVal& a, const EvVal& b) const {
    EvVal r;
    r.count = a.count + b.count;
    r.min_i = a.min_i < b.min_i ? a.min_i : b.min_i;
    r.max_i = a.max_i > b.max_i ? a.max_i : b.max_i;
    r.covered = a.covered + b.covered;
    return r;
  }
};
struct EvKeyOf {
  const uint64_t* ev;
  int bits_pos;
  __device__ __forceinline__ uint64_t operator()(uint64_t e) const { return ev[e] >> bits_pos; }
};
struct EvValOf {
  const uint64_t* ev;
  uint64_t n;
  int bits_pos;
  uint32_t r;
  __device__ __forceinline__ EvVal operator()(uint64_t e) const {
    uint64_t x = ev[e];
    uint64_t pm = (1ull << bits_pos) - 1ull;
    uint32_t i = (uint32_t)(x & pm);
    uint32_t c = r;  // last event of its (j,d) segment contributes a full interval
    if (e + 1 < n) {
      uint64_t y = ev[e + 1];
      if ((y >> bits_pos) == (x >> bits_pos)) {
        uint32_t gap = (uint32_t)(y & pm) - i;
        c = gap < r ? gap : r;
      }
    }
    EvVal v;
    v.count = 1;
    v.min_i = i;
    v.max_i = i;
    v.covered = c;
    return v;
  }
};

__global__ void __launch_bounds__(kBlock)
    split_pairs_kernel(const uint64_t* __restrict__ ukeys, const EvVal* __restrict__ agg, uint64_t n,
                       int bits_owner, uint32_t* hi, uint32_t* lo, uint32_t* count, uint32_t* min_i,
                       uint32_t* max_i, uint32_t* covered) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t k = ukeys[t];
  hi[t] = (uint32_t)(k >> bits_owner);
  lo[t] = (uint32_t)(k & ((1ull << bits_owner) - 1ull));
  EvVal v = agg[t];
  count[t] = v.count;
  min_i[t] = v.min_i;
  max_i[t] = v.max_i;
  covered[t] = v.covered;
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) ++b;
  return b;
}

}  // namespace

void SelfJoin::prepare(shn_ctx* ctx, const char* tag_, const uint64_t* d_keys, const uint32_t* d_owner,
                       const uint32_t* d_pos, uint64_t n_, int key_bits, uint32_t r_) {
  c = ctx;
  tag = tag_;
  n = n_;
  r = r_;
  if (n == 0) return;
  SHN_CHECK(n < 0xFFFFFFFFull, "self-join: more than 2^32-1 entries");
  cudaStream_t st = c->stream;
  const std::string n_sort = tag + "_sort_keys", n_heads = tag + "_runs";
  DevBuf keys_s, idx, idx_s, run_head, grp_head;
  keys_s.reserve(n * 8);
  idx.reserve(n * 4);
  idx_s.reserve(n * 4);
  std::unique_ptr<ProfScope> ps(new ProfScope(c, n_sort.c_str(), 2));
  iota_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(idx.as<uint32_t>(), n);
  KERNEL_CHECK();
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, keys_s.as<uint64_t>(),
                                             idx.as<uint32_t>(), idx_s.as<uint32_t>(), (int64_t)n, 0,
                                             key_bits, st));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, d_keys, keys_s.as<uint64_t>(),
                                             idx.as<uint32_t>(), idx_s.as<uint32_t>(), (int64_t)n, 0,
                                             key_bits, st));
  ps.reset();
  ps.reset(new ProfScope(c, n_heads.c_str(), 6));
  owner_s.reserve(n * 4);
  pos_s.reserve(n * 4);
  run_head.reserve(n * 4);
  grp_head.reserve(n * 4);
  run_start.reserve(n * 4);
  grp_start.reserve(n * 4);
  gather_heads_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(
      keys_s.as<uint64_t>(), idx_s.as<uint32_t>(), d_owner, d_pos, n, owner_s.as<uint32_t>(),
      pos_s.as<uint32_t>(), run_head.as<uint32_t>(), grp_head.as<uint32_t>());
  KERNEL_CHECK();
  tb = 0;
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, tb, run_head.as<uint32_t>(),
                                            run_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(c->tmp(tb), tb, run_head.as<uint32_t>(),
                                            run_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceScan::InclusiveScan(c->tmp(tb), tb, grp_head.as<uint32_t>(),
                                            grp_start.as<uint32_t>(), MaxOp(), (int64_t)n, st));
  // bit budget of the composite event key (j | d | i)
  DevBuf mx;
  mx.reserve(8);
  tb = 0;
  CUDA_CHECK(cub::DeviceReduce::Max(nullptr, tb, d_owner, mx.as<uint32_t>(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceReduce::Max(c->tmp(tb), tb, d_owner, mx.as<uint32_t>(), (int64_t)n, st));
  CUDA_CHECK(cub::DeviceReduce::Max(c->tmp(tb), tb, d_pos, mx.as<uint32_t>() + 1, (int64_t)n, st));
  uint32_t h[2];
  CUDA_CHECK(cudaMemcpyAsync(h, mx.p, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  bits_owner = bits_for(h[0]);
  bits_pos = bits_for(h[1]);
  SHN_CHECK(2 * bits_owner + bits_pos <= 64,
            "self-join: contig count / length exceed the 64-bit event key budget");
}

void SelfJoin::join(uint32_t lo, uint32_t hi, const uint8_t* d_status, PairTable* out) {
  out->n = 0;
  if (n == 0) return;
  cudaStream_t st = c->stream;
  const std::string n_emit = tag + "_emit_events", n_sort2 = tag + "_sort_events",
                    n_reduce = tag + "_reduce_pairs";
  const int bo = bits_owner, bp = bits_pos;
  DevBuf cnt, ev_off;
  cnt.reserve((n + 1) * 8);
  ev_off.reserve((n + 1) * 8);
  std::unique_ptr<ProfScope> ps(new ProfScope(c, n_emit.c_str(), 3));
  count_partners_kernel<<<shn_grid(n + 1, kBlock), kBlock, 0, st>>>(
      owner_s.as<uint32_t>(), run_start.as<uint32_t>(), grp_start.as<uint32_t>(), n, lo, hi, d_status,
      cnt.as<uint64_t>());
  KERNEL_CHECK();
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<uint64_t>(), ev_off.as<uint64_t>(),
                                           (int64_t)(n + 1), st));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, cnt.as<uint64_t>(), ev_off.as<uint64_t>(),
                                           (int64_t)(n + 1), st));
  uint64_t n_events = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_events, ev_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (n_events == 0) return;
  SHN_CHECK(n_events < 0x7FFFFFFFull, "self-join: more than 2^31-1 match events in one block");
  DevBuf ev, ev_s;
  ev.reserve(n_events * 8);
  ev_s.reserve(n_events * 8);
  emit_events_kernel<<<shn_grid(n, kBlock), kBlock, 0, st>>>(
      ev_off.as<uint64_t>(), run_start.as<uint32_t>(), grp_start.as<uint32_t>(), owner_s.as<uint32_t>(),
      pos_s.as<uint32_t>(), n, lo, hi, d_status, bo, bp, ev.as<uint64_t>());
  KERNEL_CHECK();
  ps.reset();
  ps.reset(new ProfScope(c, n_sort2.c_str(), 1));
  tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb, ev.as<uint64_t>(), ev_s.as<uint64_t>(),
                                            (int64_t)n_events, 0, 2 * bo + bp, st));
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(c->tmp(tb), tb, ev.as<uint64_t>(), ev_s.as<uint64_t>(),
                                            (int64_t)n_events, 0, 2 * bo + bp, st));
  ev.release();
  ps.reset();
  ps.reset(new ProfScope(c, n_reduce.c_str(), 2));
  // segmented reduction per (j, d)
  DevBuf ukeys, agg, nruns;
  ukeys.reserve(n_events * 8);
  agg.reserve(n_events * sizeof(EvVal));
  nruns.reserve(8);
  cub::CountingInputIterator<uint64_t> cit(0);
  cub::TransformInputIterator<uint64_t, EvKeyOf, cub::CountingInputIterator<uint64_t>> kin(
      cit, EvKeyOf{ev_s.as<uint64_t>(), bp});
  cub::TransformInputIterator<EvVal, EvValOf, cub::CountingInputIterator<uint64_t>> vin(
      cit, EvValOf{ev_s.as<uint64_t>(), n_events, bp, r});
  tb = 0;
  CUDA_CHECK(cub::DeviceReduce::ReduceByKey(nullptr, tb, kin, ukeys.as<uint64_t>(), vin,
                                            agg.as<EvVal>(), nruns.as<uint64_t>(), EvReduce(),
                                            (int)n_events, st));
  CUDA_CHECK(cub::DeviceReduce::ReduceByKey(c->tmp(tb), tb, kin, ukeys.as<uint64_t>(), vin,
                                            agg.as<EvVal>(), nruns.as<uint64_t>(), EvReduce(),
                                            (int)n_events, st));
  uint64_t n_pairs = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_pairs, nruns.p, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  out->hi.reserve(n_pairs * 4);
  out->lo.reserve(n_pairs * 4);
  out->count.reserve(n_pairs * 4);
  out->min_i.reserve(n_pairs * 4);
  out->max_i.reserve(n_pairs * 4);
  out->covered.reserve(n_pairs * 4);
  split_pairs_kernel<<<shn_grid(n_pairs, kBlock), kBlock, 0, st>>>(
      ukeys.as<uint64_t>(), agg.as<EvVal>(), n_pairs, bo, out->hi.as<uint32_t>(),
      out->lo.as<uint32_t>(), out->count.as<uint32_t>(), out->min_i.as<uint32_t>(),
      out->max_i.as<uint32_t>(), out->covered.as<uint32_t>());
  KERNEL_CHECK();
  ps.reset();
  CUDA_CHECK(cudaStreamSynchronize(st));
  out->n = n_pairs;
}
