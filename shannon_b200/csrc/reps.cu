// f4: faster_reps.py:60-131 -- a transcript is dropped when its first and its last
// r-mer (r = 24) both occur in ONE other transcript at a distance that matches its own length
// (+-2) and that other transcript is longer (or equally long with a smaller name).  The
// reference keeps an inverted list r-mer -> [(contig, pos)] of every r-mer of every transcript;
// here the list is a sorted array of (r-mer, entry) pairs -- the same r-mer multimap primitive as
// the duplicate filter of the hot path (selfjoin.cu) -- and every transcript end is two binary
// searches plus a merge of two short (contig, pos)-ordered ranges.
// Independent of the K1-mer key width: compiled once.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {
constexpr int kBlock = 256;
constexpr int kR = 24;  // faster_reps.py:9

__device__ __forceinline__ uint64_t seg_of(const uint64_t* __restrict__ offs, uint64_t n, uint64_t g) {
  uint64_t lo = 0, hi = n;  // last s with offs[s] <= g
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&offs[mid]) <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBlock)
    rep_counts_kernel(const uint64_t* __restrict__ offs, uint64_t n, uint64_t* __restrict__ cnt) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n) return;
  uint64_t v = 0;
  if (c < n) {
    const uint64_t len = offs[c + 1] - offs[c];
    v = len >= kR ? len - kR + 1 : 0;
  }
  cnt[c] = v;
}

// one thread per base position: the r-mer starting there (entry index = ent_off[contig] + position)
__global__ void __launch_bounds__(kBlock)
    rep_entries_kernel(const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                       const uint64_t* __restrict__ ent_off, uint64_t n, uint64_t total,
                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* bad) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const uint64_t c = seg_of(offs, n, g);
  const uint64_t b = __ldg(&offs[c]), e = __ldg(&offs[c + 1]);
  uint32_t code = shn_code_of_strict((uint8_t)__ldg(&bases[g]));
  if (code > 3) atomicAdd(bad, 1ull);
  if (g + kR > e) return;
  uint64_t key = 0;
  for (int j = 0; j < kR; ++j) key = (key << 2) | (uint64_t)(shn_code_of_strict((uint8_t)__ldg(&bases[g + j])) & 3u);
  const uint64_t idx = __ldg(&ent_off[c]) + (g - b);
  keys[idx] = key;
  vals[idx] = (uint32_t)idx;
}

__device__ __forceinline__ uint64_t lower_bound(const uint64_t* __restrict__ a, uint64_t n, uint64_t x) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(&a[mid]) < x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// duplicate_check_ends (faster_reps.py:60-93) for contig q, one strand per thread
__global__ void __launch_bounds__(kBlock)
    rep_check_kernel(const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                     const uint64_t* __restrict__ ent_off, const uint32_t* __restrict__ name_rank, uint64_t n,
                     const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ svals, uint64_t n_ent,
                     int ds, uint8_t* __restrict__ dup) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t q = t >> 1;
  const int rc = (int)(t & 1);
  if (q >= n || (rc && !ds)) return;
  const uint64_t b = __ldg(&offs[q]), len = __ldg(&offs[q + 1]) - b;
  if (len < kR) return;  // contig[:r] is the whole (short) contig: never in the index
  uint64_t head = 0, tail = 0;
  for (int j = 0; j < kR; ++j) {
    head = (head << 2) | (uint64_t)(shn_code_of_strict((uint8_t)__ldg(&bases[b + j])) & 3u);
    tail = (tail << 2) | (uint64_t)(shn_code_of_strict((uint8_t)__ldg(&bases[b + len - kR + j])) & 3u);
  }
  uint64_t first = head, last = tail;
  if (rc) {  // first r-mer of the reverse complement = rc(last r-mer), and vice versa
    first = shn_revcomp(tail, kR);
    last = shn_revcomp(head, kR);
  }
  uint64_t f0 = lower_bound(skeys, n_ent, first), f1 = lower_bound(skeys, n_ent, first + 1);
  uint64_t l0 = lower_bound(skeys, n_ent, last), l1 = lower_bound(skeys, n_ent, last + 1);
  if (f0 == f1 || l0 == l1) return;
  const uint32_t my_rank = __ldg(&name_rank[q]);
  // both ranges are ordered by (contig, position): merge them contig by contig; the reference keeps
  // the LAST list entry of a contig, i.e. its largest position
  uint64_t a = f0, z = l0;
  bool found = false;
  while (a < f1 && z < l1 && !found) {
    const uint64_t ea = __ldg(&svals[a]), ez = __ldg(&svals[z]);
    const uint64_t ca = seg_of(ent_off, n, ea), cz = seg_of(ent_off, n, ez);
    if (ca != cz) {
      // skip the whole group of the smaller contig
      const uint64_t lim = __ldg(&ent_off[(ca < cz ? ca : cz) + 1]);
      if (ca < cz) {
        while (a < f1 && __ldg(&svals[a]) < lim) ++a;
      } else {
        while (z < l1 && __ldg(&svals[z]) < lim) ++z;
      }
      continue;
    }
    const uint64_t lim = __ldg(&ent_off[ca + 1]), base = __ldg(&ent_off[ca]);
    uint64_t pf = 0, pl = 0;
    while (a < f1 && __ldg(&svals[a]) < lim) pf = __ldg(&svals[a++]) - base;
    while (z < l1 && __ldg(&svals[z]) < lim) pl = __ldg(&svals[z++]) - base;
    if (ca == q) continue;
    const long long diff = (long long)pl - (long long)pf - (long long)(len - kR);
    if (diff > -3 && diff < 3) {
      const uint64_t clen = __ldg(&offs[ca + 1]) - __ldg(&offs[ca]);
      if (len < clen || (len == clen && my_rank > __ldg(&name_rank[ca]))) found = true;
    }
  }
  if (found) dup[q] = 1;
}
}  // namespace

// dup_out[c] = 1 iff the reference would drop transcript c; name_rank[c] = rank of its name in
// string order (the tie-break of equally long transcripts); all pointers are host pointers
void shn_find_reps_impl(shn_ctx* c, const char* bases, const uint64_t* offsets, const uint32_t* name_rank,
                        uint64_t n, int ds, uint8_t* dup_out) {
  if (n == 0) return;
  cudaStream_t st = c->stream;
  const uint64_t total = offsets[n];
  memset(dup_out, 0, n);
  if (total == 0) return;
  DevBuf d_bases, d_offs, d_rank, cnt, ent_off, d_dup;
  d_bases.reserve(total);
  d_offs.reserve((n + 1) * 8);
  d_rank.reserve(n * 4);
  cnt.reserve((n + 1) * 8);
  ent_off.reserve((n + 1) * 8);
  d_dup.reserve(n);
  CUDA_CHECK(cudaMemcpyAsync(d_bases.p, bases, total, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(d_offs.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(d_rank.p, name_rank, n * 4, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemsetAsync(d_dup.p, 0, n, st));
  rep_counts_kernel<<<shn_grid(n + 1, kBlock), kBlock, 0, st>>>(d_offs.as<uint64_t>(), n, cnt.as<uint64_t>());
  KERNEL_CHECK();
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<uint64_t>(), ent_off.as<uint64_t>(), (int64_t)(n + 1), st));
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, cnt.as<uint64_t>(), ent_off.as<uint64_t>(),
                                           (int64_t)(n + 1), st));
  uint64_t n_ent = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_ent, ent_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  SHN_CHECK(n_ent < 0xFFFFFFFFull, "more than 2^32-1 r-mers");
  c->counters.reserve(64 * sizeof(unsigned long long));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  CUDA_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
  DevBuf keys, vals, skeys, svals;
  keys.reserve(std::max<uint64_t>(n_ent, 1) * 8);
  vals.reserve(std::max<uint64_t>(n_ent, 1) * 4);
  skeys.reserve(std::max<uint64_t>(n_ent, 1) * 8);
  svals.reserve(std::max<uint64_t>(n_ent, 1) * 4);
  {
    ProfScope ps(c, "reps_entries");
    rep_entries_kernel<<<shn_grid(total, kBlock), kBlock, 0, st>>>(
        d_bases.as<char>(), d_offs.as<uint64_t>(), ent_off.as<uint64_t>(), n, total, keys.as<uint64_t>(),
        vals.as<uint32_t>(), ctr);
    KERNEL_CHECK();
  }
  unsigned long long bad = 0;
  CUDA_CHECK(cudaMemcpyAsync(&bad, ctr, 8, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  SHN_CHECK(bad == 0, "transcript contains a character outside ACGT");
  if (n_ent) {
    {
      ProfScope ps(c, "reps_sort");
      tb = 0;
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint64_t>(), skeys.as<uint64_t>(),
                                                 vals.as<uint32_t>(), svals.as<uint32_t>(), (int64_t)n_ent, 0,
                                                 2 * kR, st));
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, keys.as<uint64_t>(), skeys.as<uint64_t>(),
                                                 vals.as<uint32_t>(), svals.as<uint32_t>(), (int64_t)n_ent, 0,
                                                 2 * kR, st));
    }
    ProfScope ps(c, "reps_check");
    rep_check_kernel<<<shn_grid(2 * n, kBlock), kBlock, 0, st>>>(
        d_bases.as<char>(), d_offs.as<uint64_t>(), ent_off.as<uint64_t>(), d_rank.as<uint32_t>(), n,
        skeys.as<uint64_t>(), svals.as<uint32_t>(), n_ent, ds, d_dup.as<uint8_t>());
    KERNEL_CHECK();
  }
  CUDA_CHECK(cudaMemcpyAsync(dup_out, d_dup.p, n, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
}
