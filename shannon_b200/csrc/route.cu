// Routing primitives of the hash-sharded K1-mer table: owner rank of a key and the
// stable partition of a batch by owner, feeding an NCCL all-to-all (torch.distributed) whose
// send buffers must be contiguous per destination rank.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {
constexpr int kBlock = 256;

// Independent of the bucket hash: the table uses the HIGH bits of fmix64(key) (mulhi), the owner
// uses the LOW 32 bits, so the keys of one shard still spread over all of that shard's buckets.
__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t nranks) {
  uint64_t lo = shn_mix64(key) & 0xFFFFFFFFull;
  return (uint32_t)((lo * nranks) >> 32);
}

__global__ void __launch_bounds__(kBlock)
    owner_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t nranks,
                 uint32_t* __restrict__ owner, uint32_t* __restrict__ iota) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  owner[i] = owner_of(keys[i], nranks);
  iota[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    rank_offsets_kernel(const uint32_t* __restrict__ owner_sorted, uint64_t n, uint32_t nranks,
                        unsigned long long* __restrict__ first) {
  // first[r] = first index with owner >= r, r in 0..nranks
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t prev = i == 0 ? 0u : owner_sorted[i - 1] + 1u;
  uint32_t cur = i == n ? nranks + 1u : owner_sorted[i] + 1u;
  for (uint32_t r = prev; r < cur && r <= nranks; ++r) first[r] = i;
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
    gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ perm, uint64_t n,
                  T* __restrict__ dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
    scatter_kernel(const T* __restrict__ src, const uint32_t* __restrict__ perm, uint64_t n,
                   T* __restrict__ dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}
}  // namespace

// perm: stable partition of 0..n-1 by owner rank; counts[r] = number of keys owned by rank r.
void shn_route_plan_impl(shn_ctx* c, const uint64_t* d_keys, uint64_t n, uint32_t nranks,
                         uint32_t* d_perm, uint64_t* h_counts) {
  SHN_CHECK(nranks >= 1 && nranks <= 1024, "nranks out of range");
  SHN_CHECK(n < 0xFFFFFFFFull, "at most 2^32-1 keys per routing batch");
  for (uint32_t r = 0; r < nranks; ++r) h_counts[r] = 0;
  if (n == 0) return;
  DevBuf owner, owner_s, iota, first;
  owner.reserve(n * 4);
  owner_s.reserve(n * 4);
  iota.reserve(n * 4);
  first.reserve((nranks + 2) * 8);
  ProfScope ps(c, "route_plan", 3);
  owner_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(d_keys, n, nranks, owner.as<uint32_t>(),
                                                             iota.as<uint32_t>());
  KERNEL_CHECK();
  int bits = 1;
  while ((1u << bits) < nranks) ++bits;
  size_t tb = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, owner.as<uint32_t>(), owner_s.as<uint32_t>(),
                                             iota.as<uint32_t>(), d_perm, (int64_t)n, 0, bits, c->stream));
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->tmp(tb), tb, owner.as<uint32_t>(), owner_s.as<uint32_t>(),
                                             iota.as<uint32_t>(), d_perm, (int64_t)n, 0, bits, c->stream));
  rank_offsets_kernel<<<shn_grid(n + 1, kBlock), kBlock, 0, c->stream>>>(
      owner_s.as<uint32_t>(), n, nranks, first.as<unsigned long long>());
  KERNEL_CHECK();
  std::vector<unsigned long long> h(nranks + 1);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), first.p, (nranks + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (uint32_t r = 0; r < nranks; ++r) h_counts[r] = h[r + 1] - h[r];
}

void shn_permute_impl(shn_ctx* c, const void* src, const uint32_t* d_perm, uint64_t n, int elem_bytes,
                      int scatter, void* dst) {
  if (n == 0) return;
  ProfScope ps(c, scatter ? "route_scatter" : "route_gather");
  unsigned g = shn_grid(n, kBlock);
  if (elem_bytes == 8) {
    if (scatter)
      scatter_kernel<uint64_t><<<g, kBlock, 0, c->stream>>>((const uint64_t*)src, d_perm, n, (uint64_t*)dst);
    else
      gather_kernel<uint64_t><<<g, kBlock, 0, c->stream>>>((const uint64_t*)src, d_perm, n, (uint64_t*)dst);
  } else if (elem_bytes == 4) {
    if (scatter)
      scatter_kernel<uint32_t><<<g, kBlock, 0, c->stream>>>((const uint32_t*)src, d_perm, n, (uint32_t*)dst);
    else
      gather_kernel<uint32_t><<<g, kBlock, 0, c->stream>>>((const uint32_t*)src, d_perm, n, (uint32_t*)dst);
  } else if (elem_bytes == 1) {
    if (scatter)
      scatter_kernel<uint8_t><<<g, kBlock, 0, c->stream>>>((const uint8_t*)src, d_perm, n, (uint8_t*)dst);
    else
      gather_kernel<uint8_t><<<g, kBlock, 0, c->stream>>>((const uint8_t*)src, d_perm, n, (uint8_t*)dst);
  } else {
    SHN_FAIL("element size must be 1, 4 or 8 bytes");
  }
  KERNEL_CHECK();
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
