// 2-bit packing of reads (a11 input): ASCII bases + offsets -> 32 bases per uint64 word, MSB first,
// ceil(len/32) words per read, plus (word offset, length | bad-flag).  A read containing any
// character outside "ACGT" is flagged (read.strip('ACTG'), kmers_for_component.py:336,376).
// Independent of the K1-mer key width: compiled once.
#include <cub/cub.cuh>

#include "reads.cuh"

namespace {
constexpr int kBlock = 256;
constexpr uint32_t kLenBad = 0x80000000u;

__global__ void __launch_bounds__(kBlock)
    read_words_kernel(const uint64_t* __restrict__ offs, uint64_t n, uint64_t* __restrict__ nwords) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t len = offs[i + 1] - offs[i];
  nwords[i] = (len + 31) >> 5;
}

// Eight ASCII bases (first base in the lowest byte of w) -> 16 bits, first base most significant;
// *bad is set if any byte is not one of "ACGT" (upper case only).  SWAR over the 64-bit word:
// with x = bits 2..1 of a byte, A,C,T,G = 0,1,2,3 and the code (A0 G1 C2 T3) is
// hi = bit1 ^ bit2, lo = bit2; the byte is valid iff it equals 'A' + 6*lo + 2*hi + 11*(hi&lo).
__device__ __forceinline__ uint64_t pack8(uint64_t w, bool* bad) {
  const uint64_t ones = 0x0101010101010101ull;
  const uint64_t b1 = (w >> 1) & ones, b2 = (w >> 2) & ones;
  const uint64_t hi = b1 ^ b2, lo = b2;
  const uint64_t expect = 0x4141414141414141ull + 6ull * lo + 2ull * hi + 11ull * (hi & lo);
  *bad |= expect != w;
  uint64_t c = (hi << 1) | lo;
  // byte 0 (first base) to the top, then squeeze the eight 2-bit codes together
  c = ((uint64_t)__byte_perm((uint32_t)c, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(c >> 32), 0, 0x0123);
  c = (c | (c >> 6)) & 0x000F000F000F000Full;
  c = (c | (c >> 12)) & 0x000000FF000000FFull;
  c = (c | (c >> 24)) & 0xFFFFull;
  return c;
}

// 8 lanes per read, one 32-base word per lane and iteration.  The 32 bytes of a word are fetched
// with five aligned 8-byte loads and funnel shifts (reads start at arbitrary byte offsets); the
// aligned loads never leave the allocation (256-byte granular), bytes past the read count as 'A'.
__global__ void __launch_bounds__(kBlock)
    pack_reads_kernel(const char* __restrict__ bases, const uint64_t* __restrict__ offs,
                      const uint64_t* __restrict__ woff, uint64_t n, uint64_t* __restrict__ words,
                      uint32_t* __restrict__ len_out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = t >> 3;
  int lane8 = (int)(t & 7);
  bool active = i < n;
  uint64_t start = 0, len = 0, wbase = 0;
  if (active) {
    start = __ldg(&offs[i]);
    len = __ldg(&offs[i + 1]) - start;
    wbase = __ldg(&woff[i]);
  }
  uint64_t nw = (len + 31) >> 5;
  bool bad = false;
  for (uint64_t w = lane8; w < nw; w += 8) {
    const int cnt = (int)min((uint64_t)32, len - 32 * w);
    const uint64_t addr = (uint64_t)(uintptr_t)bases + start + 32 * w;
    const uint64_t* q = reinterpret_cast<const uint64_t*>(addr & ~7ull);
    const int off = (int)(addr & 7);
    uint64_t v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = (8 * k < off + cnt) ? __ldg(q + k) : 0ull;
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t g = off ? ((v[k] >> (8 * off)) | (v[k + 1] << (64 - 8 * off))) : v[k];
      const int nb = min(max(cnt - 8 * k, 0), 8);
      const uint64_t m = nb == 8 ? ~0ull : ((1ull << (8 * nb)) - 1ull);
      g = (g & m) | (0x4141414141414141ull & ~m);  // past the end: 'A' = code 0 = left-aligned word
      x = (x << 16) | pack8(g, &bad);
    }
    words[wbase + w] = x;
  }
  // OR the bad flags of the 8 lanes of this read
  unsigned b = __ballot_sync(0xFFFFFFFFu, bad);
  unsigned grp = (b >> ((threadIdx.x & 31) & ~7)) & 0xFFu;
  if (active && lane8 == 0) {
    len_out[i] = (uint32_t)len | (grp ? kLenBad : 0u);
  }
}


void reads_state_free(shn_ctx* c) {
  ReadsState* s = static_cast<ReadsState*>(c->reads);
  if (s)
    for (int m = 0; m < 2; ++m)
      if (s->up_done[m]) cudaEventDestroy(s->up_done[m]);
  delete s;
  c->reads = nullptr;
}
}  // namespace

ReadsState* shn_reads_of(shn_ctx* c) {
  if (!c->reads) {
    c->reads = new ReadsState();
    c->reads_free = &reads_state_free;
  }
  return static_cast<ReadsState*>(c->reads);
}

void shn_reads_load(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets, uint64_t n,
                    int on_device) {
  SHN_CHECK(mate == 0 || mate == 1, "mate must be 0 or 1");
  ReadsState* s = shn_reads_of(c);
  PackedReads& pr = s->reads[mate];
  pr.n = n;
  pr.n_words = 0;
  if (n == 0) return;
  SHN_CHECK(n < 0xFFFFFFFFull, "at most 2^32-1 read records per call");
  uint64_t total = 0;
  const uint64_t* d_offs;
  if (on_device) {
    d_offs = offsets;
    CUDA_CHECK(cudaMemcpyAsync(&total, offsets + n, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  } else {
    total = offsets[n];
    d_offs = (const uint64_t*)InputView::get(c, offsets, (n + 1) * 8, 0, s->stage_b);
  }
  const char* d_bases = (const char*)InputView::get(c, bases, total, on_device, s->stage_a);
  DevBuf nwords;
  nwords.reserve((n + 1) * 8);
  pr.woff.reserve((n + 1) * 8);
  pr.len.reserve(n * 4);
  CUDA_CHECK(cudaMemsetAsync(nwords.p, 0, (n + 1) * 8, c->stream));
  {
    ProfScope ps(c, "read_word_offsets", 2);
    read_words_kernel<<<shn_grid(n, kBlock), kBlock, 0, c->stream>>>(d_offs, n, nwords.as<uint64_t>());
    KERNEL_CHECK();
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, nwords.as<uint64_t>(), pr.woff.as<uint64_t>(),
                                             (int64_t)(n + 1), c->stream));
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->tmp(tb), tb, nwords.as<uint64_t>(),
                                             pr.woff.as<uint64_t>(), (int64_t)(n + 1), c->stream));
  }
  CUDA_CHECK(cudaMemcpyAsync(&pr.n_words, pr.woff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost,
                             c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  pr.words.reserve((pr.n_words + 2) * 8);  // +2: extract_kmer may touch words past the last read
  CUDA_CHECK(cudaMemsetAsync(pr.words.as<uint64_t>() + pr.n_words, 0, 16, c->stream));
  {
    ProfScope ps(c, "pack_reads");
    pack_reads_kernel<<<shn_grid(n * 8, kBlock), kBlock, 0, c->stream>>>(
        d_bases, d_offs, pr.woff.as<uint64_t>(), n, pr.words.as<uint64_t>(), pr.len.as<uint32_t>());
    KERNEL_CHECK();
  }
  // no synchronisation here: the copies of the inputs completed before the size round trip above,
  // and the packing itself runs under whatever the host does next
}

// Starts the host->device copy of one mate file on the context's second stream and returns
// immediately; shn_l4_load_reads_staged() later waits for it and packs.  With pinned host memory
// the copy overlaps whatever the main stream does meanwhile (the whole L3 stage in the pipeline).
void shn_reads_upload_async(shn_ctx* c, int mate, const char* bases, const uint64_t* offsets, uint64_t n) {
  SHN_CHECK(mate == 0 || mate == 1, "mate must be 0 or 1");
  ReadsState* s = shn_reads_of(c);
  if (!c->stream2) CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  if (!s->up_done[mate]) CUDA_CHECK(cudaEventCreateWithFlags(&s->up_done[mate], cudaEventDisableTiming));
  s->up_n[mate] = n;
  const uint64_t total = n ? offsets[n] : 0;
  s->up_bases[mate].reserve(std::max<uint64_t>(total, 1));
  s->up_offs[mate].reserve((n + 1) * 8);
  // the staging buffers may have been used by kernels still queued on the main stream
  cudaEvent_t ev = c->prof_event();
  CUDA_CHECK(cudaEventRecord(ev, c->stream));
  CUDA_CHECK(cudaStreamWaitEvent(c->stream2, ev, 0));
  c->prof_pool.push_back(ev);
  if (total)
    CUDA_CHECK(cudaMemcpyAsync(s->up_bases[mate].p, bases, total, cudaMemcpyHostToDevice, c->stream2));
  CUDA_CHECK(cudaMemcpyAsync(s->up_offs[mate].p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream2));
  CUDA_CHECK(cudaEventRecord(s->up_done[mate], c->stream2));
}

void shn_reads_load_staged(shn_ctx* c, int mate) {
  SHN_CHECK(mate == 0 || mate == 1, "mate must be 0 or 1");
  ReadsState* s = shn_reads_of(c);
  SHN_CHECK(s->up_done[mate] != nullptr, "no upload in flight for this mate");
  CUDA_CHECK(cudaStreamWaitEvent(c->stream, s->up_done[mate], 0));
  shn_reads_load(c, mate, s->up_bases[mate].as<char>(), s->up_offs[mate].as<uint64_t>(), s->up_n[mate], 1);
}

