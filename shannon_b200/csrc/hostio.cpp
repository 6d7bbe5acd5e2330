// Host-side text IO of the front end (native; replaces the reference's per-line Python loops
// around the hot path: load_kmers' parsing, extension_correction.py:209-216, the FASTA reader of
// kmers_for_component.py:330-339/369-380 and the per-component writers :351,396-397,457-476).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <thread>

#include "common.cuh"

namespace {

struct MappedFile {
  const char* p = nullptr;
  size_t n = 0;
  int fd = -1;
  explicit MappedFile(const char* path) {
    fd = open(path, O_RDONLY);
    if (fd < 0) SHN_FAIL(std::string("cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) != 0) {
      close(fd);
      SHN_FAIL(std::string("cannot stat ") + path);
    }
    n = (size_t)st.st_size;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) {
        close(fd);
        SHN_FAIL(std::string("cannot mmap ") + path);
      }
      p = (const char*)m;
      madvise(m, n, MADV_SEQUENTIAL);
    }
  }
  ~MappedFile() {
    if (p) munmap((void*)p, n);
    if (fd >= 0) close(fd);
  }
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

unsigned n_threads_for(size_t bytes) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  unsigned want = (unsigned)std::min<size_t>(hw, bytes / (4u << 20) + 1);
  return std::max(1u, std::min(want, 64u));
}

// one `KMER count` line -> (key, count); returns an error string or nullptr
const char* parse_kmer_line(const char* b, const char* e, int* k1, uint64_t* key, uint32_t* cnt) {
  // key: 1 word for k-mers of <= 32 bases, 2 words (low word first) for 33 bases
  while (b < e && is_ws(*b)) ++b;
  const char* t0 = b;
  while (b < e && !is_ws(*b)) ++b;
  int klen = (int)(b - t0);
  if (klen == 0) return "empty line in k-mer file";
  if (klen > 33) return "k-mer longer than 33 bases (K > 32 is not supported)";
  if (*k1 == 0) *k1 = klen;
  if (klen != *k1) return "k-mers of different lengths in one file";
  u128 x = 0;
  for (const char* q = t0; q < b; ++q) {
    uint32_t code = shn_code_of((uint8_t)*q);
    if (code > 3) return "k-mer contains a character outside ACGT";
    x = (x << 2) | (u128)code;
  }
  while (b < e && is_ws(*b)) ++b;
  const char* t1 = b;
  while (b < e && !is_ws(*b)) ++b;
  if (b == t1) return "k-mer line without a count";
  const char* t1e = b;
  while (b < e && is_ws(*b)) ++b;
  if (b != e) return "k-mer line with more than two fields";
  uint64_t v = 0;
  bool plain = true;
  for (const char* q = t1; q < t1e; ++q) {
    if (*q < '0' || *q > '9') {
      plain = false;
      break;
    }
    v = v * 10 + (uint64_t)(*q - '0');
    if (v > 0x7FFFFFFFull) return "k-mer count above 2^31-1";
  }
  if (!plain) {  // float() accepts 3.0, 1e3, ...; only integral values are meaningful counts
    std::string tok(t1, t1e);
    char* endp = nullptr;
    double d = strtod(tok.c_str(), &endp);
    if (endp == tok.c_str() || *endp != 0 || !(d >= 0) || d > 2147483647.0 || d != std::floor(d))
      return "k-mer count is not a non-negative integer";
    v = (uint64_t)d;
  }
  key[0] = (uint64_t)x;
  if (klen > 32) key[1] = (uint64_t)(x >> 64);
  *cnt = (uint32_t)v;
  return nullptr;
}

}  // namespace

void shn_parse_kmer_file_impl(const char* path, uint64_t** keys_out, uint32_t** counts_out,
                              uint64_t* n_out, int* k1_out) {
  MappedFile f(path);
  *keys_out = nullptr;
  *counts_out = nullptr;
  *n_out = 0;
  *k1_out = 0;
  if (f.n == 0) SHN_FAIL(std::string("k-mer file is empty: ") + path);
  unsigned nt = n_threads_for(f.n);
  // chunk boundaries on line starts
  std::vector<size_t> cut(nt + 1);
  cut[0] = 0;
  cut[nt] = f.n;
  for (unsigned t = 1; t < nt; ++t) {
    size_t p = f.n / nt * t;
    while (p < f.n && f.p[p] != '\n') ++p;
    cut[t] = std::min(f.n, p + 1);
  }
  std::vector<uint64_t> lines(nt, 0);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        uint64_t c = 0;
        const char* p = f.p + cut[t];
        const char* e = f.p + cut[t + 1];
        while (p < e) {
          const char* nl = (const char*)memchr(p, '\n', e - p);
          ++c;
          if (!nl) break;
          p = nl + 1;
        }
        lines[t] = c;
      });
    for (auto& x : th) x.join();
  }
  std::vector<uint64_t> first(nt + 1, 0);
  for (unsigned t = 0; t < nt; ++t) first[t + 1] = first[t] + lines[t];
  uint64_t n = first[nt];
  // k1 from the very first line so that every thread checks against the same length
  int k1 = 0;
  {
    const char* e = (const char*)memchr(f.p, '\n', f.n);
    if (!e) e = f.p + f.n;
    const char* b = f.p;
    while (b < e && is_ws(*b)) ++b;
    const char* t0 = b;
    while (b < e && !is_ws(*b)) ++b;
    k1 = (int)(b - t0);
  }
  const uint64_t kw = k1 > 32 ? 2 : 1;
  uint64_t* keys = (uint64_t*)malloc(std::max<uint64_t>(n, 1) * 8 * kw);
  uint32_t* counts = (uint32_t*)malloc(std::max<uint64_t>(n, 1) * 4);
  if (!keys || !counts) {
    free(keys);
    free(counts);
    SHN_FAIL("out of host memory parsing the k-mer file");
  }
  std::vector<std::string> errs(nt);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        uint64_t i = first[t];
        const char* p = f.p + cut[t];
        const char* e = f.p + cut[t + 1];
        int kk = k1;
        while (p < e) {
          const char* nl = (const char*)memchr(p, '\n', e - p);
          const char* le = nl ? nl : e;
          const char* err = parse_kmer_line(p, le, &kk, &keys[i * kw], &counts[i]);
          if (err) {
            errs[t] = std::string(err) + " (line " + std::to_string(i + 1) + ")";
            return;
          }
          ++i;
          if (!nl) break;
          p = nl + 1;
        }
      });
    for (auto& x : th) x.join();
  }
  for (auto& e : errs)
    if (!e.empty()) {
      free(keys);
      free(counts);
      SHN_FAIL(std::string(path) + ": " + e);
    }
  *keys_out = keys;
  *counts_out = counts;
  *n_out = n;
  *k1_out = k1;
}

// FASTA reader with the reference's exact quirks (kmers_for_component.py:333-339):
//   name = readline()[:-1]; stop if empty;  read = readline()[:-1];  a record with an empty
//   read is kept and ends the input.  n_fixed >= 0: read exactly n_fixed records without the
//   stop rule (mate 2 is read in lock-step with mate 1), padding with empty reads.
void shn_load_fasta_impl(const char* path, int64_t n_fixed, char** bases_out, uint64_t** offs_out,
                         uint64_t* n_out) {
  MappedFile f(path);
  std::vector<uint64_t> offs;
  offs.push_back(0);
  std::string bases;
  bases.reserve(f.n / 2 + 16);
  const char* p = f.p;
  const char* e = f.p + f.n;
  auto readline_strip = [&](const char** b, const char** le) {
    // Python: f.readline()[:-1]  (drops the last character of the line, newline or not)
    if (p >= e) {
      *b = *le = e;
      return;
    }
    const char* nl = (const char*)memchr(p, '\n', e - p);
    const char* line_end = nl ? nl + 1 : e;  // one past the line including '\n'
    *b = p;
    *le = line_end - 1;
    p = line_end;
  };
  uint64_t n = 0;
  while (n_fixed < 0 || (int64_t)n < n_fixed) {
    const char *nb, *ne, *rb, *re;
    readline_strip(&nb, &ne);
    if (n_fixed < 0 && nb == ne) break;
    readline_strip(&rb, &re);
    bases.append(rb, re - rb);
    offs.push_back(bases.size());
    ++n;
    if (n_fixed < 0 && rb == re) break;
  }
  char* bo = (char*)malloc(std::max<size_t>(bases.size(), 1));
  uint64_t* oo = (uint64_t*)malloc(offs.size() * 8);
  if (!bo || !oo) {
    free(bo);
    free(oo);
    SHN_FAIL("out of host memory loading reads");
  }
  memcpy(bo, bases.data(), bases.size());
  memcpy(oo, offs.data(), offs.size() * 8);
  *bases_out = bo;
  *offs_out = oo;
  *n_out = n;
}

namespace {
struct OutFile {
  FILE* f;
  OutFile(const char* path, int append) {
    f = fopen(path, append ? "ab" : "wb");
    if (!f) SHN_FAIL(std::string("cannot open for writing: ") + path);
    setvbuf(f, nullptr, _IOFBF, 1 << 22);
  }
  ~OutFile() {
    if (f) fclose(f);
  }
  void write(const std::string& s) {
    if (!s.empty() && fwrite(s.data(), 1, s.size(), f) != s.size()) SHN_FAIL("short write");
  }
};
}  // namespace

void shn_write_fasta_subset_impl(const char* path, int append, const char* bases,
                                 const uint64_t* offsets, const uint32_t* read_idx, uint64_t m,
                                 uint64_t first_index, const char* suffix) {
  OutFile out(path, append);
  std::string buf;
  buf.reserve(1 << 22);
  char num[32];
  size_t slen = strlen(suffix);
  for (uint64_t e = 0; e < m; ++e) {
    uint64_t r = read_idx[e];
    int nl = snprintf(num, sizeof(num), ">%llu", (unsigned long long)(first_index + e));
    buf.append(num, nl);
    buf.append(suffix, slen);
    buf.push_back('\n');
    buf.append(bases + offsets[r], offsets[r + 1] - offsets[r]);
    buf.push_back('\n');
    if (buf.size() > (1u << 22) - 4096) {
      out.write(buf);
      buf.clear();
    }
  }
  out.write(buf);
}

// component{comp}k1mers_allowed.dict (kmers_for_component.py:457-476): for each listed contig, for
// each K1-mer window, `K1MER\tweight\n`.  weights are laid out per contig at win_off[contig].
void shn_write_k1mer_windows_impl(const char* path, const char* bases, const uint64_t* offsets,
                                  const uint32_t* contig_ids, uint64_t m, int k1,
                                  const uint32_t* weights, const uint64_t* win_off) {
  OutFile out(path, 0);
  std::string buf;
  buf.reserve(1 << 22);
  char num[16];
  for (uint64_t e = 0; e < m; ++e) {
    uint32_t c = contig_ids[e];
    uint64_t b = offsets[c], len = offsets[c + 1] - b;
    if (len < (uint64_t)k1) continue;
    const uint32_t* w = weights + win_off[c];
    for (uint64_t p = 0; p + k1 <= len; ++p) {
      buf.append(bases + b + p, k1);
      int nl = snprintf(num, sizeof(num), "\t%u\n", w[p]);
      buf.append(num, nl);
      if (buf.size() > (1u << 22) - 4096) {
        out.write(buf);
        buf.clear();
      }
    }
  }
  out.write(buf);
}

// k1mer.dict_org writer (`jellyfish dump -c -t` format, shannon.py:441): `KMER\tcount\n` per entry,
// formatted by all host threads in batches.  keys: 1 or 2 words per key (low word first).
void shn_write_kmer_file_impl(const char* path, const uint64_t* keys, const uint32_t* counts, uint64_t n,
                              int k1) {
  if (k1 < 1 || k1 > 33) SHN_FAIL("k1 must be in 1..33");
  OutFile out(path, 0);
  const uint64_t kw = k1 > 32 ? 2 : 1;
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(nt, 32u));
  const uint64_t batch = 1ull << 22;  // lines per thread and round
  std::vector<std::string> bufs(nt);
  for (uint64_t lo = 0; lo < n; lo += batch * nt) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        std::string& b = bufs[t];
        b.clear();
        const uint64_t a = std::min(n, lo + batch * t), e = std::min(n, a + batch);
        b.reserve((e - a) * (k1 + 8));
        char line[64];
        for (uint64_t i = a; i < e; ++i) {
          u128 x = kw == 2 ? (((u128)keys[2 * i + 1] << 64) | keys[2 * i]) : (u128)keys[i];
          for (int j = k1 - 1; j >= 0; --j) {
            line[j] = shn_base_of((uint32_t)x & 3u);
            x >>= 2;
          }
          int m = k1;
          line[m++] = '\t';
          char num[12];
          int d = 0;
          uint32_t v = counts[i];
          do {
            num[d++] = (char)('0' + v % 10);
            v /= 10;
          } while (v);
          while (d) line[m++] = num[--d];
          line[m++] = '\n';
          b.append(line, m);
        }
      });
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < nt; ++t) out.write(bufs[t]);
  }
}

// FASTA as rc_s.py reads it (rc_s.py:7-19): blank lines are dropped, a line whose first field starts
// with '>' is a header (kept stripped), any other line is one sequence = its first whitespace-
// delimited field.  The driver's files have one sequence line per header; a header without a
// sequence line or a sequence line without a header is refused (rc_s.py would still copy it, but
// nothing downstream reads such files).  Returns names (without '>') and sequences, concatenated.
void shn_load_fasta_named_impl(const char* path, char** names_out, uint64_t** name_offs_out, char** bases_out,
                               uint64_t** offs_out, uint64_t* n_out) {
  MappedFile f(path);
  std::string names, bases;
  std::vector<uint64_t> noffs(1, 0), offs(1, 0);
  bases.reserve(f.n / 2 + 16);
  const char* p = f.p;
  const char* e = f.p + f.n;
  bool have_header = false;
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', e - p);
    const char* le = nl ? nl : e;
    const char* b = p;
    p = nl ? nl + 1 : e;
    while (b < le && (is_ws(*b) || *b == '\n')) ++b;
    while (le > b && (is_ws(le[-1]) || le[-1] == '\n')) --le;
    if (b == le) continue;  // blank line
    if (*b == '>') {
      if (have_header) SHN_FAIL(std::string(path) + ": header line without a sequence line");
      names.append(b + 1, le - b - 1);
      noffs.push_back(names.size());
      have_header = true;
    } else {
      if (!have_header) SHN_FAIL(std::string(path) + ": sequence line without a header (multi-line FASTA records are not supported)");
      const char* t = b;
      while (t < le && !is_ws(*t)) ++t;  // fields[0]
      bases.append(b, t - b);
      offs.push_back(bases.size());
      have_header = false;
    }
  }
  if (have_header) SHN_FAIL(std::string(path) + ": header line without a sequence line");
  const uint64_t n = offs.size() - 1;
  char* no = (char*)malloc(std::max<size_t>(names.size(), 1));
  char* bo = (char*)malloc(std::max<size_t>(bases.size(), 1));
  uint64_t* nf = (uint64_t*)malloc(noffs.size() * 8);
  uint64_t* of = (uint64_t*)malloc(offs.size() * 8);
  if (!no || !bo || !nf || !of) {
    free(no); free(bo); free(nf); free(of);
    SHN_FAIL("out of host memory loading reads");
  }
  memcpy(no, names.data(), names.size());
  memcpy(bo, bases.data(), bases.size());
  memcpy(nf, noffs.data(), noffs.size() * 8);
  memcpy(of, offs.data(), offs.size() * 8);
  *names_out = no;
  *name_offs_out = nf;
  *bases_out = bo;
  *offs_out = of;
  *n_out = n;
}

// `>name\nSEQ\n` per record (what rc_s.py writes: '\n'.join(lines) + '\n')
void shn_write_fasta_named_impl(const char* path, int append, const char* names, const uint64_t* name_offs,
                                const char* bases, const uint64_t* offs, uint64_t n) {
  OutFile out(path, append);
  std::string buf;
  buf.reserve(1 << 22);
  for (uint64_t i = 0; i < n; ++i) {
    buf.push_back('>');
    buf.append(names + name_offs[i], name_offs[i + 1] - name_offs[i]);
    buf.push_back('\n');
    buf.append(bases + offs[i], offs[i + 1] - offs[i]);
    buf.push_back('\n');
    if (buf.size() > (1u << 22) - 65536) {
      out.write(buf);
      buf.clear();
    }
  }
  out.write(buf);
}
