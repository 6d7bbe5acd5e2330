"""Seeded synthetic RNA-Seq read generator for BASELINE.json configs 3-5.

Transcripts are built from *genes with shared exons* (4-10 random exons of 100-400 bp, 2-4
isoforms per gene, each exon kept with probability 0.75) so that contigs of one gene overlap
and form multi-contig components.  Read pairs are 300 bp fragments: mate 1 = first 100 bp,
mate 2 = reverse complement of the last 100 bp, i.i.d. 1 % substitution errors, no N.

All per-pair randomness is a counter-based hash (``mix64``) of (seed, pair index, lane), so
the CUDA generator ``shn_synth_pairs`` (csrc/synth.cu) and :func:`make_pairs` below produce
the same bytes; tests use the numpy version at small sizes, bench.py the CUDA one at 10 M pairs.
"""
import numpy as np

U64 = np.uint64
_M1 = U64(0xFF51AFD7ED558CCD)
_M2 = U64(0xC4CEB9FE1A85EC53)
_GOLD = U64(0x9E3779B97F4A7C15)

ASCII = np.frombuffer(b"AGCT", dtype=np.uint8)          # code -> base (A=0,G=1,C=2,T=3)


def mix64(x):
    """murmur3 fmix64 on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=U64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> U64(33)
        x *= _M1
        x ^= x >> U64(33)
        x *= _M2
        x ^= x >> U64(33)
    return x


def stream(seed, index, lane):
    """64 random bits for (seed, index, lane)."""
    with np.errstate(over="ignore"):
        a = mix64(U64(seed) * _GOLD + np.asarray(index, dtype=U64))
        return mix64(a ^ (np.asarray(lane, dtype=U64) * _GOLD))


def make_transcripts(n_transcripts, seed=1234):
    """Returns a list of 2-bit-code uint8 arrays (A=0,G=1,C=2,T=3), each >= 300 bp."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n_transcripts:
        n_exons = int(rng.integers(4, 11))
        exons = [rng.integers(0, 4, size=int(rng.integers(100, 401)), dtype=np.uint8)
                 for _ in range(n_exons)]
        for _ in range(int(rng.integers(2, 5))):
            keep = rng.random(n_exons) < 0.75
            if not keep.any():
                continue
            t = np.concatenate([e for e, k in zip(exons, keep) if k])
            if len(t) >= 300 and len(out) < n_transcripts:
                out.append(t)
    return out


def expression_thresholds(n_transcripts, lengths, skewed=False):
    """Cumulative pick thresholds (uint64, 63-bit scale) for choosing a transcript per pair:
    probability proportional to expression x length; expression uniform or Zipf(1)."""
    expr = 1.0 / np.arange(1, n_transcripts + 1) if skewed else np.ones(n_transcripts)
    w = expr * np.asarray(lengths, dtype=np.float64)
    cdf = np.cumsum(w) / w.sum()
    thr = np.minimum(cdf * float(1 << 63), float((1 << 63) - 1024)).astype(U64)
    thr[-1] = U64((1 << 63) - 1)
    return thr


def pack_transcripts(transcripts):
    offs = np.zeros(len(transcripts) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(t) for t in transcripts])
    return np.concatenate(transcripts).astype(np.uint8), offs


ERR_THRESHOLD_24 = 167772          # round(0.01 * 2**24)


def make_pairs(tx_codes, tx_offs, thr, n_pairs, seed, first_pair=0, read_len=100, frag_len=300,
               err_threshold=ERR_THRESHOLD_24):
    """numpy twin of csrc/synth.cu.  Returns two (n_pairs, read_len) uint8 ASCII arrays."""
    p = np.arange(first_pair, first_pair + n_pairs, dtype=U64)
    r_t = stream(seed, p, 0) >> U64(1)
    t = np.searchsorted(thr, r_t, side="right").astype(np.int64)
    t = np.minimum(t, len(thr) - 1)
    tlen = (tx_offs[t + 1] - tx_offs[t]).astype(U64)
    start = stream(seed, p, 1) % (tlen - U64(frag_len) + U64(1))
    base = (tx_offs[t].astype(U64) + start).astype(np.int64)
    j = np.arange(read_len, dtype=np.int64)
    m1 = tx_codes[base[:, None] + j[None, :]]
    m2 = 3 - tx_codes[base[:, None] + (frag_len - 1 - j)[None, :]]
    mates = []
    for m, codes in enumerate((m1, m2)):
        lane = U64(2 + m * read_len) + j.astype(U64)
        e = stream(seed, p[:, None], lane[None, :])
        is_err = (e & U64(0xFFFFFF)) < U64(err_threshold)
        shift = ((e >> U64(24)) % U64(3)).astype(np.uint8) + np.uint8(1)
        codes = np.where(is_err, (codes + shift) & np.uint8(3), codes).astype(np.uint8)
        mates.append(ASCII[codes])
    return mates[0], mates[1]


def write_fasta(path, reads, names=None):
    """reads: (n, L) uint8 ASCII array or list of str."""
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            s = r if isinstance(r, str) else bytes(r).decode()
            f.write(">%s\n%s\n" % (names[i] if names is not None else i, s))


_RC_TABLE = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _RC_TABLE[_a] = _b


def revcomp_ascii(reads):
    """(n, L) uint8 ASCII -> reverse complement per row."""
    return _RC_TABLE[reads[:, ::-1]]


def rc_double(m1, m2=None):
    """The driver's RC doubling (shannon.py:395-424) for fixed-length read arrays:
    SE: reads = [R ; rc(R)];  PE ds: reads_1 = [R1 ; rc(R2)], reads_2 = [rc(R1) ; R2]."""
    if m2 is None:
        return np.concatenate([m1, revcomp_ascii(m1)])
    return (np.concatenate([m1, revcomp_ascii(m2)]),
            np.concatenate([revcomp_ascii(m1), m2]))
