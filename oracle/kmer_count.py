"""TEST INFRASTRUCTURE ONLY -- deterministic stand-in for `jellyfish count -m K+1` +
`jellyfish dump -c -t -L 1` (shannon.py:439-441; the jellyfish binary is absent here and no
version is pinned by the reference, so K1-mer *line order* is unpinned -- SURVEY 8c).

Counts every (K+1)-mer window of every sequence line of the given FASTA files, skipping
windows that contain a non-ACGT character (what jellyfish does with N), not canonical (the
reference never passes -C, shannon.py:427,436).  Lines are emitted `KMER<TAB>count` in
ascending ASCII order of the k-mer, which is the documented order the GPU counter
(shn_count_k1mers) reproduces.
"""
import sys


def count_k1mers(fasta_paths, k1):
    counts = {}
    ok = set("ACGT")
    for path in fasta_paths:
        with open(path) as f:
            for line in f:
                if line.startswith(">"):
                    continue
                s = line.strip()
                bad = -1            # index of the last non-ACGT char seen so far
                for i, ch in enumerate(s):
                    if ch not in ok:
                        bad = i
                    st = i - k1 + 1
                    if st >= 0 and bad < st:
                        km = s[st:i + 1]
                        counts[km] = counts.get(km, 0) + 1
    return counts


def count_k1mers_numpy(fasta_paths, k1):
    """Same result as count_k1mers (checked in tests/test_cpu_host_logic.py), vectorised with numpy
    for the large fixtures (k1 <= 32): returns (kmers as an (n, k1) uint8 ASCII matrix in ascending
    ASCII order, counts)."""
    import numpy as np
    assert k1 <= 32
    code = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):          # ASCII order = integer order of the packed word
        code[ch] = i
    by_len = {}
    for path in fasta_paths:
        with open(path, "rb") as f:
            for line in f:
                if line.startswith(b">"):
                    continue
                s = line.strip()
                if len(s) >= k1:
                    by_len.setdefault(len(s), []).append(s)
    chunks = []
    for L, seqs in by_len.items():
        a = code[np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), L)]
        nw = L - k1 + 1
        key = np.zeros((len(seqs), nw), dtype=np.uint64)
        bad = np.zeros((len(seqs), nw), dtype=bool)
        for j in range(k1):
            col = a[:, j:j + nw]
            bad |= col > 3
            key = (key << np.uint64(2)) | (col & 3).astype(np.uint64)
        chunks.append(key[~bad])
    allk = np.concatenate(chunks) if chunks else np.empty(0, dtype=np.uint64)
    keys, counts = np.unique(allk, return_counts=True)
    shifts = (2 * (k1 - 1 - np.arange(k1))).astype(np.uint64)
    mat = np.frombuffer(b"ACGT", dtype=np.uint8)[((keys[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)]
    return mat, counts


def write_dict_numpy(mat, counts, out_path):
    k1 = mat.shape[1] if len(mat) else 0
    txt = mat.tobytes().decode()
    with open(out_path, "w") as f:
        f.write("".join("%s\t%d\n" % (txt[i * k1:(i + 1) * k1], c) for i, c in enumerate(counts.tolist())))


def write_dict(counts, out_path, min_count=1):
    with open(out_path, "w") as f:
        for km in sorted(counts):
            if counts[km] >= min_count:
                f.write("%s\t%d\n" % (km, counts[km]))


if __name__ == "__main__":
    k1 = int(sys.argv[1])
    write_dict(count_k1mers(sys.argv[3:], k1), sys.argv[2])
