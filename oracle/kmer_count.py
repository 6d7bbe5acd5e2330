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


def write_dict(counts, out_path, min_count=1):
    with open(out_path, "w") as f:
        for km in sorted(counts):
            if counts[km] >= min_count:
                f.write("%s\t%d\n" % (km, counts[km]))


if __name__ == "__main__":
    k1 = int(sys.argv[1])
    write_dict(count_k1mers(sys.argv[3:], k1), sys.argv[2])
