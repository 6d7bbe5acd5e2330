"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's faster_reps.py (SURVEY.md 8f row
f4): removal of transcripts whose first and last r-mer (r = 24) both occur in another, longer (or
equally long, earlier-named) transcript at the right distance.  Insertion-ordered dict semantics
(CPython >= 3.7), like the rest of the oracle.  Pinned to the real faster_reps.py in
tests/test_oracle_vs_reference.py."""

R = 24            # faster_reps.py:9
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def reverse_complement(s):
    """faster_reps.py:15-22 (upper-case ACGT; other characters pass through unchanged there)."""
    return "".join(_COMP.get(b, b) for b in reversed(s))


def read_contigs(infile):
    """faster_reps.py:99-116: name = first token of the header without '>', sequence = stripped
    line; returns (contigs dict, rmer index rmer -> [[name, pos], ...] in file order)."""
    contigs = {}
    index = {}
    name = None
    with open(infile) as f:
        for line in f:
            if line[0] == ">":
                name = line.strip().split()[0][1:]
                continue
            seq = line.strip()
            contigs[name] = seq
            for i in range(len(seq) - R + 1):
                index.setdefault(seq[i:i + R], []).append([name, i])
    return contigs, index


def duplicate_check_ends(contigs, index, name, rc):
    """faster_reps.py:60-93."""
    contig = contigs[name]
    if rc:
        contig = reverse_complement(contig)
    first, last = contig[:R], contig[-R:]
    if first in index and last in index:
        hits = {}
        for c, p in index[first]:
            if c == name:
                continue
            if c in hits:
                hits[c][0] = p
            else:
                hits[c] = [p, -1]
        for c, p in index[last]:
            if c == name:
                continue
            if c in hits:
                hits[c][1] = p
            else:
                hits[c] = [-1, p]
        for c, (pf, pl) in hits.items():
            if pf >= 0 and pl >= 0 and abs((pl - pf) - (len(contig) - R)) < 3:
                if len(contig) < len(contigs[c]) or (len(contig) == len(contigs[c]) and name > c):
                    return True
    return False


def find_reps(infile, outfile, ds):
    """faster_reps.py:99-131."""
    contigs, index = read_contigs(infile)
    out = []
    for name in contigs:
        dup = duplicate_check_ends(contigs, index, name, False)
        if ds:
            dup = dup or duplicate_check_ends(contigs, index, name, True)
        if not dup:
            out.append(">" + name + "\n")
            out.append(contigs[name] + "\n")
    with open(outfile, "w") as f:
        f.writelines(out)
