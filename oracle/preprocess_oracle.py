"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the driver's read pre-processing in front of the
hot path: rc_s.py's per-line reverse complement and the file concatenations of
shannon.py:395-424.  Pinned to the real rc_s.py in tests/test_oracle_vs_reference.py.  K1-mer
counting (row f1) is oracle/kmer_count.py."""
import os
import shutil

_D = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def reverse_complement_file(infile, outfile):
    """rc_s.py:7-19: blank lines dropped; header lines kept stripped; every other line replaced by
    the reverse complement of its first field (KeyError on characters outside ACGTN)."""
    out = []
    with open(infile) as f:
        for line in f:
            if not line.strip():
                continue
            fields = line.strip().split()
            if fields[0][0] == ">":
                out.append(line.strip())
            else:
                out.append("".join(_D[b] for b in reversed(fields[0])))
    with open(outfile, "w") as f:
        f.write("\n".join(out) + "\n")


def _cat(parts, dst):
    with open(dst, "wb") as out:
        for p in parts:
            with open(p, "rb") as f:
                shutil.copyfileobj(f, out)


def rc_double(reads_files, kmer_directory, paired_end, double_stranded):
    """shannon.py:395-424 -> the reads_files the rest of the driver uses."""
    d = kmer_directory
    if not paired_end:
        if not double_stranded:
            return list(reads_files)
        reverse_complement_file(reads_files[0], d + "/rc.fasta")
        _cat([reads_files[0], d + "/rc.fasta"], d + "/reads.fasta")
        os.remove(d + "/rc.fasta")
        return [d + "/reads.fasta"]
    if not double_stranded:
        reverse_complement_file(reads_files[1], d + "/rc_2.fasta")
        return [reads_files[0], d + "/rc_2.fasta"]
    reverse_complement_file(reads_files[0], d + "/rc_1.fasta")
    reverse_complement_file(reads_files[1], d + "/rc_2.fasta")
    _cat([reads_files[0], d + "/rc_2.fasta"], d + "/reads_1.fasta")
    _cat([d + "/rc_1.fasta", reads_files[1]], d + "/reads_2.fasta")
    os.remove(d + "/rc_1.fasta")
    os.remove(d + "/rc_2.fasta")
    return [d + "/reads_1.fasta", d + "/reads_2.fasta"]
