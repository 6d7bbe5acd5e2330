"""The golden cases: how each input is (re)built and with which driver parameters it is run.
Shared by make_golden.py (reference side, build container) and the tests (oracle / GPU side)."""
import gzip
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


def sample_seqs(name, limit=None):
    with gzip.open(os.path.join(HERE, "samples", name + ".seqs.gz"), "rt") as f:
        seqs = f.read().split("\n")[:-1]
    return seqs[:limit] if limit else seqs


def repeat_rich_reads(seed, n_reads, read_len, genome_len, n_genomes):
    """Low-entropy genomes with copied segments: forces weight ties, repeated r-mers/C-mers
    and self-overlapping contigs (the tie-break paths of duplicate_check and of the DFS)."""
    rnd = random.Random(seed)
    genomes = []
    for _ in range(n_genomes):
        g = [rnd.choice("ACGT") for _ in range(genome_len)]
        for _ in range(3):
            a = rnd.randrange(0, genome_len - 40)
            b = rnd.randrange(0, genome_len - 40)
            g[b:b + 30] = g[a:a + 30]
        genomes.append("".join(g))
    for g in list(genomes):
        a = rnd.randrange(20, len(g) // 2)
        genomes.append(g[:a] + g[a + 25:])
    reads = []
    for _ in range(n_reads):
        g = rnd.choice(genomes)
        s = rnd.randrange(0, len(g) - read_len + 1)
        r = list(g[s:s + read_len])
        if rnd.random() < 0.3:
            r[rnd.randrange(read_len)] = rnd.choice("ACGT")
        reads.append("".join(r))
    return reads


def degenerate_reads(kind):
    """Inputs on which (almost) nothing survives: the front end must still leave the reference's
    (empty) files behind."""
    rnd = random.Random(5)

    def rand(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))
    if kind == "low_cov":                 # every K1-mer has weight 1-2: no seed
        return [rand(100) for _ in range(40)]
    if kind == "lowcomplexity":           # homopolymer reads: (nearly) every K1-mer is filtered
        return ["A" * 100] * 30 + ["T" * 60 + "A" * 40] * 10
    if kind == "short_reads":             # walks shorter than min_length
        return [rand(30) for _ in range(200)] * 4
    if kind == "one_read":                # one contig, a single component of size 1
        return [rand(100)] * 50
    if kind == "with_N":                  # every read is dropped by the read partition
        return [rand(50) + "N" + rand(49) for _ in range(20)] * 5
    raise ValueError(kind)


CASES = {
    # BASELINE.json config 1 (bundled single-end sample; FASTA twin of the fastq, no Quorum)
    "sample_se": {"K": 24, "input": ("sample_se",), "run": {}},
    # BASELINE.json config 1 as literally written: the sequence lines of Samples/SE_reads.fastq
    # (= PE_read_1.fastq ++ PE_read_2.fastq, 5 000 x 100 bp, many N; checked by make_golden.py
    # against the fastq itself) in single-end mode, no Quorum (absent: SURVEY 8c)
    "sample_se_fastq": {"K": 24, "input": ("sample_se_fastq",), "run": {}},
    # BASELINE.json config 2 (bundled paired-end sample, double-stranded RC doubling)
    "sample_pe": {"K": 24, "input": ("sample_pe",), "run": {}},
    "synth_pe_default": {"K": 24, "input": ("synth", 12, 1500, 1), "run": {}},
    "synth_pe_metis2": {"K": 24, "input": ("synth", 30, 3000, 2), "run": {"partition_size": 2}},
    "synth_pe_metis1": {"K": 24, "input": ("synth", 40, 6000, 3), "run": {"partition_size": 1}},
    "synth_se_k16_inmem": {"K": 16, "input": ("synth_se", 10, 1500, 7),
                           "run": {"inMem": True, "partition_size": 2, "min_length": 50}},
    "synth_pe_dflag": {"K": 24, "input": ("synth", 8, 800, 11), "rc_double": False,
                       "run": {"double_stranded_load": True, "ec_inMem": False}},
    # key-width boundaries: K1 = 32 fills one 64-bit word, K1 = 33 (shannon.py's largest -K, 32)
    # needs two words per key
    "synth_pe_k31": {"K": 31, "input": ("synth", 10, 1200, 17), "run": {"partition_size": 2}},
    "synth_pe_k32": {"K": 32, "input": ("synth", 12, 1500, 13), "run": {}},
    "synth_se_k32_inmem": {"K": 32, "input": ("synth_se", 10, 1500, 19),
                           "run": {"inMem": True, "partition_size": 1, "min_length": 60}},
    # raw K1-mer components of 10^5..10^6 nodes: the default speculative walk tiers (>= 60 000
    # nodes) fire without any environment override.  Stored as digests (sha256 per output).
    "synth_pe_big": {"K": 24, "input": ("synth", 10, 80000, 23), "run": {}, "digest": True,
                     "fast_count": True},
    "repeat_rich_k32": {"K": 32, "input": ("repeat", 9, 500, 60, 220, 4),
                        "run": {"min_weight": 2, "min_length": 40, "partition_size": 2}},
}
for _k in ("low_cov", "lowcomplexity", "short_reads", "one_read", "with_N"):
    CASES["degenerate_" + _k] = {"K": 24, "input": ("degenerate", _k), "run": {}}
for _s in range(6):
    CASES["repeat_rich_%d" % _s] = {
        "K": [8, 10, 12, 15][_s % 4], "input": ("repeat", _s, 400, 40, 160, 4),
        "run": {"min_weight": 2, "min_length": 20 + _s, "partition_size": 1 + _s % 3,
                "inMem": bool(_s & 1)}}


def case_inputs(spec):
    import helpers
    kind = spec["input"][0]
    if kind == "sample_se":
        return sample_seqs("SE_read"), None
    if kind == "sample_se_fastq":
        return sample_seqs("PE_read_1") + sample_seqs("PE_read_2"), None
    if kind == "sample_pe":
        return sample_seqs("PE_read_1"), sample_seqs("PE_read_2")
    if kind == "synth":
        _, ntx, npairs, seed = spec["input"]
        return helpers.synthetic_seqs(ntx, npairs, seed)
    if kind == "synth_se":
        _, ntx, npairs, seed = spec["input"]
        return helpers.synthetic_seqs(ntx, npairs, seed)[0], None
    if kind == "repeat":
        return repeat_rich_reads(*spec["input"][1:]), None
    if kind == "degenerate":
        return degenerate_reads(spec["input"][1]), None
    raise ValueError(kind)


def digest(data):
    import hashlib
    if isinstance(data, str):
        data = data.encode()
    return hashlib.sha256(data).hexdigest()


def digest_allowed(allowed):
    """sha256 over the sorted `KMER\tweight` lines of an allowed_kmer_dict."""
    return digest("".join("%s\t%d\n" % kv for kv in sorted(allowed.items())))


def load_golden(name):
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "rt") as f:
        return json.load(f)
