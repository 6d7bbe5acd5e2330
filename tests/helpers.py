"""Shared test plumbing: build an input case (read files + k1mer.dict_org), replay the driver's
call sequence of shannon.py:395-467 against an implementation (reference / oracle / GPU), and
snapshot everything observable for comparison."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import kmer_count  # noqa: E402
from shannon_b200 import synth  # noqa: E402

GPMETIS = os.path.join(ROOT, "oracle", "gpmetis")


def read_fasta_seqs(path):
    with open(path) as f:
        return [l.rstrip("\n") for i, l in enumerate(f) if i % 2 == 1]


def rc_str(s):
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def write_reads(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">r%d\n%s\n" % (i, s))


class Case(object):
    """Input files for one front-end run, laid out like the driver does."""

    def __init__(self, workdir, K, paired_end):
        self.workdir, self.K, self.paired_end = workdir, K, paired_end
        self.algo_input = os.path.join(workdir, "in_algo_input")
        os.makedirs(self.algo_input, exist_ok=True)
        self.k1mer_org = os.path.join(self.algo_input, "k1mer.dict_org")
        self.reads_files = []

    def outdir(self, name):
        d = os.path.join(self.workdir, name)
        os.makedirs(os.path.join(d, "algo_input"), exist_ok=True)
        return d


def make_case(workdir, K, seqs1, seqs2=None, double_stranded=True, fast_count=False):
    """RC-double the reads like shannon.py:395-424 and count (K+1)-mers with the stand-in
    (fast_count: its numpy twin, for the large fixtures)."""
    case = Case(workdir, K, seqs2 is not None)
    if seqs2 is None:
        reads = seqs1 + [rc_str(s) for s in seqs1] if double_stranded else list(seqs1)
        p = os.path.join(case.algo_input, "reads.fasta")
        write_reads(p, reads)
        case.reads_files = [p]
    else:
        if double_stranded:
            r1 = seqs1 + [rc_str(s) for s in seqs2]
            r2 = [rc_str(s) for s in seqs1] + seqs2
        else:
            r1, r2 = list(seqs1), [rc_str(s) for s in seqs2]
        p1 = os.path.join(case.algo_input, "reads_1.fasta")
        p2 = os.path.join(case.algo_input, "reads_2.fasta")
        write_reads(p1, r1)
        write_reads(p2, r2)
        case.reads_files = [p1, p2]
    if fast_count:
        kmer_count.write_dict_numpy(*kmer_count.count_k1mers_numpy(case.reads_files, K + 1), case.k1mer_org)
    else:
        kmer_count.write_dict(kmer_count.count_k1mers(case.reads_files, K + 1), case.k1mer_org)
    return case


def synthetic_seqs(n_transcripts, n_pairs, seed, skewed=False, read_len=100, frag_len=300,
                   err_threshold=synth.ERR_THRESHOLD_24):
    tx = synth.make_transcripts(n_transcripts, seed)
    codes, offs = synth.pack_transcripts(tx)
    thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], skewed)
    m1, m2 = synth.make_pairs(codes, offs, thr, n_pairs, seed, read_len=read_len,
                              frag_len=frag_len, err_threshold=err_threshold)
    return ([bytes(r).decode() for r in m1], [bytes(r).decode() for r in m2])


def run_frontend(ec_fn, kfc_fn, case, outname, min_weight=3, min_length=75, partition_size=500,
                 inMem=False, ec_inMem=True, repartition=True, only_reads=False,
                 double_stranded_load=False, extra_kfc=None):
    """shannon.py:450-467.  ec_fn / kfc_fn are the two entry points of an implementation."""
    out = case.outdir(outname)
    k1dict = os.path.join(out, "algo_input", "k1mer.dict")
    args = []
    if double_stranded_load:
        args.append("-d")
    args += [case.k1mer_org, k1dict, str(min_weight), str(min_length), out, str(partition_size),
             "1"] + case.reads_files
    k1mer_dictionary, reads = ec_fn(args, ec_inMem)
    kw = dict(extra_kfc or {})
    ds_kfc = kw.pop("double_stranded", False)
    ret = kfc_fn(k1mer_dictionary, os.path.join(out, "algo_input"), reads, case.reads_files, out,
                 "contigs.txt", True, ds_kfc, case.paired_end, repartition, partition_size, 2,
                 case.K, GPMETIS, 5, only_reads, inMem, 1, **kw)
    return out, dict(k1mer_dictionary), reads, ret


SKIP_FILES = ("before_sp_log.txt",)


def snapshot(outdir):
    """{relative path: bytes} of every file under outdir except timestamped logs."""
    snap = {}
    for base, _, files in os.walk(outdir):
        for fn in files:
            if fn in SKIP_FILES:
                continue
            p = os.path.join(base, fn)
            with open(p, "rb") as f:
                snap[os.path.relpath(p, outdir)] = f.read()
    return snap


def normalise_ret(ret):
    """kmers_for_component's return value as plain comparable data."""
    components_broken, new_comps, contig_weights, rps = ret
    if isinstance(new_comps, dict):
        new_comps = [(k, list(v)) for k, v in new_comps.items()]
    else:
        new_comps = list(new_comps)
    if isinstance(contig_weights, dict):
        contig_weights = [(k, [list(map(int, w)) for w in v]) for k, v in contig_weights.items()]
    else:
        contig_weights = list(contig_weights)
    rps = [(k, [[list(x) for x in part] if part and isinstance(part[0], list) else list(part)
                for part in v]) for k, v in rps.items()]
    return dict(components_broken), new_comps, contig_weights, rps


def assert_same_run(a, b, label=""):
    out_a, dict_a, reads_a, ret_a = a
    out_b, dict_b, reads_b, ret_b = b
    assert dict_a == dict_b, label + ": allowed_kmer_dict differs"
    assert list(reads_a) == list(reads_b)
    sa, sb = snapshot(out_a), snapshot(out_b)
    assert sorted(sa) == sorted(sb), label + ": file sets differ: %s" % (
        sorted(set(sa) ^ set(sb)),)
    for k in sa:
        if k.endswith("algo_input/k1mer.dict"):
            # written from a Python set in the reference (extension_correction.py:406):
            # line order is hash-seed dependent, compare as a multiset of lines
            assert sorted(sa[k].splitlines()) == sorted(sb[k].splitlines()), label + ": " + k
        else:
            assert sa[k] == sb[k], label + ": file %s differs" % k
    assert normalise_ret(ret_a) == normalise_ret(ret_b), label + ": return value differs"


def check_against_golden(ec_fn, kfc_fn, name, workdir, label):
    """Re-create golden case ``name``, run an implementation on it and compare with what the
    real reference produced (tests/golden/<name>.json.gz)."""
    import hashlib
    gdir = os.path.join(ROOT, "tests", "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import cases
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    seqs1, seqs2 = cases.case_inputs(spec)
    case = make_case(workdir, spec["K"], seqs1, seqs2, double_stranded=spec.get("rc_double", True),
                     fast_count=spec.get("fast_count", False))
    with open(case.k1mer_org, "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == gold["k1mer_dict_org_sha256"], \
            "input k1mer.dict_org differs from the one the golden was generated from"
    out, allowed, reads, ret = run_frontend(ec_fn, kfc_fn, case, "run", **spec["run"])
    if gold.get("digest"):      # large fixture: sha256 of every output
        assert len(allowed) == gold["n_allowed"], label + ": allowed_kmer_dict size differs from golden"
        assert cases.digest_allowed(dict(allowed)) == gold["allowed_sha256"], label + ": allowed_kmer_dict"
        assert list(reads) == []
        snap = snapshot(out)
        assert sorted(snap) == sorted(gold["files_sha256"]), label + ": file set differs"
        for k, v in snap.items():
            if k.endswith("algo_input/k1mer.dict"):
                v = b"".join(sorted(v.splitlines(True)))
            assert cases.digest(v) == gold["files_sha256"][k], label + ": file %s differs from golden" % k
        cb, new_comps, cw, rps = normalise_ret(ret)
        got = {"components_broken": dict((str(k), v) for k, v in cb.items()),
               "new_comps": new_comps, "contig_weights": cw, "rps": rps}
        assert cases.digest(json.dumps(json.loads(json.dumps(got)), sort_keys=True)) == gold["ret_sha256"], \
            label + ": return value differs from golden"
        return case, out
    assert allowed == gold["allowed_kmer_dict"], label + ": allowed_kmer_dict differs from golden"
    assert list(reads) == []
    snap = snapshot(out)
    assert sorted(snap) == sorted(gold["files"]), label + ": file set differs: %s" % (
        sorted(set(snap) ^ set(gold["files"])),)
    for k, v in snap.items():
        g = gold["files"][k].encode()
        if k.endswith("algo_input/k1mer.dict"):
            assert sorted(v.splitlines()) == sorted(g.splitlines()), label + ": " + k
        else:
            assert v == g, label + ": file %s differs from golden" % k
    cb, new_comps, cw, rps = normalise_ret(ret)
    got = json.loads(json.dumps({"components_broken": dict((str(k), v) for k, v in cb.items()),
                                 "new_comps": new_comps, "contig_weights": cw, "rps": rps}))
    assert got == gold["ret"], label + ": return value differs from golden"
    return case, out


def transcripts_file(path, seed, n=60):
    """Transcript-like FASTA with the cases faster_reps.py is about: exact sub-sequences, equal
    copies (name order decides), reverse-complement copies, near misses (length off by >= 3)."""
    import random
    rnd = random.Random(seed)
    rand = lambda k: "".join(rnd.choice("ACGT") for _ in range(k))  # noqa: E731
    base = [rand(rnd.randrange(60, 400)) for _ in range(n // 3)]
    seqs = list(base)
    for s in base:
        a = rnd.randrange(0, len(s) // 3)
        b = rnd.randrange(2 * len(s) // 3, len(s))
        kind = rnd.randrange(6)
        if kind == 0:
            seqs.append(s[a:b])                                   # contained
        elif kind == 1:
            seqs.append(s)                                        # identical copy
        elif kind == 2:
            seqs.append(rc_str(s[a:b]))                           # contained on the other strand
        elif kind == 3:
            seqs.append(s[a:a + 30] + rand(5) + s[a + 30:b])      # ends match, length off by 5
        elif kind == 4:
            seqs.append(s[a:a + 40] + s[a + 42:b])                # ends match, length off by 2
        else:
            seqs.append(s[:20])                                   # shorter than r
    rnd.shuffle(seqs)
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">T%d_%d len=%d\n%s\n" % (rnd.randrange(1000), i, len(s), s))
    return path


def debruijn_case(path, seed, K=8, n_seqs=6, length=120, acyclic=True):
    """A `k1mer.dict`-style file (K1-mer<TAB>prevalence) of a small graph with branches, tips, a
    repeated line and shared segments; acyclic=True rejects inputs that contain a pure cycle of
    unambiguous edges (its condensation depends on the visiting order, mbgraph.py:479)."""
    import random
    rnd = random.Random(seed)
    while True:
        seqs = []
        for _ in range(n_seqs):
            s = "".join(rnd.choice("ACGT") for _ in range(length))
            seqs.append(s)
            if rnd.random() < 0.7:                       # a variant sharing both ends: a bubble
                p = rnd.randrange(K + 2, length - K - 2)
                seqs.append(s[:p] + rnd.choice("ACGT") + s[p + 1:])
            if rnd.random() < 0.5:                       # a tip
                p = rnd.randrange(K, length - K)
                seqs.append(s[p:p + K + 1 + rnd.randrange(1, 6)][:-1] + "".join(rnd.choice("ACGT") for _ in range(4)))
        counts = {}
        for s in seqs:
            for i in range(len(s) - K):
                km = s[i:i + K + 1]
                counts[km] = counts.get(km, 0) + rnd.randrange(1, 9)
        lines = list(counts.items())
        rnd.shuffle(lines)
        lines.append(lines[0])                            # a repeated line = a parallel edge
        if acyclic and _has_unambiguous_cycle(lines, K):
            continue
        with open(path, "w") as f:
            for km, c in lines:
                f.write("%s\t%d\n" % (km, c))
        return path


def _has_unambiguous_cycle(lines, K):
    out_deg, in_deg, nxt = {}, {}, {}
    for km, _ in lines:
        a, b = km[:-1], km[1:]
        out_deg[a] = out_deg.get(a, 0) + 1
        in_deg[b] = in_deg.get(b, 0) + 1
    for km, _ in lines:
        a, b = km[:-1], km[1:]
        if out_deg[a] == 1 and in_deg[b] == 1 and a != b:
            nxt[a] = b
    for start in nxt:
        seen, cur = set(), start
        while cur in nxt and cur not in seen:
            seen.add(cur)
            cur = nxt[cur]
        if cur in seen:
            return True
    return False
