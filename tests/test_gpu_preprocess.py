"""Rows f1/f2: RC doubling of read files and K1-mer counting on the device against their
CPU restatements (oracle/preprocess_oracle.py pinned to rc_s.py; oracle/kmer_count.py, the documented
jellyfish stand-in), and the FASTA -> partition chain without intermediate text."""
import os
import sys

import numpy as np
import pytest

import helpers
from oracle import kmer_count
from oracle import preprocess_oracle as po
from oracle import shannon_oracle as so

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "golden"))
import cases  # noqa: E402


def _ctx():
    from shannon_b200.extension_correction import get_context
    return get_context()


def _write(path, names, seqs):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n%s\n" % (n, s))


def _sample_files(workdir, kind):
    """bundled samples (reads of ~50 and 100 bp, a third of them with N) as raw FASTA files"""
    os.makedirs(workdir, exist_ok=True)
    if kind == "se":
        seqs = cases.sample_seqs("SE_read")
        p = os.path.join(workdir, "se.fasta")
        _write(p, ["SE_%d some text" % i for i in range(len(seqs))], seqs)
        return [p]
    out = []
    for m in (1, 2):
        seqs = cases.sample_seqs("PE_read_%d" % m)[:1500]
        p = os.path.join(workdir, "pe_%d.fasta" % m)
        _write(p, ["PE_%d/%d" % (i, m) for i in range(len(seqs))], seqs)
        out.append(p)
    return out


@pytest.mark.parametrize("kind,paired,ds", [("se", False, True), ("pe", True, True), ("pe", True, False),
                                            ("se", False, False)])
def test_rc_double_files_equal_driver(workdir, kind, paired, ds):
    from shannon_b200 import preprocess
    files = _sample_files(os.path.join(workdir, "in"), kind)
    if kind == "se":     # rc_s.py's line handling: blank lines, padded lines, a second field
        with open(files[0], "a") as f:
            f.write("\n>padded  \n  ACGTNACGT  \n>two fields\nGATTACA xyz\n")
    a, b = os.path.join(workdir, "a"), os.path.join(workdir, "b")
    os.makedirs(a)
    os.makedirs(b)
    exp = po.rc_double(files, a, paired, ds)
    got, n, l = preprocess.rc_double(_ctx(), files, b, paired, ds)
    assert [os.path.basename(x) for x in exp] == [os.path.basename(x) for x in got]
    for x, y in zip(exp, got):
        assert open(x, "rb").read() == open(y, "rb").read(), os.path.basename(x)
    assert sorted(os.listdir(a)) == sorted(os.listdir(b))
    assert n > 1000 and 40 < l <= 100


def test_rc_rejects_what_rc_s_rejects(workdir):
    from shannon_b200 import _lib, preprocess
    p = os.path.join(workdir, "bad.fasta")
    _write(p, ["x"], ["ACGTacgt"])
    with pytest.raises(KeyError):
        po.reverse_complement_file(p, p + ".rc")
    with pytest.raises(_lib.ShnError):
        preprocess.rc_double(_ctx(), [p], workdir, False, True)


@pytest.mark.parametrize("kind,K,cutoff", [("se", 24, 1), ("pe", 24, 1), ("pe", 32, 1), ("se", 24, 2)])
def test_jellyfish_count_equals_stand_in(workdir, kind, K, cutoff, monkeypatch):
    from shannon_b200 import preprocess
    monkeypatch.setattr(preprocess, "COUNT_CHUNK_BASES", 40000)      # many chunks
    files = _sample_files(os.path.join(workdir, "in"), kind)
    doubled = po.rc_double(files, workdir, kind == "pe", True)
    exp = os.path.join(workdir, "exp.dict")
    kmer_count.write_dict(kmer_count.count_k1mers(doubled, K + 1), exp, min_count=cutoff)
    got = os.path.join(workdir, "got.dict")
    d_keys, d_counts, n = preprocess.jellyfish_count(_ctx(), doubled, K, got, cutoff)
    assert open(exp).read() == open(got).read()
    assert n == sum(1 for _ in open(exp)) > 1000


def test_count_table_grows_when_estimate_is_too_small(workdir):
    from shannon_b200 import preprocess
    s1, _ = helpers.synthetic_seqs(6, 800, 3)
    p = os.path.join(workdir, "r.fasta")
    _write(p, ["r%d" % i for i in range(len(s1))], s1)
    ctx = _ctx()
    _, _, bases, offs = ctx.load_fasta_named(p)
    preprocess._count_arrays(ctx, [(bases, offs)], 25, expected=64)   # far too small: doubled until it fits
    _, _, n = ctx.count_finish(1)
    assert n == len(kmer_count.count_k1mers([p], 25))


def test_frontend_from_fasta_equals_file_pipeline(workdir):
    """raw FASTA -> device RC doubling -> device counting -> table -> walks -> partition, against the
    oracle run on the files the driver would have written."""
    from shannon_b200 import preprocess
    s1, s2 = helpers.synthetic_seqs(30, 5000, 52)
    s1[7] = s1[7][:40] + "N" + s1[7][41:]
    s2[9] = s2[9][:60]
    raw = [os.path.join(workdir, "raw_1.fasta"), os.path.join(workdir, "raw_2.fasta")]
    _write(raw[0], ["p%d/1" % i for i in range(len(s1))], s1)
    _write(raw[1], ["p%d/2" % i for i in range(len(s2))], s2)
    case = helpers.make_case(os.path.join(workdir, "case"), 24, s1, s2)     # the driver's files
    out, allowed, _, ret = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case,
                                                "ora", partition_size=3, inMem=True, repartition=False)
    _, new_comps, _, rps = ret
    cor, comp_offs, rec_idx, stats, mates = preprocess.frontend_from_fasta(
        _ctx(), raw, 24, True, True, 3, 75, 3)
    assert cor.contigs.strings() == open(out + "/algo_input/k1mer.dict_contig").read().split()
    names = []
    for i in sorted(ret[0]):
        names += ["c%d_%d" % (i + 1, p) for p in range(ret[0][i])]
    names += [c for c in new_comps if c.startswith("cremaining")]
    assert len(names) == stats["n_partitions"] > 1
    recs = [helpers.read_fasta_seqs(f) for f in case.reads_files]
    for cid, name in enumerate(names):
        sel = rec_idx[comp_offs[cid]:comp_offs[cid + 1]].tolist()
        assert [recs[0][r] for r in sel] == rps[name][0][0]
        assert [recs[1][r] for r in sel] == rps[name][1][0]
