"""Pins oracle/shannon_oracle.py to the REAL reference (loaded from /root/reference through
oracle/ref_loader.py).  Runs only where the reference tree exists (the build container);
on the GPU box the committed fixtures in tests/golden/ carry the same pin."""
import os

import pytest

import helpers
from oracle import ref_loader, shannon_oracle

pytestmark = pytest.mark.skipif(not ref_loader.available(),
                                reason="reference tree not present (GPU box)")

SAMPLES = os.path.join(ref_loader.REFERENCE_DIR, "Samples")


def both(case, **kw):
    ref_ec = ref_loader.load("extension_correction")
    ref_kfc = ref_loader.load("kmers_for_component")
    a = helpers.run_frontend(ref_ec.extension_correction, ref_kfc.kmers_for_component,
                             case, "ref", **kw)
    b = helpers.run_frontend(shannon_oracle.extension_correction,
                             shannon_oracle.kmers_for_component, case, "ora", **kw)
    helpers.assert_same_run(a, b, "oracle vs reference")
    return a, b


def test_sample_se(workdir):
    seqs = helpers.read_fasta_seqs(os.path.join(SAMPLES, "SE_read.fasta"))
    both(helpers.make_case(workdir, 24, seqs))


def test_sample_pe_subset(workdir):
    s1 = helpers.read_fasta_seqs(os.path.join(SAMPLES, "PE_read_1.fasta"))[:600]
    s2 = helpers.read_fasta_seqs(os.path.join(SAMPLES, "PE_read_2.fasta"))[:600]
    both(helpers.make_case(workdir, 24, s1, s2))


@pytest.mark.parametrize("seed,ntx,npairs,psize", [(1, 12, 1500, 500), (2, 30, 3000, 2),
                                                   (3, 40, 6000, 1)])
def test_synthetic_pe(workdir, seed, ntx, npairs, psize):
    s1, s2 = helpers.synthetic_seqs(ntx, npairs, seed)
    case = helpers.make_case(workdir, 24, s1, s2)
    a, _ = both(case, partition_size=psize)
    if psize < 10:
        assert any(f.startswith("component1") for f in os.listdir(a[0])), \
            "case was meant to exercise the gpmetis branch"


def test_synthetic_se_inmem_and_small_k(workdir):
    s1, _ = helpers.synthetic_seqs(10, 1500, 7)
    both(helpers.make_case(workdir, 16, s1), inMem=True, partition_size=2, min_length=50)


def test_double_stranded_load_flag(workdir):
    s1, s2 = helpers.synthetic_seqs(8, 800, 11)
    case = helpers.make_case(workdir, 24, s1, s2, double_stranded=False)
    both(case, double_stranded_load=True, ec_inMem=False)


def _repeat_rich_reads(seed, n_reads, read_len, genome_len, n_genomes):
    """Low-entropy genomes with copied segments: forces weight ties, repeated r-mers/C-mers,
    self-overlapping contigs -- the tie-break paths of duplicate_check and the DFS."""
    import random
    rnd = random.Random(seed)
    genomes = []
    for _ in range(n_genomes):
        g = [rnd.choice("ACGT") for _ in range(genome_len)]
        for _ in range(3):                       # copy a segment somewhere else
            a = rnd.randrange(0, genome_len - 40)
            b = rnd.randrange(0, genome_len - 40)
            g[b:b + 30] = g[a:a + 30]
        genomes.append("".join(g))
    # isoform-like variants: splice out a middle piece
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


@pytest.mark.parametrize("seed", range(12))
def test_repeat_rich_small_k(workdir, seed):
    reads = _repeat_rich_reads(seed, 400, 40, 160, 4)
    K = [8, 10, 12, 15][seed % 4]
    case = helpers.make_case(workdir, K, reads)
    both(case, min_weight=2, min_length=20 + seed, partition_size=1 + seed % 3,
         inMem=bool(seed & 1))


def test_rc_file_oracle_equals_rc_s(workdir):
    """oracle/preprocess_oracle.py against the real rc_s.py (row f2: RC doubling)."""
    from oracle import preprocess_oracle as po
    src = os.path.join(workdir, "in.fasta")
    with open(src, "w") as f:
        f.write(">r0 some description\nACGTNNACGT\n\n>r1\nTTTTGCA  \n  >r2\t\nGATTACA trailing\n>r3\nN\n")
    rc_s = ref_loader.load("rc_s")
    rc_s.reverse_complement_serial(src, os.path.join(workdir, "ref.fasta"))
    po.reverse_complement_file(src, os.path.join(workdir, "ora.fasta"))
    assert open(os.path.join(workdir, "ref.fasta")).read() == open(os.path.join(workdir, "ora.fasta")).read()
    for name in ("SE_read", "PE_read_1"):
        p = os.path.join(SAMPLES, name + ".fasta")
        rc_s.reverse_complement_serial(p, os.path.join(workdir, "ref2.fasta"))
        po.reverse_complement_file(p, os.path.join(workdir, "ora2.fasta"))
        assert open(os.path.join(workdir, "ref2.fasta")).read() == open(os.path.join(workdir, "ora2.fasta")).read()


@pytest.mark.parametrize("seed,ds", [(1, False), (2, True), (3, True), (4, False)])
def test_reps_oracle_equals_faster_reps(workdir, seed, ds):
    """oracle/reps_oracle.py against the real faster_reps.py (row f4)."""
    from oracle import reps_oracle
    src = helpers.transcripts_file(os.path.join(workdir, "t.fasta"), seed)
    ref = ref_loader.load("faster_reps")
    ref.find_reps(src, os.path.join(workdir, "ref.fasta"), ds)
    reps_oracle.find_reps(src, os.path.join(workdir, "ora.fasta"), ds)
    a, b = open(os.path.join(workdir, "ref.fasta")).read(), open(os.path.join(workdir, "ora.fasta")).read()
    assert a == b
    assert 0 < a.count(">") < open(src).read().count(">")


@pytest.mark.parametrize("seed,K", [(1, 8), (2, 8), (3, 12), (4, 24), (5, 31)])
def test_mbgraph_oracle_equals_reference(workdir, seed, K):
    """oracle/mbgraph_oracle.py against the real multibridging.load_single_jellyfish +
    Node.condense_all (row f3)."""
    from oracle import mbgraph_oracle
    path = helpers.debruijn_case(os.path.join(workdir, "k1mer.dict"), seed, K=K, acyclic=False)
    mb = ref_loader.load("multibridging")
    mb.Read.K = K
    mb.load_single_jellyfish(path)
    mb.Node.condense_all()
    ref_nodes = sorted((n.bases, float(n.count), float(n.prevalence), float(n.norm), float(n.copy_count))
                       for n in mb.Node.nodes)
    ref_edges = sorted((n.bases, e.out_node.bases, int(e.weight), float(e.copy_count))
                       for n in mb.Node.nodes for e in n.out_edges)
    g = mbgraph_oracle.load_and_condense(path, K)
    nodes, edges = g.snapshot()
    assert nodes == ref_nodes and edges == ref_edges
    assert len(nodes) > 5 and any(c > 1 for _, c, _, _, _ in nodes)


def test_mbgraph_oracle_on_pipeline_output(workdir):
    """the real consumer path: a per-component k1mer.dict written by kmers_for_component"""
    from oracle import mbgraph_oracle
    s1, s2 = helpers.synthetic_seqs(10, 1500, 5)
    case = helpers.make_case(workdir, 24, s1, s2)
    out, _, _, ret = helpers.run_frontend(shannon_oracle.extension_correction,
                                          shannon_oracle.kmers_for_component, case, "ora")
    path = os.path.join(out, "component" + ret[1][0] + "k1mers_allowed.dict")
    mb = ref_loader.load("multibridging")
    mb.Read.K = 24
    mb.load_single_jellyfish(path)
    mb.Node.condense_all()
    g = mbgraph_oracle.load_and_condense(path, 24)
    nodes, edges = g.snapshot()
    assert nodes == sorted((n.bases, float(n.count), float(n.prevalence), float(n.norm), float(n.copy_count))
                           for n in mb.Node.nodes)
    assert edges == sorted((n.bases, e.out_node.bases, int(e.weight), float(e.copy_count))
                           for n in mb.Node.nodes for e in n.out_edges)


@pytest.mark.parametrize("paired,inmem", [(False, False), (True, False), (True, True)])
def test_kfc_double_stranded_true_nJobs1(workdir, paired, inmem):
    """kmers_for_component(double_stranded=True) -- never passed by shannon.py:427, deterministic for
    nJobs = 1: every chunk of reads is followed by its reverse complements (:36-52,117-141,341,381)."""
    s1, s2 = helpers.synthetic_seqs(12, 1500, 21)
    s1[4] = s1[4][:50] + "N" + s1[4][51:]
    case = helpers.make_case(workdir, 24, s1, s2 if paired else None, double_stranded=False)
    both(case, inMem=inmem, extra_kfc={"double_stranded": True}, min_weight=2, min_length=60)
