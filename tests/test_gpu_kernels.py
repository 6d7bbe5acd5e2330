"""GPU parity tests, stage by stage, through the C-ABI (ctypes) against the oracle."""
import os
import sys

import numpy as np
import pytest

import helpers
from oracle import kmer_count
from oracle import shannon_oracle as so
from shannon_b200 import _lib, synth
from shannon_b200 import extension_correction as ec

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "golden"))
import cases  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _enc(kmers, k1=None):
    kmers = list(kmers)
    return ec.keys_array([ec.encode_kmer(k) for k in kmers], k1 or len(kmers[0]))


def _dec(keys, k1):
    if len(keys) == 0:
        return []
    t = ec.decode_kmers(keys, k1).tobytes().decode()
    return [t[i:i + k1] for i in range(0, len(t), k1)]


def _case(workdir, kind, seed):
    if kind == "synth":
        s1, s2 = helpers.synthetic_seqs(10, 1200, seed)
        return helpers.make_case(workdir, 24, s1, s2), 3, 75
    if kind in ("synth31", "synth32"):      # K1 = 32: full 64-bit keys; K1 = 33: two-word keys
        s1, s2 = helpers.synthetic_seqs(10, 1200, seed)
        return helpers.make_case(workdir, int(kind[-2:]), s1, s2), 3, 75
    if kind == "repeat32":
        reads = cases.repeat_rich_reads(seed, 600, 60, 220, 5)
        return helpers.make_case(workdir, 32, reads), 2, 45
    reads = cases.repeat_rich_reads(seed, 500, 40, 200, 5)
    return helpers.make_case(workdir, [8, 10, 12, 15][seed % 4], reads), 2, 25


CASES = [("synth", 1), ("synth", 2), ("repeat", 0), ("repeat", 1), ("repeat", 2), ("repeat", 3),
         ("synth31", 5), ("synth32", 6), ("repeat32", 7), ("repeat32", 8)]


# ---- a1/a2 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,seed", CASES)
@pytest.mark.parametrize("ds", [False, True])
def test_table_build_equals_load_kmers(ctx, workdir, kind, seed, ds):
    case, _, _ = _case(workdir, kind, seed)
    kmers, k1 = so.load_kmers(case.k1mer_org, ds)
    keys, counts, k1g = ctx.parse_kmer_file(case.k1mer_org)
    assert k1g == k1
    ctx.table_build(keys, counts, k1, ds)
    st = ctx.table_stats()
    assert st["n_distinct"] == len(kmers)
    with open(case.k1mer_org) as f:
        n_low = sum(so.low_complexity(l.split()[0]) for l in f)
    assert st["n_lowcomplexity"] == n_low
    gk, gw, gi = ctx.table_dump()
    # dump order = first-occurrence order = the reference dict's insertion order
    assert _dec(gk, k1) == list(kmers)
    assert gw.tolist() == list(kmers.values())
    assert np.all(np.diff(gi.astype(np.int64)) > 0)
    # lookups: every key, plus absent keys
    w, f = ctx.table_lookup(gk)
    assert f.all() and np.array_equal(w, gw)
    rng = np.random.default_rng(seed)
    if k1 <= 32:
        probe = rng.integers(0, (1 << (2 * k1)) - 1, size=5000, dtype=np.uint64, endpoint=True)
    else:
        probe = np.stack([rng.integers(0, (1 << 64) - 1, size=5000, dtype=np.uint64, endpoint=True),
                          rng.integers(0, 4, size=5000, dtype=np.uint64)], axis=1)
    # half of the probes: one-base mutations of present keys (near misses in the same word)
    mut = gk[rng.integers(0, len(gk), size=2500)].copy()
    if k1 <= 32:
        mut ^= np.uint64(1) << rng.integers(0, 2 * k1, size=2500).astype(np.uint64)
    else:
        bit = rng.integers(0, 2 * k1, size=2500)
        mut[np.arange(2500), bit // 64] ^= np.uint64(1) << (bit % 64).astype(np.uint64)
    probe = np.concatenate([probe, mut])
    w, f = ctx.table_lookup(probe)
    exp = [kmers.get(s) for s in _dec(probe, k1)]
    assert f.tolist() == [int(e is not None) for e in exp]
    assert w.tolist() == [e or 0 for e in exp]


def test_pack_and_lowcomplexity_edge_cases(ctx):
    k1 = 25
    ks = ["A" * 25, "A" * 23 + "CG", "A" * 22 + "CGT", "ACGT" * 6 + "A", "T" * 24 + "g",
          "c" * 25, "acgtacgtacgtacgtacgtacgtT"]
    keys = ctx.pack_kmers("".join(ks).encode(), len(ks), k1)
    assert _dec(keys, k1) == [k.upper() for k in ks]
    ctx.table_build(keys, np.arange(1, len(ks) + 1, dtype=np.uint32), k1, False)
    keep = [k.upper() for k in ks if not so.low_complexity(k.upper())]
    gk, gw, _ = ctx.table_dump()
    assert _dec(gk, k1) == keep
    with pytest.raises(_lib.ShnError):
        ctx.pack_kmers(b"ACGTN" * 5, 1, 25)
    # duplicate lines accumulate; k1 = 32 (all 64 key bits used)
    ks = ["ACGTTGCAACGTTGCAACGTTGCAACGTTGCA", "TTTTTTTTGGGGGGGGCCCCCCCCAAAAAAAT"]
    keys = ctx.pack_kmers("".join(ks).encode(), 2, 32)
    ctx.table_build(np.concatenate([keys, keys[:1]]), np.asarray([4, 9, 6], np.uint32), 32, False)
    gk, gw, gi = ctx.table_dump()
    assert _dec(gk, 32) == ks and gw.tolist() == [10, 9] and gi.tolist() == [0, 1]
    w, f = ctx.table_lookup(_enc([so.reverse_complement(ks[0])]))
    assert f.tolist() == [0]
    # k1 = 33: two words per key; keys that differ only in the high word are distinct
    ks = ["A" + "ACGTTGCAACGTTGCAACGTTGCAACGTTGCA", "G" + "ACGTTGCAACGTTGCAACGTTGCAACGTTGCA",
          "T" * 32 + "G", "TGCA" * 8 + "C"]
    assert [so.low_complexity(k) for k in ks] == [False, False, True, False]
    keys = ctx.pack_kmers("".join(ks).encode(), 4, 33)
    assert keys.shape == (4, 2) and np.array_equal(keys, _enc(ks))
    ctx.table_build(np.concatenate([keys, keys[1:2]]), np.asarray([4, 9, 6, 2, 5], np.uint32), 33, False)
    gk, gw, gi = ctx.table_dump()
    assert _dec(gk, 33) == [ks[0], ks[1], ks[3]] and gw.tolist() == [4, 14, 2] and gi.tolist() == [0, 1, 3]
    w, f = ctx.table_lookup(_enc(["C" + ks[0][1:], ks[1], so.reverse_complement(ks[3])]))
    assert f.tolist() == [0, 1, 0] and w.tolist() == [0, 14, 0]
    ctx.table_build(keys, np.arange(1, 5, dtype=np.uint32), 33, True)      # + reverse complements
    w, f = ctx.table_lookup(_enc([so.reverse_complement(ks[3]), so.reverse_complement(ks[0])]))
    assert f.tolist() == [1, 1] and w.tolist() == [4, 1]
    with pytest.raises(_lib.ShnError):
        ctx.pack_kmers(b"A" * 34, 1, 34)
    # empty input
    ctx.table_build(np.empty(0, np.uint64), np.empty(0, np.uint32), 25, False)
    assert ctx.table_stats()["n_distinct"] == 0


# ---- a3-a9 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,seed", CASES)
@pytest.mark.parametrize("spec", [False, True])
def test_l3_stages_equal_oracle(ctx, workdir, kind, seed, spec, monkeypatch):
    # spec=True forces the speculative windowed walk kernel (normally only used for components
    # with >= 60 k K1-mers) onto every component
    monkeypatch.setenv("SHN_SPEC_MIN_NODES", "1" if spec else "1000000000")
    monkeypatch.setenv("SHN_SPEC_TIER16", "2")   # first two components: 16-warp CTAs, the rest: 8-warp CTAs
    # hyperbola filter: decided on the device, or (borderline walks; here: all) by the host's libm
    monkeypatch.setenv("SHN_SHAPE_TOL", "1e9" if spec else "1e-9")
    # component sizes: block-private shared-memory histogram, or global atomics (many components)
    monkeypatch.setenv("SHN_COMP_HIST_MAX_BYTES", "0" if spec else "40960")
    case, min_weight, min_length = _case(workdir, kind, seed)
    out = case.outdir("o")
    res = so.run_correction(case.k1mer_org, out + "/k", min_weight, min_length, False, out, 2,
                            True, True)
    k1 = res.k1
    keys, counts, _ = ctx.parse_kmer_file(case.k1mer_org)
    ctx.table_build(keys, counts, k1, False)
    dump_before = ctx.table_dump()
    sz = ctx.l3_run(min_weight, min_length)
    # the walks borrow the first-occurrence word of every slot: the table must come back unchanged
    for x, y in zip(dump_before, ctx.table_dump()):
        assert np.array_equal(x, y)
    # walks, in pop order
    seed_k, nl, nr, tot, flags = ctx.l3_walks()
    exp = res.walks
    assert sz["n_walks"] == len(exp), (sz, len(exp))
    assert _dec(seed_k, k1) == [w.seed for w in exp]
    assert nl.tolist() == [w.n_left for w in exp]
    assert nr.tolist() == [w.n_right for w in exp]
    assert tot.tolist() == [w.tot_wt for w in exp]
    assert sz["n_traversed"] == len(res.traversed)
    assert (sz["n_spec_comps"] > 0) == spec
    assert [bool(f & 1) for f in flags] == [w.passes_shape for w in exp]
    assert [bool(f & 2) for f in flags] == [w.passes_shape and w.duplicate for w in exp]
    assert [bool(f & 4) for f in flags] == [w.accepted for w in exp]
    # contigs in acceptance order
    bases, offs = ctx.l3_contigs()
    txt = bases.tobytes().decode()
    got = [txt[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1)]
    assert got == res.contigs[1:]
    # allowed K1-mers + weights
    ak, aw = ctx.l3_allowed()
    assert dict(zip(_dec(ak, k1), aw.tolist())) == res.allowed_kmer_dict
    assert len(ak) == len(res.allowed_kmer_dict)
    # contig graph: distinct edges with multiplicities, first positions, component labels
    a, b, w, fp = ctx.l3_edges()
    exp_edges = sorted((x, y, wt) for y, nb in res.connections.items() for x, wt in nb.items() if x < y)
    assert sorted(zip(a.tolist(), b.tolist(), w.tolist())) == exp_edges
    assert list(zip(b.tolist(), a.tolist())) == sorted(zip(b.tolist(), a.tolist()))
    n = len(res.contigs) - 1
    adj = ec.contig_adjacency(n, a, b, w, fp)
    for x in range(1, n + 1):
        assert adj[x] == list(res.connections[x].items())
    lab = ctx.l3_labels()
    assert lab[1:].tolist() == [res.contig2component[x] for x in range(1, n + 1)]


# ---- a10-a12 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [24, 32])
@pytest.mark.parametrize("paired", [False, True])
def test_l4_assign_equals_oracle(ctx, workdir, paired, K):
    s1, s2 = helpers.synthetic_seqs(10, 1200, 4)
    # ragged + dirty reads: short, exactly K1, K1+1, with N, lower case, empty mate
    s1[5], s2[5] = s1[5][:K + 1], s2[5][:K + 2]
    s1[6] = s1[6][:10]
    s1[7] = s1[7][:40] + "N" + s1[7][41:]
    s2[8] = s2[8].lower()
    s1[9] = s1[9][:57]
    s2[10] = s2[10][:2 * (K + 1)]
    s1[11] = s1[11][:2 * (K + 1) + 1]
    case = helpers.make_case(workdir, K, s1, s2 if paired else None)
    out = case.outdir("o")
    res = so.run_correction(case.k1mer_org, out + "/k", 3, 75, False, out, 500, True, True)
    k1 = res.k1
    # two-membership map: every contig in a 'c' part and an 'r2' part, as repartition does
    contigs = res.contigs[1:]
    comp_a = [i % 3 for i in range(len(contigs))]
    comp_b = [3 + (i % 2) for i in range(len(contigs))]
    entries = [(c, comp_a[i]) for i, c in enumerate(contigs)] + \
              [(c, comp_b[i]) for i, c in enumerate(contigs)]
    text = "".join(c for c, _ in entries)
    offs = np.zeros(len(entries) + 1, np.uint64)
    offs[1:] = np.cumsum([len(c) for c, _ in entries])
    total = sum(max(len(c) - k1 + 1, 0) for c, _ in entries)
    ctx.l4_map_add_contigs(np.frombuffer(text.encode(), np.uint8), offs,
                           np.asarray([c for _, c in entries], np.uint32), k1, True, total)
    k2c = {}
    for c, cid in entries:
        for p in range(len(c) - k1 + 1):
            k2c.setdefault(c[p:p + k1], [set(), 0])[0].add(cid)
    files = case.reads_files
    rb, ro = ctx.load_fasta(files[0])
    ctx.l4_load_reads(0, rb, ro)
    recs = [helpers.read_fasta_seqs(f) for f in files]
    if paired:
        rb1, ro1 = ctx.load_fasta(files[1], len(ro) - 1)
        ctx.l4_load_reads(1, rb1, ro1)
    na, nlook, nvalid = ctx.l4_assign(paired, k1)
    comp_offs, idx = ctx.l4_assignments(5, na)
    exp = [[] for _ in range(5)]
    exp_valid = exp_look = 0
    for r in range(len(recs[0])):
        mates = [rr[r] for rr in recs]
        if any(m.strip("ACTG") for m in mates):
            continue
        exp_valid += 1
        comps = set()
        for m in mates:
            comps |= so.read_components(m, k2c, k1)
            exp_look += len(so.sample_k1mers(m, k1)) if len(m) >= k1 else 0
        for c in comps:
            exp[c].append(r)
    got = [idx[int(comp_offs[c]):int(comp_offs[c + 1])].tolist() for c in range(5)]
    assert got == exp
    assert nvalid == exp_valid and nlook == exp_look
    assert sum(len(e) for e in exp) == na > 0
    # window weights (a12)
    ak = _enc(res.allowed_kmer_dict, k1)
    ctx.l4_map_set_weights(ak, np.asarray(list(res.allowed_kmer_dict.values()), np.uint32))
    ww, woff = ctx.l4_map_window_weights(np.frombuffer(text.encode(), np.uint8), offs, k1)
    expw = [res.allowed_kmer_dict.get(c[p:p + k1], 0) for c, _ in entries
            for p in range(len(c) - k1 + 1)]
    assert ww.tolist() == expw


def test_l4_more_than_two_components_is_a_loud_error(ctx):
    c = "ACGTTGCAAGGCTTAACCGGTTAACGATCGATTAGC"
    text = c * 3
    offs = np.asarray([0, len(c), 2 * len(c), 3 * len(c)], np.uint64)
    with pytest.raises(_lib.ShnError):
        ctx.l4_map_add_contigs(np.frombuffer(text.encode(), np.uint8), offs,
                               np.asarray([0, 1, 2], np.uint32), 25, True, 100)


# ---- inputs of the path ----------------------------------------------------------------------
def test_synth_revcomp_count_equal_cpu_twins(ctx, workdir):
    tx = synth.make_transcripts(8, 21)
    codes, offs = synth.pack_transcripts(tx)
    thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], True)
    n, L = 700, 100
    m1, m2 = synth.make_pairs(codes, offs, thr, n, 77, first_pair=13)
    d_tx, d_off, d_thr = ctx.to_device(codes), ctx.to_device(offs), ctx.to_device(thr)
    d1, d2 = ctx.dev_alloc(2 * n * L), ctx.dev_alloc(2 * n * L)
    ctx.synth_pairs(d_tx, d_off, d_thr, len(tx), n, 13, 77, L, 300, synth.ERR_THRESHOLD_24, d1, d2)
    g1 = ctx.d2h(np.empty((n, L), np.uint8), d1)
    g2 = ctx.d2h(np.empty((n, L), np.uint8), d2)
    assert np.array_equal(g1, m1) and np.array_equal(g2, m2)
    # RC doubling: reads_1 = [R1 ; rc(R2)], reads_2 = [rc(R1) ; R2]
    ctx.revcomp_reads(d2, d1 + n * L, n, L)
    ctx.revcomp_reads(d1, d2 + n * L, n, L)   # writes rc(R1) behind R2 ... then swap halves below
    r1 = ctx.d2h(np.empty((2 * n, L), np.uint8), d1)
    r2 = ctx.d2h(np.empty((2 * n, L), np.uint8), d2)
    e1, e2 = synth.rc_double(m1, m2)
    assert np.array_equal(r1, e1)
    assert np.array_equal(r2[n:], e2[:n]) and np.array_equal(r2[:n], e2[n:])
    # k-mer counting stand-in
    p1, p2 = os.path.join(workdir, "a.fa"), os.path.join(workdir, "b.fa")
    synth.write_fasta(p1, r1)
    synth.write_fasta(p2, r2)
    exp = kmer_count.count_k1mers([p1, p2], 25)
    dk, dc, nd = ctx.count_k1mers([d1, d2], [2 * n, 2 * n], L, 25, 4 * n * 76)
    assert nd == len(exp)
    gk = ctx.d2h(np.empty(nd, np.uint64), dk)
    gc = ctx.d2h(np.empty(nd, np.uint32), dc)
    ks = sorted(exp)
    assert _dec(gk, 25) == ks
    assert gc.tolist() == [exp[k] for k in ks]
    for d in (d_tx, d_off, d_thr, d1, d2):
        ctx.dev_free(d)
