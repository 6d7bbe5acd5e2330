"""Size-independent properties of the GPU path at a size the Python oracle cannot reach in a test
(1 M read pairs, ~37 M K1-mers): invariants I1-I8 of SURVEY 8a, determinism, conservation laws."""
import numpy as np
import pytest

import helpers  # noqa: F401
from shannon_b200 import _lib, pipeline, synth

pytestmark = pytest.mark.gpu

K1, L = 25, 100


@pytest.fixture(scope="module")
def run():
    ctx = _lib.Context(0)
    n_pairs, n_tx, seed = 1000000, 500, 77
    tx = synth.make_transcripts(n_tx, seed)
    codes, offs = synth.pack_transcripts(tx)
    thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], True)
    d_tx, d_off, d_thr = ctx.to_device(codes), ctx.to_device(offs), ctx.to_device(thr)
    n_rec = 2 * n_pairs
    d1, d2 = ctx.dev_alloc(n_rec * L), ctx.dev_alloc(n_rec * L)
    half = n_pairs * L
    ctx.synth_pairs(d_tx, d_off, d_thr, len(tx), n_pairs, 0, seed, L, 300, synth.ERR_THRESHOLD_24,
                    d1, d2 + half)
    ctx.revcomp_reads(d2 + half, d1 + half, n_pairs, L)
    ctx.revcomp_reads(d1, d2, n_pairs, L)
    dk, dc, nk = ctx.count_k1mers([d1, d2], [n_rec, n_rec], L, K1, 2 * n_rec * 76 // 3)
    h_offs = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(L)
    d_offs = ctx.to_device(h_offs)
    mates = [(d1, d_offs, n_rec, True), (d2, d_offs, n_rec, True)]
    out = []
    for _ in range(2):
        cor, comp_offs, rec_idx, stats = pipeline.frontend_in_memory(
            ctx, dk, dc, K1, mates, True, 3, 75, 500, True, nk)
        walks = ctx.l3_walks()
        allowed = ctx.l3_allowed()
        out.append((cor, comp_offs, rec_idx.copy(), stats, walks, allowed))   # rec_idx: view of a re-used pinned buffer
    reads = (ctx.d2h(np.empty((n_rec, L), np.uint8), d1), ctx.d2h(np.empty((n_rec, L), np.uint8), d2))
    keys = ctx.d2h(np.empty(nk, np.uint64), dk)
    counts = ctx.d2h(np.empty(nk, np.uint32), dc)
    yield ctx, out, reads, keys, counts
    ctx.close()


def test_deterministic_across_runs(run):
    _, out, _, _, _ = run
    a, b = out
    assert a[0].contigs.strings() == b[0].contigs.strings()
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for x, y in zip(a[4], b[4]):
        assert np.array_equal(x, y)
    assert a[3]["n_traversed"] == b[3]["n_traversed"]


def test_walk_conservation_and_order(run):
    ctx, out, _, keys, counts = run
    cor, _, _, stats, (seed, nl, nr, tot, flags), _ = out[0]
    # every traversed K1-mer belongs to exactly one walk (I1)
    assert int(nl.sum() + nr.sum() + len(seed)) == stats["n_traversed"] <= stats["n_loaded"]
    # pop order: seed weights never increase along the walk list; all >= min_weight (I3)
    w, f = ctx.table_lookup(seed)
    assert f.all() and w.min() >= 3
    assert np.all(np.diff(w.astype(np.int64)) <= 0)
    # a walk's weight is at least its seed weight plus one per extra K1-mer
    assert np.all(tot >= w.astype(np.uint64) + nl + nr)
    # accepted => passes shape and not duplicate
    assert np.all((flags & 4 == 0) | ((flags & 1 == 1) & (flags & 2 == 0)))
    assert int((flags & 4 != 0).sum()) == stats["n_contigs"]


def test_contigs_are_paths_of_the_table_and_disjoint(run):
    ctx, out, _, keys, counts = run
    cor, _, _, stats, _, (ak, aw) = out[0]
    # allowed K1-mers: all distinct (I1), all in the table with the same weight (a7)
    assert len(np.unique(ak)) == len(ak) == stats["n_allowed"]
    w, f = ctx.table_lookup(ak)
    assert f.all() and np.array_equal(w, aw)
    # and they are exactly the windows of the accepted contigs, in contig order
    from shannon_b200.extension_correction import encode_kmer
    some = cor.contigs.strings()[:39]
    exp = [encode_kmer(c[i:i + K1]) for c in some for i in range(len(c) - K1 + 1)]
    assert ak[:len(exp)].tolist() == exp
    assert all(len(c) >= 75 for c in cor.contigs.strings())
    # low-complexity K1-mers never enter the table (a2)
    lowc = encode_kmer("A" * 23 + "CG")
    assert ctx.table_lookup(np.asarray([lowc], np.uint64))[1].tolist() == [0]


def test_read_partition_properties(run):
    ctx, out, reads, _, _ = run
    cor, comp_offs, rec_idx, stats, _, (ak, _) = out[0]
    n_comps = stats["n_partitions"]
    assert comp_offs[0] == 0 and comp_offs[-1] == len(rec_idx) == stats["assignments"]
    allowed = set(ak.tolist())
    from shannon_b200.extension_correction import encode_kmer
    # per component: record indices strictly ascending (input order preserved, no duplicates)
    for c in range(n_comps):
        seg = rec_idx[comp_offs[c]:comp_offs[c + 1]].astype(np.int64)
        assert np.all(np.diff(seg) > 0)
    # spot check: an assigned record has a sampled K1-mer among the allowed ones; records whose
    # sampled K1-mers are all absent are assigned nowhere
    assigned = np.zeros(reads[0].shape[0], bool)
    assigned[rec_idx] = True
    rng = np.random.default_rng(3)
    for r in rng.integers(0, reads[0].shape[0], size=300).tolist():
        hit = False
        for m in reads:
            s = bytes(m[r]).decode()
            for st in (0, 25, 50, 75):
                hit |= encode_kmer(s[st:st + K1]) in allowed
        # contigs of single-contig components are not partitioned, so a hit need not be assigned,
        # but an assignment always implies a hit
        assert (not assigned[r]) or hit
    assert stats["lookups"] == 8 * stats["valid_records"] == 8 * reads[0].shape[0]


def test_speculative_walks_equal_serial_walks_at_scale(run, monkeypatch):
    """The speculative windowed kernel (racy by construction, exact by in-order commit) must
    reproduce the one-warp-per-component replay bit for bit on every walk of the 1 M-pair input,
    whichever components it is applied to."""
    ctx, out, _, keys, counts = run
    results = []
    for min_nodes in ("1000000000", "1", "20000"):
        monkeypatch.setenv("SHN_SPEC_MIN_NODES", min_nodes)
        ctx.table_build(keys, counts, K1, False)
        sz = ctx.l3_run(3, 75)
        results.append((sz, ctx.l3_walks(), ctx.l3_contigs()))
    assert results[0][0]["n_spec_comps"] == 0 and results[1][0]["n_spec_comps"] > 0
    for sz, walks, (bases, offs) in results[1:]:
        assert sz["n_traversed"] == results[0][0]["n_traversed"]
        for x, y in zip(walks, results[0][1]):
            assert np.array_equal(x, y)
        assert np.array_equal(bases, results[0][2][0]) and np.array_equal(offs, results[0][2][1])


def test_speculative_walks_equal_serial_walks_wide_keys(monkeypatch):
    """Same check for K = 32 (128-bit keys, two slots per bucket) at 300 k read pairs: the serial
    one-warp replay, both speculative tiers on every component, and the default mix agree on every
    walk and every contig."""
    ctx = _lib.Context(0)
    try:
        n_pairs, n_tx, seed, k1 = 300000, 150, 91, 33
        tx = synth.make_transcripts(n_tx, seed)
        codes, offs = synth.pack_transcripts(tx)
        thr = synth.expression_thresholds(len(tx), [len(t) for t in tx], True)
        d_tx, d_off, d_thr = ctx.to_device(codes), ctx.to_device(offs), ctx.to_device(thr)
        n_rec = 2 * n_pairs
        d1, d2 = ctx.dev_alloc(n_rec * L), ctx.dev_alloc(n_rec * L)
        half = n_pairs * L
        ctx.synth_pairs(d_tx, d_off, d_thr, len(tx), n_pairs, 0, seed, L, 300, synth.ERR_THRESHOLD_24,
                        d1, d2 + half)
        ctx.revcomp_reads(d2 + half, d1 + half, n_pairs, L)
        ctx.revcomp_reads(d1, d2, n_pairs, L)
        dk, dc, nk = ctx.count_k1mers([d1, d2], [n_rec, n_rec], L, k1, 2 * n_rec * (L - k1 + 1) // 3)
        results = []
        for min_nodes, tier16 in (("1000000000", "0"), ("1", "3"), ("20000", "1")):
            monkeypatch.setenv("SHN_SPEC_MIN_NODES", min_nodes)
            monkeypatch.setenv("SHN_SPEC_TIER16", tier16)
            ctx.table_build(dk, dc, k1, False, on_device=True, n=nk)
            sz = ctx.l3_run(3, 75)
            results.append((sz, ctx.l3_walks(), ctx.l3_contigs()))
        assert results[0][0]["n_spec_comps"] == 0 and results[1][0]["n_spec_comps"] > 3
        assert results[0][0]["n_contigs"] > 50 and results[0][1][0].shape[1] == 2
        for sz, walks, (bases, offs_c) in results[1:]:
            assert sz["n_traversed"] == results[0][0]["n_traversed"]
            for x, y in zip(walks, results[0][1]):
                assert np.array_equal(x, y)
            assert np.array_equal(bases, results[0][2][0]) and np.array_equal(offs_c, results[0][2][1])
    finally:
        ctx.close()
