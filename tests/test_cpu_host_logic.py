"""CPU-only checks of everything around the kernels: the C-ABI library loads and exports every
symbol the header declares, the native host IO, the host-side contig ordering (adjacency order +
DFS) and the closed forms the kernels implement, all against the oracle.  No GPU compute."""
import os
import re
import sys

import numpy as np
import pytest

import helpers
from oracle import shannon_oracle as so
from shannon_b200 import _lib
from shannon_b200 import extension_correction as ec
from shannon_b200.weight_updated_graph import weight_updated_graph

ROOT = helpers.ROOT


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "shannon_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(shn_[a-z0-9_]+)\s*\(", text))
    assert len(declared) > 40
    for name in declared:
        assert hasattr(lib, name), "libshannon_b200.so does not export " + name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in lib.shn_version()


def test_no_gpu_means_loud_failure_not_fallback():
    import ctypes
    lib = _lib.load()
    h = ctypes.c_void_p()
    n_dev_ok = lib.shn_create(0, ctypes.byref(h)) == 0
    if n_dev_ok:
        lib.shn_destroy(h)
        pytest.skip("a GPU is present")
    assert b"no CPU fallback" in lib.shn_last_error(None)
    with pytest.raises(_lib.ShnError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "shannon_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp")):
                with open(os.path.join(base, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_parse_kmer_file_matches_oracle_loader(workdir):
    s1, s2 = helpers.synthetic_seqs(6, 300, 5)
    case = helpers.make_case(workdir, 24, s1, s2)
    keys, counts, k1 = _lib.HostIO().parse_kmer_file(case.k1mer_org)
    assert k1 == 25
    with open(case.k1mer_org) as f:
        lines = [l.split() for l in f]
    assert len(lines) == len(keys)
    assert [ec.encode_kmer(k) for k, _ in lines[:2000]] == keys[:2000].tolist()
    assert [int(c) for _, c in lines] == counts.tolist()
    assert ec.decode_kmers(keys[:5], 25).tobytes().decode() == "".join(k for k, _ in lines[:5])


def test_parse_kmer_file_two_word_keys(workdir):
    """K = 32 (shannon.py's largest -K): 33-base k-mers come back as (n, 2) uint64, low word first."""
    s1, s2 = helpers.synthetic_seqs(4, 200, 9)
    case = helpers.make_case(workdir, 32, s1, s2)
    keys, counts, k1 = _lib.HostIO().parse_kmer_file(case.k1mer_org)
    with open(case.k1mer_org) as f:
        lines = [l.split() for l in f]
    assert k1 == 33 and keys.shape == (len(lines), 2) and keys.dtype == np.uint64
    assert ec.keys_as_ints(keys) == [ec.encode_kmer(k) for k, _ in lines]
    assert np.array_equal(ec.keys_array(ec.keys_as_ints(keys), 33), keys)
    assert [int(c) for _, c in lines] == counts.tolist()
    txt = ec.decode_kmers(keys, 33).tobytes().decode()
    assert [txt[i:i + 33] for i in range(0, len(txt), 33)] == [k for k, _ in lines]
    d = ec.AllowedKmerDict(keys[:50], counts[:50], 33)
    assert dict(d) == dict((k, int(c)) for k, c in lines[:50])
    assert d.get(lines[7][0]) == int(lines[7][1]) and d.get("A" * 33, 0) == 0
    p = os.path.join(workdir, "k34")
    with open(p, "w") as f:
        f.write("A" * 34 + "\t1\n")
    with pytest.raises(_lib.ShnError):
        _lib.HostIO().parse_kmer_file(p)


def test_load_fasta_quirks(workdir):
    io = _lib.HostIO()
    p = os.path.join(workdir, "r.fa")
    with open(p, "w") as f:
        f.write(">a\nACGT\n>b\nACNT\n>c\n\n>d\nAAAA\n")
    b, o = io.load_fasta(p)
    # the record with the empty read is kept and ends the input (kmers_for_component.py:338-339,355)
    assert o.tolist() == [0, 4, 8, 8] and b.tobytes() == b"ACGTACNT"
    with open(p, "w") as f:
        f.write(">a\nACGT\n\n>b\nACGT\n")
    b, o = io.load_fasta(p)
    assert o.tolist() == [0, 4]            # empty name line stops the reader
    b, o = io.load_fasta(p, 3)             # lock-step mate: fixed count, padded
    assert len(o) == 4


def _edges_from_oracle(res):
    """distinct edges (a<b, w, fp) from the oracle's contig_connections, fp recomputed from the
    contig strings (first C-mer position in b shared with a)."""
    c_len = res.k1 - 1
    a_l, b_l, w_l, fp_l = [], [], [], []
    for b, nbrs in res.connections.items():
        cb = res.contigs[b]
        for a, w in nbrs.items():
            if a < b:
                ca = res.contigs[a]
                cm = set(ca[i:i + c_len] for i in range(len(ca) - c_len + 1))
                fp = next(i for i in range(len(cb) - c_len + 1) if cb[i:i + c_len] in cm)
                a_l.append(a), b_l.append(b), w_l.append(w), fp_l.append(fp)
    order = np.lexsort((b_l, a_l)) if a_l else np.empty(0, np.int64)
    f = lambda x: np.asarray(x, dtype=np.uint32)[order]
    return f(a_l), f(b_l), f(w_l), f(fp_l)


def _oracle_result(workdir, seed, K=12, psize=2):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    reads = cases.repeat_rich_reads(seed, 500, 40, 200, 5)
    case = helpers.make_case(workdir, K, reads)
    out = case.outdir("o")
    return so.run_correction(case.k1mer_org, out + "/k", 2, 25, False, out, psize, True, True)


@pytest.mark.parametrize("seed", range(6))
def test_host_adjacency_and_dfs_reproduce_reference_order(workdir, seed):
    res = _oracle_result(workdir, seed)
    n = len(res.contigs) - 1
    a, b, w, fp = _edges_from_oracle(res)
    adj = ec.contig_adjacency(n, a, b, w, fp)
    for x in range(1, n + 1):
        assert adj[x] == list(res.connections[x].items()), "neighbour order of contig %d" % x
    comp, comp_of = ec.dfs_components(n, adj)
    assert list(comp.items()) == list(res.component2contig.items())
    assert any(len(v) > 2 for v in comp.values()), "case too simple to pin the DFS order"


def _pair_table(contigs, j, accepted_before, r):
    """brute-force (count, last_i, covered) of candidate j against every earlier accepted d --
    the quantities selfjoin.cu computes."""
    cj = contigs[j]
    rows = []
    for d in accepted_before:
        cd = contigs[d]
        occ = {}
        for p in range(len(cd) - r + 1):
            occ[cd[p:p + r]] = occ.get(cd[p:p + r], 0) + 1
        hits = [(i, occ[cj[i:i + r]]) for i in range(len(cj) - r + 1) if cj[i:i + r] in occ]
        if hits:
            cov = set()
            for i, _ in hits:
                cov.update(range(i, i + r))
            rows.append((d, sum(m for _, m in hits), hits[-1][0], len(cov)))
    return rows


@pytest.mark.parametrize("seed", range(6))
def test_duplicate_closed_form_equals_reference_loop(workdir, seed):
    """dup_round_kernel's rule: best = argmax over accepted partners of (count, last_i, id);
    duplicate iff 2*covered(best) > len.  Checked against duplicate_check() on every walk."""
    if seed < 3:
        res = _oracle_result(workdir, seed, K=10 + seed % 3)
    else:   # error bubbles of simulated reads: most short walks duplicate an accepted contig
        s1, s2 = helpers.synthetic_seqs(10, 1200, seed)
        case = helpers.make_case(workdir, 24, s1, s2)
        out = case.outdir("o")
        res = so.run_correction(case.k1mer_org, out + "/k", 3, 75, False, out, 500, True, True)
    contigs = [w.contig for w in res.walks]
    accepted = []
    n_dup = 0
    for j, wk in enumerate(res.walks):
        rows = _pair_table(contigs, j, accepted, so.R_MER)
        dup = False
        if rows:
            best = max(rows, key=lambda t: (t[1], t[2], t[0]))
            dup = 2 * best[3] > len(wk.contig)
        assert dup == wk.duplicate, "walk %d" % j
        n_dup += dup
        if wk.accepted:
            accepted.append(j)
    assert seed < 3 or n_dup > 0


def _frontier_resolution(cands, contigs, n_blocks, r):
    """CPU model of the duplicate filter as the GPU runs it (l3.cu: l3_filter + dup_round_kernel,
    selfjoin.cu): candidates in rank blocks; a block's pair table holds the partners d < j that are in
    the block or ACCEPTED in an earlier block; inside a block synchronous frontier rounds resolve every
    candidate whose `best` can no longer change.  Returns {candidate: 1 accepted / 2 rejected}, rounds."""
    occ = {}

    def rows_of(j, partners):
        cj = contigs[j]
        rows = []
        for d in partners:
            if d not in occ:
                cd = contigs[d]
                o = occ[d] = {}
                for p in range(len(cd) - r + 1):
                    o[cd[p:p + r]] = o.get(cd[p:p + r], 0) + 1
            o = occ[d]
            hits = [(i, o[cj[i:i + r]]) for i in range(len(cj) - r + 1) if cj[i:i + r] in o]
            if hits:
                cov = set()
                for i, _ in hits:
                    cov.update(range(i, i + r))
                rows.append((d, sum(m for _, m in hits), hits[-1][0], len(cov)))
        return rows

    status = dict((j, 0) for j in cands)
    size = max(1, (len(cands) + n_blocks - 1) // n_blocks)
    total_rounds = 0
    for b0 in range(0, len(cands), size):
        block = cands[b0:b0 + size]
        lo = block[0]
        table = dict((j, rows_of(j, [d for d in cands if d < j and (d >= lo or status[d] == 1)])) for j in block)
        for rounds in range(len(block) + 2):
            new = dict(status)
            unresolved = 0
            for j in block:
                if status[j]:
                    continue
                have = have_u = u_after = False
                best = (0, 0, 0)
                u = (0, 0)
                for d, cnt, last, cov in table[j]:          # ascending d, like the pair table
                    sd = status[d]
                    if sd == 2:
                        continue
                    if sd == 1:
                        if not have or cnt > best[0] or (cnt == best[0] and last >= best[1]):
                            have, best, u_after = True, (cnt, last, cov), False
                    else:
                        if not have_u or cnt > u[0] or (cnt == u[0] and last >= u[1]):
                            have_u, u = True, (cnt, last)
                        if have and cnt == best[0] and last == best[1]:
                            u_after = True
                ready = not have_u
                if have_u and have:
                    u_wins = u[0] > best[0] or (u[0] == best[0] and u[1] > best[1]) or \
                        (u[0] == best[0] and u[1] == best[1] and u_after)
                    ready = not u_wins
                if not ready:
                    unresolved += 1
                    continue
                new[j] = 2 if (have and 2 * best[2] > len(contigs[j])) else 1
            status = new
            total_rounds += 1
            if unresolved == 0:
                break
        else:
            raise AssertionError("the frontier rounds do not converge")
    return status, total_rounds


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n_blocks", [1, 4])
def test_duplicate_frontier_rounds_in_rank_blocks_equal_the_sequential_loop(workdir, seed, n_blocks):
    """The parallel resolution order of the GPU's duplicate filter (rank blocks, accepted-only
    partners from earlier blocks, frontier rounds) gives the accept / reject decisions of the
    reference's sequential loop (extension_correction.py:358-361)."""
    if seed < 3:
        res = _oracle_result(workdir, seed, K=10 + seed % 3)
    else:   # low thresholds: hundreds of error-bubble candidates, nearly all duplicates of a few contigs
        s1, s2 = helpers.synthetic_seqs(10, 3000 if seed == 5 else 1200, seed)
        case = helpers.make_case(workdir, 24, s1, s2)
        out = case.outdir("o")
        res = so.run_correction(case.k1mer_org, out + "/k", 2, 40, False, out, 500, True, True)
    contigs = [w.contig for w in res.walks]
    cands = [j for j, w in enumerate(res.walks) if w.passes_shape]   # the GPU's candidates (a5)
    status, rounds = _frontier_resolution(cands, contigs, n_blocks, so.R_MER)
    for j in cands:
        assert (status[j] == 1) == bool(res.walks[j].accepted), "candidate %d" % j
    assert rounds >= n_blocks and (seed < 3 or any(v == 2 for v in status.values()))


@pytest.mark.parametrize("seed", range(4))
def test_cmer_closed_form_equals_reference_loop(workdir, seed):
    """w(a,b) = sum over shared C-mers of occ_a * occ_b."""
    res = _oracle_result(workdir, seed)
    c_len = res.k1 - 1
    occ = {}
    for idx in range(1, len(res.contigs)):
        c = res.contigs[idx]
        o = occ[idx] = {}
        for i in range(len(c) - c_len + 1):
            o[c[i:i + c_len]] = o.get(c[i:i + c_len], 0) + 1
    for b, nbrs in res.connections.items():
        for a, w in nbrs.items():
            assert w == sum(n * occ[a].get(cm, 0) for cm, n in occ[b].items())


def test_shape_filter_matches_python_pow():
    """passes_shape() in l3.cu uses the same libm pow as math.pow; here: the oracle's expression
    is the reference's (extension_correction.py:353,361) on boundary-ish inputs."""
    assert so.passes_shape(150, 3 * 126, 126, 3, 75)
    assert not so.passes_shape(149, 3 * 125, 125, 3, 75)
    assert not so.passes_shape(74, 10 ** 6, 50, 3, 75)


def test_weight_updated_graph_matches_oracle(workdir):
    with open(os.path.join(workdir, "g.txt"), "w") as f:
        f.write("4\t4\t001\n2\t3\t3\t1\t\n1\t3\t4\t2\t\n1\t1\t4\t7\t\n2\t2\t3\t7\t\n")
    with open(os.path.join(workdir, "g.txt.part.2"), "w") as f:
        f.write("0\n0\n1\n1\n")
    weight_updated_graph(workdir, "/g.txt.part.2", "/g.txt", "/a.txt", "/c", "/c", 5, False)
    so.weight_updated_graph(workdir, "/g.txt.part.2", "/g.txt", "/b.txt", 5)
    assert open(workdir + "/a.txt").read() == open(workdir + "/b.txt").read()
    assert "\t5\t" in open(workdir + "/a.txt").read()


def test_allowed_kmer_dict_behaves_like_a_dict():
    kmers = ["ACGTA", "TTGCA", "GGGAC"]
    keys = np.asarray([ec.encode_kmer(k) for k in kmers], dtype=np.uint64)
    d = ec.AllowedKmerDict(keys, np.asarray([5, 7, 9], dtype=np.uint32), 5)
    assert dict(d) == {"ACGTA": 5, "TTGCA": 7, "GGGAC": 9}
    assert list(d) == kmers and len(d) == 3
    assert d.get("TTGCA", 0) == 7 and d.get("AAAAA", 0) == 0 and d.get("ACG", 0) == 0
    assert "GGGAC" in d and "NNNNN" not in d
    d.clear()
    assert len(d) == 0 and dict(d) == {}


def test_synth_generator_is_deterministic_and_strand_consistent():
    from shannon_b200 import synth
    tx = synth.make_transcripts(5, 3)
    codes, offs = synth.pack_transcripts(tx)
    thr = synth.expression_thresholds(len(tx), [len(t) for t in tx])
    a = synth.make_pairs(codes, offs, thr, 50, 9)
    b = synth.make_pairs(codes, offs, thr, 20, 9, first_pair=30)
    assert np.array_equal(a[0][30:], b[0]) and np.array_equal(a[1][30:], b[1])
    r1, r2 = synth.rc_double(a[0], a[1])
    assert r1.shape == (100, 100) and bytes(r2[0]).decode() == helpers.rc_str(bytes(a[0][0]).decode())


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores without a GPU and prints one JSON line with
    the keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "reads_partitioned_per_sec"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and "workload" in line["config"]


def test_assign_components_balances_and_is_deterministic():
    """owner rank of every K1-mer graph component in the sharded path (shannon_b200/sharded.py)"""
    import numpy as np
    from shannon_b200 import sharded
    sizes = np.array([100, 90, 10, 10, 10, 5, 5, 1, 1, 1])
    own = sharded.assign_components(sizes, 3)
    load = np.bincount(own, weights=sizes, minlength=3)
    assert load.max() <= 100 and (own == sharded.assign_components(sizes, 3)).all()
    assert (sharded.assign_components(sizes, 1) == 0).all()
    big = np.concatenate([np.full(5, 1000), np.ones(40000, dtype=np.int64)])
    own = sharded.assign_components(big, 4, exact_top=8)
    load = np.bincount(own, weights=big, minlength=4)
    assert load.max() - load.min() <= 2048
