"""The front end on hash-sharded tables (shannon_b200/sharded.py) against the single-GPU path, on
ONE GPU: several virtual ranks run as threads, each with its own shn context, and exchange through
sharded.ThreadComm -- the same kernels and the same host code as under NCCL, so the single-GPU
test box exercises the whole N-rank data path (minimizer routing, shard build, local and cross-rank
components, component re-sharding, per-rank walks, candidate merge, replicated filters, range-
sharded read partition).  The NCCL variant is tests/test_gpu_multirank.py."""
import threading

import numpy as np
import pytest
import torch

import helpers
from shannon_b200 import _lib, pipeline, sharded
from shannon_b200.dist import shard_range

pytestmark = pytest.mark.gpu


def load_case_arrays(ctx, case):
    keys, counts, k1 = ctx.parse_kmer_file(case.k1mer_org)
    mates = []
    for i, f in enumerate(case.reads_files):
        b, o = ctx.load_fasta(f) if i == 0 else ctx.load_fasta(f, len(mates[0][1]) - 1)
        mates.append((b, o))
    return keys, counts, k1, mates


def single_gpu(keys, counts, k1, mates, paired, partition_size, min_weight=3, min_length=75, ds=False):
    ctx = _lib.Context(0)
    try:
        if ds:
            cor = pipeline.correct(ctx, keys, counts, k1, True, min_weight, min_length)
            return snapshot(ctx, cor, None, None, None)
        cor, offs, idx, stats = pipeline.frontend_in_memory(
            ctx, keys, counts, k1, [(b, o, None, False) for b, o in mates], paired, min_weight, min_length,
            partition_size)
        return snapshot(ctx, cor, offs, idx.copy(), stats)
    finally:
        ctx.close()


def snapshot(ctx, cor, offs, idx, stats):
    ak, aw = ctx.l3_allowed()
    return {"contigs": cor.contigs.strings(), "allowed_keys": np.asarray(ak).copy(),
            "allowed_w": np.asarray(aw).copy(), "edges": [np.asarray(x).copy() for x in ctx.l3_edges()],
            "labels": ctx.l3_labels().copy(), "n_loaded": cor.n_loaded,
            "offs": None if offs is None else np.asarray(offs, dtype=np.int64).copy(),
            "idx": None if idx is None else np.asarray(idx).copy(), "stats": stats}


def run_sharded(world, keys, counts, k1, mates, paired, partition_size, min_weight=3, min_length=75,
                ds=False):
    """`world` virtual ranks on GPU 0; returns the snapshots of every rank."""
    hub = sharded.ThreadHub(world)
    dev = torch.device("cuda", 0)
    results, errors = [None] * world, []
    kw = 2 if k1 > 32 else 1
    n_lines = len(counts)
    n_rec = len(mates[0][1]) - 1 if mates else 0

    def work(rank):
        ctx = ops = None
        try:
            torch.cuda.set_device(0)
            ctx = _lib.Context(0)
            ops = sharded.GpuOps(ctx, dev)
            comm = sharded.ThreadComm(hub, rank, dev)
            lo, hi = shard_range(n_lines, rank, world)
            d_keys = ctx.to_device(np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(-1)[lo * kw:hi * kw]))
            d_counts = ctx.to_device(np.ascontiguousarray(counts[lo:hi]))
            if ds:
                n_loaded = sharded.correct_sharded(comm, ops, d_keys, d_counts, hi - lo, lo, k1, True,
                                                   min_weight, min_length)
                cor = pipeline.collect_correction(ctx, k1, n_loaded)
                results[rank] = snapshot(ctx, cor, None, None, None)
            else:
                rlo, rhi = shard_range(n_rec, rank, world)
                mine = []
                for b, o in mates:
                    o = np.asarray(o, dtype=np.uint64)
                    mine.append((np.ascontiguousarray(b[int(o[rlo]):int(o[rhi])]),
                                 np.ascontiguousarray(o[rlo:rhi + 1] - o[rlo]), None, False))
                cor, offs, idx, stats = sharded.frontend_sharded(
                    comm, ops, ctx, d_keys, d_counts, hi - lo, lo, k1, mine, rlo, paired, min_weight,
                    min_length, partition_size)
                results[rank] = snapshot(ctx, cor, offs, idx, stats)
            ctx.dev_free(d_keys)
            ctx.dev_free(d_counts)
        except BaseException as e:  # noqa: BLE001 -- a dead rank must not leave the others at a barrier
            errors.append((rank, e))
            hub.barrier.abort()
        finally:
            if ops is not None:
                ops.close()
            if ctx is not None:
                ctx.close()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0][1]
    return results


def assert_same_result(a, b, label, partition=True):
    assert a["contigs"] == b["contigs"], label + ": contigs differ"
    assert a["n_loaded"] == b["n_loaded"], label
    assert np.array_equal(a["allowed_keys"], b["allowed_keys"]), label + ": allowed K1-mers differ"
    assert np.array_equal(a["allowed_w"], b["allowed_w"]), label + ": allowed weights differ"
    for x, y in zip(a["edges"], b["edges"]):
        assert np.array_equal(x, y), label + ": contig graph differs"
    assert np.array_equal(a["labels"], b["labels"]), label
    if partition:
        assert np.array_equal(a["offs"], b["offs"]), label + ": component sizes differ"
        assert np.array_equal(a["idx"], b["idx"]), label + ": read partition differs"


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_sharded_equals_single_gpu(workdir, world):
    s1, s2 = helpers.synthetic_seqs(30, 6000, 61)
    s1[5] = s1[5][:30] + "N" + s1[5][31:]          # a dirty read and two short ones
    s2[9] = s2[9][:20]
    s1[40] = s1[40][:26]
    case = helpers.make_case(workdir, 24, s1, s2)
    ctx = _lib.Context(0)
    keys, counts, k1, mates = load_case_arrays(ctx, case)
    ctx.close()
    ref = single_gpu(keys, counts, k1, mates, True, 3)
    assert len(ref["contigs"]) > 20 and ref["stats"]["n_partitions"] > 1 and len(ref["idx"]) > 1000
    res = run_sharded(world, keys, counts, k1, mates, True, 3)
    assert_same_result(ref, res[0], "world %d rank 0" % world)
    for r in range(1, world):                       # the L3 result is replicated on every rank
        assert_same_result(res[0], res[r], "rank %d vs rank 0" % r, partition=False)
        assert res[r]["offs"] is None
    st = res[0]["stats"]
    assert st["n_raw_comps_global"] == ref["stats"]["n_raw_comps"]
    if world > 1:
        assert st["cross_edges"] > 0                # components really span ranks before re-sharding
    # ... and the single-GPU path equals the oracle on this input (the chain of trust)
    from oracle import shannon_oracle as so
    out, allowed, _, ret = helpers.run_frontend(so.extension_correction, so.kmers_for_component, case,
                                                "ora", partition_size=3, inMem=True, repartition=False)
    assert ref["contigs"] == open(out + "/algo_input/k1mer.dict_contig").read().split()


def test_sharded_wide_keys_single_end(workdir):
    """K = 32: 128-bit keys on the wire (32-byte records), single-end reads."""
    s1, _ = helpers.synthetic_seqs(12, 2500, 67)
    case = helpers.make_case(workdir, 32, s1, None)
    ctx = _lib.Context(0)
    keys, counts, k1, mates = load_case_arrays(ctx, case)
    ctx.close()
    assert k1 == 33
    ref = single_gpu(keys, counts, k1, mates, False, 2)
    assert len(ref["contigs"]) > 5
    res = run_sharded(3, keys, counts, k1, mates, False, 2)
    assert_same_result(ref, res[0], "wide keys, world 3")


def test_sharded_double_stranded_load_and_tie_breaks(workdir):
    """-d (every line also adds its reverse complement as the next dict entry) on repeat-rich input
    with many equal weights: the seed tie-break (later input line first) must survive two re-shards."""
    import sys
    import os
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "golden"))
    import cases
    reads = cases.repeat_rich_reads(3, 600, 50, 200, 5)
    case = helpers.make_case(workdir, 12, reads, None, double_stranded=False)
    ctx = _lib.Context(0)
    keys, counts, k1, _ = load_case_arrays(ctx, case)
    ctx.close()
    ref = single_gpu(keys, counts, k1, [], False, 2, min_weight=2, min_length=22, ds=True)
    assert len(ref["contigs"]) >= 3
    for world in (2, 3):
        res = run_sharded(world, keys, counts, k1, [], False, 2, min_weight=2, min_length=22, ds=True)
        assert_same_result(ref, res[0], "-d, world %d" % world, partition=False)


def test_sharded_small_k_and_empty_ranks(workdir):
    """k1 below the minimizer length (owner = hash of the whole key) and more ranks than seeds."""
    reads = ["ACGTTGCAAGGCTTAACCGGTTAGCTAGCTAGGATCCGATCGGATATCGCGCGATTAGCAT" * 2] * 8
    case = helpers.make_case(workdir, 8, reads, None)
    ctx = _lib.Context(0)
    keys, counts, k1, mates = load_case_arrays(ctx, case)
    ctx.close()
    ref = single_gpu(keys, counts, k1, mates, False, 500, min_weight=2, min_length=20)
    res = run_sharded(4, keys, counts, k1, mates, False, 500, min_weight=2, min_length=20)
    assert_same_result(ref, res[0], "k1=9, world 4")


def test_minimizer_owner_matches_numpy_twin():
    """the CPU stand-in of the gloo protocol tests routes like the kernel (csrc/shard.cu owner_of)"""
    import dist_testlib
    rng = np.random.default_rng(17)
    for k1, world in ((25, 8), (15, 3), (32, 5)):
        n = 50000
        hi = (1 << (2 * k1)) - 1
        keys = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
        counts = np.ones(n, dtype=np.uint32)
        ctx = _lib.Context(0)
        d_keys, d_counts = ctx.to_device(keys), ctx.to_device(counts)
        got = ctx.route_lines(d_keys, d_counts, n, 0, False, k1, world)
        exp = np.bincount(dist_testlib.minimizer_owner(keys, k1, world), minlength=world).tolist()
        assert got == exp
        assert min(got) > 0.7 * n / world            # minimizer owners are balanced
        ctx.close()
