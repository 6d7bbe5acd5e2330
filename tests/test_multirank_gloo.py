"""world_size-2 gloo tests (CPU) of the N>1 host logic: routing plan, variable all-to-all,
sharded build / lookup protocol of shannon_b200.dist, and the range sharding bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)
from shannon_b200 import dist as sdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_input(seed=5, n=4000, k1=25):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 1 << (2 * k1), size=n, dtype=np.uint64)
    keys[::7] = keys[3]                 # repeated lines accumulate; first line index is the minimum
    counts = rng.integers(1, 50, size=n).astype(np.int32)
    return keys, counts


def _worker(rank, world, port, out_dir):
    import dist_testlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys, counts = _make_input()
        lo, hi = sdist.shard_range(len(keys), rank, world)
        ops = dist_testlib.NumpyOps()
        tab = sdist.ShardedKmerTable(ops)
        n_local = tab.build(torch.from_numpy(keys[lo:hi].view(np.int64).copy()),
                            torch.from_numpy(counts[lo:hi].copy()), lo, 25)
        # every key landed on its owner, with summed counts and the global first line
        own = dist_testlib.owner_of(np.array(list(ops.table), dtype=np.uint64), world) \
            if ops.table else np.zeros(0)
        assert (own == rank).all()
        exp = {}
        for i, (k, c) in enumerate(zip(keys.tolist(), counts.tolist())):
            e = exp.setdefault(k, [0, i])
            e[0] += c
        mine = dict((k, v) for k, v in exp.items()
                    if dist_testlib.owner_of(np.array([k], dtype=np.uint64), world)[0] == rank)
        assert ops.table == mine
        tot = torch.tensor([n_local])
        dist.all_reduce(tot)
        assert int(tot) == len(keys)
        # lookups from this rank: present keys (owned by anyone), absent keys, ragged batch size
        rng = np.random.default_rng(100 + rank)
        q = np.concatenate([keys[rng.integers(0, len(keys), size=500 + 37 * rank)],
                            rng.integers(0, 1 << 50, size=300, dtype=np.uint64)])
        rng.shuffle(q)
        w, f = tab.lookup(torch.from_numpy(q.view(np.int64).copy()))
        assert w.tolist() == [exp.get(k, [0])[0] for k in q.tolist()]
        assert f.tolist() == [int(k in exp) for k in q.tolist()]
        # an empty query batch on one rank must not dead-lock the collective
        qe = q[:0] if rank == 0 else q[:5]
        w, f = tab.lookup(torch.from_numpy(qe.view(np.int64).copy()))
        assert len(w) == len(qe)
        open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_sharded_table_protocol_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1001):
        for world in (1, 2, 3, 8):
            r = [sdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_owner_hash_is_independent_of_bucket_hash():
    """keys of one shard must still spread over all buckets of that shard's table (the table uses
    the high bits of fmix64, the owner the low 32 bits)."""
    import dist_testlib
    from shannon_b200 import synth
    keys = np.random.default_rng(1).integers(0, 1 << 50, size=200000, dtype=np.uint64)
    own = dist_testlib.owner_of(keys, 8)
    assert np.bincount(own, minlength=8).min() > 200000 / 8 * 0.9
    hi = (synth.mix64(keys[own == 3]) >> np.uint64(54)).astype(np.int64)      # top 10 bits
    assert np.bincount(hi, minlength=1024).min() > 0


def _partition_worker(rank, world, port, out_dir):
    import dist_testlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        contigs, n_comps, mates = dist_testlib.partition_case()
        ctx = dist_testlib.OracleL4Ctx()
        offs, idx = sdist.partition_reads_sharded(ctx, mates, True, 25,
                                                  contigs if rank == 0 else None,
                                                  n_comps if rank == 0 else None)
        if rank == 0:
            # one "rank" over all the records = the reference's own loop order
            ref = dist_testlib.OracleL4Ctx()
            ref.l4_map_add_contigs(contigs[0], contigs[1], contigs[2], 25, True, 10 ** 9)
            for m, (b, o) in enumerate(mates):
                ref.l4_load_reads(m, b, o)
            na, _, _ = ref.l4_assign(True, 25)
            eo, ei = ref.l4_assignments(n_comps, na)
            assert offs.tolist() == eo.tolist() and idx.tolist() == ei.tolist() and na > 100
        else:
            assert offs is None and idx is None
        open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_read_partition_equals_single_rank(tmp_path, world):
    """Reads sharded by record range against a replicated component map: the merged
    per-component lists are the single-rank lists, in input order."""
    mp.spawn(_partition_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok%d" % r for r in range(world)]


def test_merge_partitions_orders_by_rank_inside_components():
    a = (np.array([0, 2, 2, 5]), np.array([1, 3, 0, 2, 4], dtype=np.uint32))
    b = (np.array([0, 1, 3, 3]), np.array([10, 11, 12], dtype=np.uint32))
    offs, idx = sdist.merge_partitions([a, b], 3)
    assert offs.tolist() == [0, 3, 5, 8] and idx.tolist() == [1, 3, 10, 11, 12, 0, 2, 4]
    offs, idx = sdist.merge_partitions([(np.zeros(4, np.int64), np.zeros(0, np.uint32))] * 2, 3)
    assert offs.tolist() == [0, 0, 0, 0] and len(idx) == 0


# ---- the whole L3 stage on sharded tables (shannon_b200/sharded.py) under gloo ----------------------
def _sharded_worker(rank, world, port, out_dir, ds):
    import dist_testlib
    import helpers
    from oracle import shannon_oracle as so
    from shannon_b200 import sharded
    from shannon_b200.pipeline import encode_kmer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        work = os.path.join(out_dir, "case%d" % rank)
        s1, s2 = helpers.synthetic_seqs(6, 500, 77)
        case = helpers.make_case(work, 14, s1, s2, double_stranded=not ds)     # same files on every rank
        lines = [l.split() for l in open(case.k1mer_org)]
        keys = np.array([encode_kmer(k) for k, _ in lines], dtype=np.uint64)
        counts = np.array([int(c) for _, c in lines], dtype=np.int64)
        lo, hi = sdist.shard_range(len(keys), rank, world)
        ops = dist_testlib.OracleShardOps()
        comm = sharded.TorchComm()
        st = {}
        n_loaded = sharded.correct_sharded(comm, ops, keys[lo:hi], counts[lo:hi], hi - lo, lo, 15, ds, 3, 40,
                                           stats=st)
        # the reference's whole seed loop on the un-sharded input
        res = so.run_correction(case.k1mer_org, os.path.join(work, "k1mer.dict"), 3, 40, ds, work, 500,
                                write_files=False)
        assert n_loaded == len(res.kmers)
        assert ops.contigs[1:] == res.contigs[1:], "contigs differ from the sequential loop"
        assert len(res.contigs) > 4
        assert dict(zip(ops.allowed, ops.allowed_w)) == res.allowed_kmer_dict
        assert list(ops.allowed) == list(res.allowed_kmer_dict)
        assert dict((a, dict(b)) for a, b in ops.connections.items()) == \
            dict((a, dict(b)) for a, b in res.connections.items())
        assert st["cross_edges"] > 0 and st["n_super"] > st["n_raw_comps_global"]
        open(os.path.join(out_dir, "sok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ds", [(2, False), (3, False), (2, True)])
def test_sharded_correction_equals_sequential_loop(tmp_path, world, ds):
    """Routing by minimizer owner, cross-rank components, component re-sharding, per-rank walks,
    merged candidates and the replicated accept loop reproduce the reference's sequential seed
    loop (oracle) on every rank -- protocol test with CPU stand-ins for the kernels."""
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path), ds), nprocs=world, join=True)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("sok")) == ["sok%d" % r for r in range(world)]


def test_merge_partitions_device_matches_host_merge():
    from shannon_b200 import sharded
    rng = np.random.default_rng(3)
    n_comps, world = 5, 3
    parts = []
    for r in range(world):
        sizes = rng.integers(0, 6, size=n_comps)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        idx = np.sort(rng.integers(0, 100, size=int(offs[-1]))).astype(np.uint32) + 1000 * r
        parts.append((offs, idx))
    eo, ei = sdist.merge_partitions(parts, n_comps)
    go, gi = sharded.merge_partitions_device([torch.from_numpy(o) for o, _ in parts],
                                             [torch.from_numpy(i.astype(np.int64)) for _, i in parts],
                                             n_comps, torch.device("cpu"))
    assert go.tolist() == eo.tolist() and gi.tolist() == ei.tolist()
