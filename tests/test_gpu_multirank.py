"""2-GPU test (NCCL) of the hash-sharded K1-mer table: device routing kernels + all-to-all +
indexed build + distributed lookups, checked against the numpy twin.  Skipped with < 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import dist_testlib
    from shannon_b200 import _lib
    from shannon_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(9)
        n, k1 = 300000, 25
        keys = rng.integers(0, 1 << (2 * k1), size=n, dtype=np.uint64)
        keys[::11] = keys[5]
        # keep only keys that survive the low-complexity filter trivially (random 25-mers do)
        counts = rng.integers(1, 90, size=n).astype(np.int32)
        lo, hi = sdist.shard_range(n, rank, world)
        ctx = _lib.Context(rank)
        ops = sdist.GpuOps(ctx, rank)
        tab = sdist.ShardedKmerTable(ops)
        dk = torch.from_numpy(keys[lo:hi].view(np.int64).copy()).cuda()
        dc = torch.from_numpy(counts[lo:hi].copy()).cuda()
        # device routing plan == numpy twin
        perm, cnts = ops.plan(dk, world)
        own = dist_testlib.owner_of(keys[lo:hi], world)
        assert cnts == np.bincount(own, minlength=world).tolist()
        assert perm.cpu().numpy().tolist() == np.argsort(own, kind="stable").tolist()
        tab.build(dk, dc, lo, k1)
        exp = {}
        for i, (k, c) in enumerate(zip(keys.tolist(), counts.tolist())):
            e = exp.setdefault(k, [0, i])
            e[0] += c
        mine = dict((k, v) for k, v in exp.items()
                    if dist_testlib.owner_of(np.array([k], dtype=np.uint64), world)[0] == rank)
        gk, gw, gi = ctx.table_dump()
        assert dict(zip(gk.tolist(), zip(gw.tolist(), gi.tolist()))) == \
            dict((k, (v[0], v[1])) for k, v in mine.items())
        q = np.concatenate([keys[rng.integers(0, n, size=200000)],
                            rng.integers(0, 1 << 50, size=100000, dtype=np.uint64)])
        np.random.default_rng(rank).shuffle(q)
        w, f = tab.lookup(torch.from_numpy(q.view(np.int64).copy()).cuda())
        ops.sync()
        assert w.cpu().tolist() == [exp.get(k, [0])[0] for k in q.tolist()]
        assert f.cpu().tolist() == [int(k in exp) for k in q.tolist()]
        ctx.close()
        open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_sharded_table_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def _partition_worker(rank, world, port, out_dir):
    import dist_testlib
    from shannon_b200 import _lib
    from shannon_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        contigs, n_comps, mates = dist_testlib.partition_case(seed=4, n_tx=12, n_pairs=20000)
        ctx = _lib.Context(rank)
        offs, idx = sdist.partition_reads_sharded(ctx, mates, True, 25,
                                                  contigs if rank == 0 else None,
                                                  n_comps if rank == 0 else None,
                                                  device=torch.device("cuda", rank))
        if rank == 0:
            # the same partition on one GPU over all the records, and the oracle's loop
            lens = np.diff(contigs[1].astype(np.int64))
            ctx.l4_map_add_contigs(contigs[0], contigs[1], contigs[2], 25, True,
                                   int(np.maximum(lens - 24, 0).sum()))
            for m, (b, o) in enumerate(mates):
                ctx.l4_load_reads(m, b, o)
            na, _, _ = ctx.l4_assign(True, 25)
            eo, ei = ctx.l4_assignments(n_comps, na)
            assert offs.tolist() == eo.astype(np.int64).tolist() and np.array_equal(idx, ei) and na > 5000
            ref = dist_testlib.OracleL4Ctx()
            ref.l4_map_add_contigs(contigs[0], contigs[1], contigs[2], 25, True, 10 ** 9)
            for m, (b, o) in enumerate(mates):
                ref.l4_load_reads(m, b, o)
            assert ref.l4_assign(True, 25)[0] == na
            ro, ri = ref.l4_assignments(n_comps, na)
            assert ro.tolist() == eo.tolist() and ri.tolist() == ei.tolist()
        ctx.close()
        open(os.path.join(out_dir, "pok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_sharded_read_partition_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_partition_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["pok0", "pok1"]


def _sharded_worker(rank, world, port, out_dir, case_dir):
    """The whole front end over NCCL (sharded.TorchComm) against the single-GPU path."""
    from shannon_b200 import _lib, sharded
    from shannon_b200 import dist as sdist
    import test_gpu_sharded as tgs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        s1, s2 = helpers.synthetic_seqs(40, 20000, 71)
        case = helpers.make_case(os.path.join(case_dir, "r%d" % rank), 24, s1, s2, fast_count=True)
        ctx = _lib.Context(rank)
        keys, counts, k1, mates = tgs.load_case_arrays(ctx, case)
        ref = None
        if rank == 0:
            cor, offs, idx, stats = __import__("shannon_b200.pipeline", fromlist=["x"]).frontend_in_memory(
                ctx, keys, counts, k1, [(b, o, None, False) for b, o in mates], True, 3, 75, 4)
            ref = tgs.snapshot(ctx, cor, offs, idx.copy(), stats)
        ops = sharded.GpuOps(ctx, dev)
        comm = sharded.TorchComm(device=dev)
        lo, hi = sdist.shard_range(len(counts), rank, world)
        d_keys = ctx.to_device(np.ascontiguousarray(keys[lo:hi]))
        d_counts = ctx.to_device(np.ascontiguousarray(counts[lo:hi]))
        n_rec = len(mates[0][1]) - 1
        rlo, rhi = sdist.shard_range(n_rec, rank, world)
        mine = []
        for b, o in mates:
            o = np.asarray(o, dtype=np.uint64)
            mine.append((np.ascontiguousarray(b[int(o[rlo]):int(o[rhi])]),
                         np.ascontiguousarray(o[rlo:rhi + 1] - o[rlo]), None, False))
        for _ in range(2):     # twice: the path is re-runnable on the same contexts
            cor, offs, idx, stats = sharded.frontend_sharded(comm, ops, ctx, d_keys, d_counts, hi - lo, lo, k1,
                                                             mine, rlo, True, 3, 75, 4)
        got = tgs.snapshot(ctx, cor, offs, idx, stats)
        if rank == 0:
            tgs.assert_same_result(ref, got, "NCCL world %d" % world)
            assert stats["cross_edges"] > 0 and len(ref["idx"]) > 5000 and comm.bytes_sent > 0
        ops.close()
        ctx.close()
        open(os.path.join(out_dir, "shok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_sharded_frontend_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    case_dir = str(tmp_path / "cases")
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path), case_dir), nprocs=world, join=True)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("shok")) == ["shok%d" % r for r in range(world)]
