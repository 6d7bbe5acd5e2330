"""Hash-sharded K1-mer table over the GPUs of one box (north_star: "the k-mer tables
shard naturally by k-mer hash across the 8 GPUs, with an NCCL all-to-all over NVLink routing k-mer
batches to owner ranks").

One process per GPU; ``torch.distributed`` is the plumbing (``all_to_all_single`` over NCCL /
NVLink; gloo in the CPU tests).  Every key lives on ``owner(key)`` = low 32 bits of fmix64(key)
scaled to the rank count -- independent of the in-table bucket hash, which uses the high bits.

  build : partition the local slice of the input lines by owner  -> all-to-all (key, count, global
          line index) -> shn_table_build_indexed on the owner
  lookup: partition the queries by owner -> all-to-all keys -> local probe -> all-to-all answers
          back -> un-permute into query order

The exchange protocol (split sizes, permutation, un-permutation) is the same code for both
backends; only the four array primitives differ (``GpuOps`` below: kernels of libshannon_b200.so;
the gloo test supplies a numpy twin).  Round 1 status: the sharded table serves distributed
lookups (half of the headline metric); the walks still run per rank on rank-local tables -- the
component-wise re-sharding that makes them global is described in DESIGN.md section 8.
"""
import numpy as np
import torch
import torch.distributed as dist


class GpuOps(object):
    """Array primitives on the device of one shn context (tensors are CUDA tensors)."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = torch.device("cuda", device)

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.sync()

    def plan(self, keys, world):
        """stable partition by owner: (perm int32 tensor, per-rank counts)."""
        n = keys.numel()
        perm = self.empty(n, torch.int32)
        self.sync()
        counts = self.ctx.route_plan(keys.data_ptr(), n, world, perm.data_ptr())
        return perm, counts

    def gather(self, src, perm):
        out = self.empty(perm.numel(), src.dtype)
        self.sync()
        self.ctx.permute(src.data_ptr(), perm.data_ptr(), perm.numel(), src.element_size(),
                         out.data_ptr(), scatter=False)
        return out

    def scatter(self, src, perm):
        out = self.empty(perm.numel(), src.dtype)
        self.sync()
        self.ctx.permute(src.data_ptr(), perm.data_ptr(), perm.numel(), src.element_size(),
                         out.data_ptr(), scatter=True)
        return out

    def build(self, keys, counts, line_idx, k1):
        self.sync()
        self.ctx.table_build_indexed(keys.data_ptr(), counts.data_ptr(), line_idx.data_ptr(),
                                     keys.numel(), k1)

    def lookup(self, keys):
        n = keys.numel()
        w = self.empty(n, torch.int32)
        f = self.empty(n, torch.uint8)
        self.sync()
        if n:
            self.ctx.table_lookup_dev(keys.data_ptr(), n, w.data_ptr(), f.data_ptr())
        return w, f


def all_to_all_v(ops, send, send_counts, group=None):
    """Variable all-to-all of a 1-D tensor that is already contiguous per destination rank.
    Returns (recv tensor, recv_counts)."""
    world = dist.get_world_size(group)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = ops.empty(sum(recv_counts), send.dtype)
    # zero-size splits are legal; give empty tensors a valid storage
    dist.all_to_all_single(recv if recv.numel() else ops.empty(0, send.dtype),
                           send if send.numel() else ops.empty(0, send.dtype),
                           output_split_sizes=recv_counts, input_split_sizes=list(send_counts),
                           group=group)
    return recv, recv_counts


class ShardedKmerTable(object):
    """K1-mer -> weight table sharded by key hash over the ranks of ``group``.
    Keys travel as int64 bit patterns of the packed uint64 keys, counts/weights as int32."""

    def __init__(self, ops, group=None):
        self.ops = ops
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_local_lines = 0

    def build(self, keys, counts, first_line, k1):
        """keys/counts: this rank's slice of the input lines (line i of the slice is global line
        first_line + i: the dict insertion order that breaks seed ties)."""
        ops = self.ops
        n = keys.numel()
        perm, send_counts = ops.plan(keys, self.world)
        line = (torch.arange(n, dtype=torch.int64, device=keys.device) + int(first_line)).to(torch.int32)
        rk, _ = all_to_all_v(ops, ops.gather(keys, perm), send_counts, self.group)
        rc, _ = all_to_all_v(ops, ops.gather(counts, perm), send_counts, self.group)
        rl, _ = all_to_all_v(ops, ops.gather(line, perm), send_counts, self.group)
        self.n_local_lines = rk.numel()
        ops.build(rk, rc, rl, k1)
        return self.n_local_lines

    def lookup(self, keys):
        """weights (int32, 0 if absent) and found flags (uint8) in query order."""
        ops = self.ops
        perm, send_counts = ops.plan(keys, self.world)
        rk, recv_counts = all_to_all_v(ops, ops.gather(keys, perm), send_counts, self.group)
        w, f = ops.lookup(rk)
        bw, _ = all_to_all_v(ops, w, recv_counts, self.group)
        bf, _ = all_to_all_v(ops, f, recv_counts, self.group)
        return ops.scatter(bw, perm), ops.scatter(bf, perm)


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of n items for `rank` (input lines / read records by range)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---- read partition sharded by record range (second regime) ---------------------
def _bcast_array(a, src, device, group=None):
    """numpy array of rank `src` -> the same array on every rank (tensor broadcast; the shape and
    dtype travel as a small object first)."""
    meta = [None if a is None else (a.shape, a.dtype.str)]
    dist.broadcast_object_list(meta, src=src, group=group)
    shape, dtype = meta[0]
    if dist.get_rank(group) != src:
        a = np.empty(shape, dtype=np.dtype(dtype))
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
    if device is not None:
        t = t.to(device)
    if t.numel():
        dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy().view(np.dtype(dtype)).reshape(shape)


def merge_partitions(per_rank, n_comps):
    """[(comp_offsets, record_idx)] in rank order -> one (comp_offsets, record_idx): inside every
    component the records of rank 0 come first, then rank 1, ...  Ranks own ascending contiguous
    record ranges, so this is the global input order the reference appends reads in
    (kmers_for_component.py:345-351)."""
    sizes = np.stack([np.diff(np.asarray(o, dtype=np.int64)) for o, _ in per_rank])  # [world, n_comps]
    total = sizes.sum(axis=0)
    goff = np.zeros(n_comps + 1, dtype=np.int64)
    goff[1:] = np.cumsum(total)
    before = np.cumsum(sizes, axis=0) - sizes                                        # earlier ranks
    out = np.empty(int(goff[-1]), dtype=np.uint32)
    for r, (offs, idx) in enumerate(per_rank):
        offs = np.asarray(offs, dtype=np.int64)
        n = int(offs[-1])
        if n == 0:
            continue
        comp_of_entry = np.repeat(np.arange(n_comps, dtype=np.int64), sizes[r])
        dst = goff[comp_of_entry] + before[r][comp_of_entry] + (np.arange(n, dtype=np.int64) - offs[comp_of_entry])
        out[dst] = np.asarray(idx[:n], dtype=np.uint32)
    return goff, out


def partition_reads_sharded(ctx, mates, paired, k1, contigs=None, n_comps=None, src=0, device=None,
                            group=None):
    """K1-mer -> component map replicated on every rank, reads sharded by record range, no
    communication during the lookups.

    mates : [(bases uint8, offsets uint64)] host arrays of the WHOLE read files on every rank
            (each rank packs and looks up only its own record range)
    contigs : on rank `src` (bases uint8 ASCII, offsets uint64, comp_of_contig uint32), the
            contigs of the partition and their component ids (SHN none = 0xFFFFFFFF is skipped);
            ignored elsewhere
    Returns (comp_offsets int64, record_idx uint32) on rank `src` -- identical to
    pipeline.partition_reads on one GPU -- and (None, None) on the other ranks."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cb, co, cc = contigs if rank == src else (None, None, None)
    cb = _bcast_array(cb, src, device, group)
    co = _bcast_array(co, src, device, group)
    cc = _bcast_array(cc, src, device, group)
    nc = [n_comps]
    dist.broadcast_object_list(nc, src=src, group=group)
    n_comps = int(nc[0])
    lens = np.diff(co.astype(np.int64))
    ctx.l4_map_add_contigs(cb, co, cc, k1, True, int(np.maximum(lens - k1 + 1, 0).sum()))
    n_rec = len(mates[0][1]) - 1
    lo, hi = shard_range(n_rec, rank, world)
    for m, (bases, offs) in enumerate(mates):
        o = np.asarray(offs, dtype=np.uint64)
        ctx.l4_load_reads(m, bases[int(o[lo]):int(o[hi])], o[lo:hi + 1] - o[lo])
    n_assign, _, _ = ctx.l4_assign(paired, k1)
    offs, idx = ctx.l4_assignments(n_comps, n_assign)
    idx = idx.astype(np.uint32) + np.uint32(lo)
    gathered = [None] * world if rank == src else None
    dist.gather_object((offs.astype(np.int64), idx), gathered, dst=src, group=group)
    if rank != src:
        return None, None
    return merge_partitions(gathered, n_comps)
