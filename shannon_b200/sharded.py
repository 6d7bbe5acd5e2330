"""The whole front end on hash-sharded K1-mer tables, one table shard per rank (SURVEY.md 8e,
north_star: "the k-mer tables shard naturally by k-mer hash across the 8 GPUs of one box, with an
NCCL all-to-all over NVLink routing k-mer batches to owner ranks").

One process per GPU.  Every rank starts with a contiguous slice of the lines of k1mer.dict_org
(extension_correction.py:209-219) and a contiguous range of the read records, and the stages are

  1. route lines to owner(K1-mer) = hash of its minimizer  -- all-to-all #1 -- build the shard
  2. local union-find; successor candidates owned elsewhere are asked for by all-to-all #2; the
     answers are the edges of the graph of local components, all-gathered and labelled on every
     rank: global connected components of the K1-mer successor graph
  3. whole components are assigned to ranks (largest first onto the least loaded rank) and every
     table entry moves to its component's rank -- all-to-all #3 -- second table build
  4. the unchanged single-GPU seed loop / greedy walks / shape filter per rank
     (extension_correction.py:343-356: a walk never leaves its component)
  5. the candidates of all ranks are all-gathered and merged into the global pop order
     (seed weight descending, later input line first, :334); duplicate filter, allowed set,
     contig C-mer graph and its components run replicated on every rank (:358-450, contig-level
     data); the weights of the allowed K1-mers come from their owners by all-reduce
  6. every rank builds the K1-mer -> component map from its replica and partitions ITS range of the
     read records (kmers_for_component.py:322-423); the per-component lists are merged in rank
     order, which is the input order

The collectives are torch.distributed calls (NCCL over NVLink; gloo in the CPU tests) issued on the
stream the library works on, so kernels and collectives are ordered by the stream and the host
only synchronises to read split sizes.  `ThreadComm` runs the same code with several virtual ranks
as threads on ONE GPU (tests on a single-GPU box).
"""
import heapq
import threading
import time

import numpy as np
import torch

from . import pipeline

NONE32 = 0xFFFFFFFF
GLINE_SHIFT = 30


# ---- communicators ------------------------------------------------------------------------------
class TorchComm(object):
    """torch.distributed process group (one process per rank)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cpu")
        self.bytes_sent = 0

    def exchange_ints(self, values):
        """list of ints of this rank -> [list of every rank]"""
        t = torch.tensor(list(values), dtype=torch.int64, device=self.device)
        out = torch.empty(self.world * len(values), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().view(self.world, len(values)).tolist()

    def all_to_all_rows(self, send, send_counts):
        """send: (n, W) tensor whose rows are contiguous per destination rank."""
        sc = torch.tensor(list(send_counts), dtype=torch.int64, device=self.device)
        rc = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self.dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.cpu().tolist()
        n_recv = sum(recv_counts)   # zero-size splits are legal; keep a valid storage behind empty tensors
        recv = torch.empty((max(n_recv, 1),) + tuple(send.shape[1:]), dtype=send.dtype,
                           device=send.device)[:n_recv]
        self.dist.all_to_all_single(recv, send, output_split_sizes=recv_counts,
                                    input_split_sizes=list(send_counts), group=self.group)
        self.bytes_sent += (sum(send_counts) - send_counts[self.rank]) * send.element_size() * \
            int(np.prod(send.shape[1:]))
        return recv, recv_counts

    def all_gather_rows(self, t):
        """rows of every rank concatenated in rank order; returns (tensor, rows per rank)."""
        counts = [c[0] for c in self.exchange_ints([t.shape[0]])]
        mx = max(counts)
        tail = tuple(t.shape[1:])
        pad = torch.empty((mx,) + tail, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * mx,) + tail, dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        self.bytes_sent += (self.world - 1) * t.numel() * t.element_size()
        if all(c == mx for c in counts):
            return out, counts
        return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(self.world)]), counts

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(group=self.group)


class ThreadHub(object):
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class ThreadComm(object):
    """Virtual ranks = threads of one process sharing one device (each with its own shn context and
    its own torch stream); a collective is a pair of barriers around plain tensor copies."""

    def __init__(self, hub, rank, device):
        self.hub, self.rank, self.world, self.device = hub, rank, hub.world, device
        self.bytes_sent = 0

    def _swap(self, obj):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self.hub.slots[self.rank] = obj
        self.hub.barrier.wait()
        return list(self.hub.slots)

    def _done(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self.hub.barrier.wait()

    def exchange_ints(self, values):
        out = [list(v) for v in self._swap(list(values))]
        self._done()
        return out

    def all_to_all_rows(self, send, send_counts):
        everyone = self._swap((send, list(send_counts)))
        recv_counts = [everyone[s][1][self.rank] for s in range(self.world)]
        parts = []
        for s in range(self.world):
            t, sc = everyone[s]
            lo = sum(sc[:self.rank])
            parts.append(t[lo:lo + sc[self.rank]])
        recv = torch.cat(parts) if parts else send[:0]
        self._done()
        return recv, recv_counts

    def all_gather_rows(self, t):
        everyone = self._swap(t)
        out = torch.cat(everyone)
        counts = [x.shape[0] for x in everyone]
        self._done()
        return out, counts

    def all_reduce_sum(self, t):
        everyone = self._swap(t.clone())
        t.copy_(torch.stack(everyone).sum(dim=0))
        self._done()
        return t

    def barrier(self):
        self.hub.barrier.wait()


# ---- device primitives on one shn context -----------------------------------------------------------
class GpuOps(object):
    """The per-rank kernels of libshannon_b200.so on torch CUDA tensors.  The context is switched to
    torch's current stream, so library kernels, torch ops and NCCL collectives are stream-ordered."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        # a dedicated non-default stream becomes this thread's current torch stream (the default
        # stream's handle is 0, which shn_use_stream reads as "back to the context's own stream")
        self._prev = torch.cuda.current_stream(self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.stream.wait_stream(self._prev)
        torch.cuda.set_stream(self.stream)
        ctx.use_stream(self.stream.cuda_stream)

    def close(self):
        self.stream.synchronize()
        self.ctx.use_stream(None)
        torch.cuda.set_stream(self._prev)

    @staticmethod
    def rec_words(k1):
        return 2 if k1 <= 32 else 4

    def _empty(self, shape, dtype):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(shape))
        if n == 0:      # a valid (non-null) storage for zero-size buffers
            return torch.empty(max(int(np.prod(shape[1:])), 1), dtype=dtype, device=self.device)[:0].view(
                (0,) + shape[1:])
        return torch.empty(shape, dtype=dtype, device=self.device)

    def route_lines(self, d_keys, d_counts, n, first_line, double_stranded, k1, world):
        counts = self.ctx.route_lines(d_keys, d_counts, n, first_line, double_stranded, k1, world)
        send = self._empty((sum(counts), self.rec_words(k1)), torch.int64)
        self.ctx.route_lines(d_keys, d_counts, n, first_line, double_stranded, k1, world, counts,
                             send.data_ptr())
        return send, counts

    def build_from_records(self, recs, k1):
        self.ctx.table_build_records(recs.data_ptr(), recs.shape[0], k1)

    def n_distinct(self):
        return self.ctx.table_stats()["n_distinct"]

    def cc_local(self):
        return self.ctx.cc_local()

    def cc_cross(self, world, rank, gid_base, k1):
        counts = self.ctx.cc_cross(world, rank, gid_base)
        send = self._empty((sum(counts), self.rec_words(k1)), torch.int64)
        self.ctx.cc_cross(world, rank, gid_base, counts, send.data_ptr())
        return send, counts

    def cc_resolve(self, recs, gid_base):
        edges = self._empty(recs.shape[0], torch.int64)
        ne = self.ctx.cc_resolve(recs.data_ptr(), recs.shape[0], gid_base, edges.data_ptr())
        return edges[:ne]

    def cc_merge(self, edges, n_super):
        return self.ctx.cc_merge(edges.data_ptr() if edges.numel() else None, edges.numel(), n_super)

    def cc_sizes(self, gid_base, n_final):
        sizes = torch.empty(max(n_final, 1), dtype=torch.int64, device=self.device)
        self.ctx.cc_sizes(gid_base, sizes.data_ptr())
        return sizes[:n_final]

    def cc_route(self, owner_of_final, gid_base, world, k1):
        own = owner_of_final.to(device=self.device, dtype=torch.int32).contiguous()
        counts = self.ctx.cc_route(own.data_ptr(), gid_base, world)
        send = self._empty((sum(counts), self.rec_words(k1)), torch.int64)
        self.ctx.cc_route(own.data_ptr(), gid_base, world, counts, send.data_ptr())
        return send, counts

    def cc_free(self):
        self.ctx.cc_free()

    def l3_walks(self, min_weight, min_length):
        self.ctx.l3_walks_phase(min_weight, min_length)

    def cand_export(self):
        n, nb = self.ctx.l3_cand_sizes()
        w = self._empty(n, torch.int32)
        line = self._empty(n, torch.int64)
        offs = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        codes = self._empty(nb, torch.uint8)
        self.ctx.l3_cand_export(w.data_ptr(), line.data_ptr(), offs.data_ptr(), codes.data_ptr())
        return w, line, offs, codes

    def l3_filter(self, codes, offs, n_cand):
        return self.ctx.l3_filter_phase(codes.data_ptr() if codes.numel() else None, offs.data_ptr(),
                                        n_cand, external=True, allow_missing=True)

    def allowed_weights(self, n_allowed):
        w = torch.zeros(max(n_allowed, 1), dtype=torch.int32, device=self.device)
        self.ctx.l3_allowed_copy(None, w.data_ptr())
        return w[:n_allowed]

    def set_allowed_weights(self, w):
        self.ctx.l3_set_allowed_weights(w.data_ptr() if w.numel() else None)

    def relieve(self, min_free=0.35):
        """Two caching allocators share the device (torch's and the library's): when less than
        `min_free` of the memory is free, both hand their cached blocks back to the driver."""
        _, free, total = self.ctx.device_info()
        if free < min_free * total:
            torch.cuda.current_stream(self.device).synchronize()
            torch.cuda.empty_cache()
            self.ctx.trim()

    def assignments(self, n_comps, n_assign, first_record):
        """(comp_offsets int64 [n_comps+1], global record indices int32 [n_assign]) on the device"""
        offs = torch.empty(n_comps + 1, dtype=torch.int64, device=self.device)
        idx = self._empty(n_assign, torch.int32)
        self.ctx.l4_assignments_dev(n_comps, first_record, offs.data_ptr(), idx.data_ptr())
        return offs, idx

    def to_host(self, t):
        """device tensor -> numpy through a cached page-locked buffer (valid until the next call)"""
        nbytes = t.numel() * t.element_size()
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < nbytes:
            buf = self._pinned = torch.empty(max(nbytes + nbytes // 2, 1), dtype=torch.uint8, pin_memory=True)
        view = buf[:nbytes].view(t.dtype)
        view.copy_(t.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return view.numpy()


# ---- host-side decisions ---------------------------------------------------------------------------
def assign_components(sizes, world, exact_top=1 << 14):
    """Owner rank of every K1-mer graph component, the same on every rank: the largest components
    one by one onto the least loaded rank (longest processing time first), the long tail of small
    ones in blocks."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = len(sizes)
    owner = np.zeros(n, dtype=np.int32)
    if world == 1 or n == 0:
        return owner
    order = np.argsort(-sizes, kind="stable")
    heap = [(0, r) for r in range(world)]
    top = order[:exact_top]
    for c, s in zip(top.tolist(), sizes[top].tolist()):
        load, r = heapq.heappop(heap)
        owner[c] = r
        heapq.heappush(heap, (load + s, r))
    rest = order[exact_top:]
    block = 1024
    for lo in range(0, len(rest), block):
        ids = rest[lo:lo + block]
        load, r = heapq.heappop(heap)
        owner[ids] = r
        heapq.heappush(heap, (load + int(sizes[ids].sum()), r))
    return owner


def check_component_fits(ops, sizes, owner, world, k1):
    """The walks need every K1-mer graph component whole on one rank.  Error K1-mers link unrelated
    transcripts whenever two of them share a (K1-1)-mer by chance; beyond ~2.5 x 10^9 distinct K1-mers
    (k1 = 25) those chance links percolate and most of the input becomes ONE component, which no
    single table can hold.  Refuse loudly instead of running out of memory in the table build."""
    from . import _lib
    load = np.bincount(owner, weights=sizes.astype(np.float64), minlength=world)
    worst = int(load.max())
    slot_bytes = 16 if k1 <= 32 else 32
    need = worst * 2 * slot_bytes + worst * 2 * 4 * 4       # table at load 0.5 + four 32-bit arrays per slot
    limit = None
    info = getattr(ops, "ctx", None)
    if info is not None:
        _, _, total = info.device_info()
        limit = 0.9 * total
    if worst * 2 >= 0xFFFFFFFF or (limit is not None and need > limit):
        raise _lib.ShnError(
            "sharded front end: rank load of %d K1-mers (largest K1-mer graph component: %d of %d K1-mers, "
            "%.1f %%) needs a %d GB table shard: the component-sharded walks cannot hold it (chance links "
            "between error K1-mers have percolated into a giant component)"
            % (worst, int(sizes.max()), int(sizes.sum()), 100.0 * sizes.max() / max(sizes.sum(), 1), need >> 30))


def merge_candidates(comm, w, gline, offs, codes):
    """All ranks' candidates (rank-local pop order) -> the global pop order of the reference's seed
    loop: seed weight descending, then LATER input line first (stable ascending sort by weight
    consumed from the end, extension_correction.py:334,343-344).  Returns (codes, offsets, n)."""
    lens = offs[1:] - offs[:-1]
    meta, counts = comm.all_gather_rows(torch.stack([w.to(torch.int64), gline, lens], dim=1))
    all_codes, _ = comm.all_gather_rows(codes)
    n = meta.shape[0]
    dev = meta.device
    if n == 0:
        return all_codes[:0], torch.zeros(1, dtype=torch.int64, device=dev), 0
    gw, gl, glen = meta[:, 0], meta[:, 1], meta[:, 2]
    src_off = torch.cumsum(glen, 0) - glen                 # gathered order = concatenation of the ranks
    order = torch.argsort(gl, descending=True, stable=True)
    order = order[torch.argsort(gw[order], descending=True, stable=True)]
    mlen = glen[order]
    moffs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    moffs[1:] = torch.cumsum(mlen, 0)
    total = int(moffs[-1].item())
    if total == 0:
        return all_codes[:0], moffs, n
    start = torch.repeat_interleave(src_off[order] - moffs[:-1], mlen)
    merged = all_codes[start + torch.arange(total, dtype=torch.int64, device=dev)]
    return merged, moffs, n


class _Trace(object):
    """SHN_SHARD_TRACE=1: wall-clock time of every sub-step (with a device synchronisation in front
    of each mark, so the times are attributable) on stderr."""

    def __init__(self, ops, rank):
        import os
        self.on = bool(os.environ.get("SHN_SHARD_TRACE"))
        self.ops, self.rank = ops, rank
        self.t = time.perf_counter()

    def __call__(self, what):
        if not self.on:
            return
        import sys
        if getattr(self.ops, "device", None) is not None and self.ops.device.type == "cuda":
            torch.cuda.synchronize(self.ops.device)
        now = time.perf_counter()
        sys.stderr.write("[shard %d] %-28s %8.2f ms\n" % (self.rank, what, 1000.0 * (now - self.t)))
        self.t = now


# ---- the sharded L3 stage ----------------------------------------------------------------------------
def correct_sharded(comm, ops, d_keys, d_counts, n_lines, first_line, k1, double_stranded, min_weight,
                    min_length, timings=None, stats=None):
    """load_kmers .. contig components (extension_correction.py:317-450) over the ranks of `comm`.
    d_keys / d_counts: device pointers of this rank's slice of the input lines (line i of the slice
    is global line first_line + i).  On return every rank's context holds the SAME L3 result as a
    single-GPU shn_l3_run over the whole input (contigs, allowed set with weights, contig graph,
    labels).  Returns n_loaded (distinct K1-mers over all ranks)."""
    tm = timings if timings is not None else {}
    st = stats if stats is not None else {}
    world, rank = comm.world, comm.rank
    trace = _Trace(ops, rank)
    t0 = time.perf_counter()
    # 1. lines -> minimizer owners
    send, counts = ops.route_lines(d_keys, d_counts, n_lines, first_line, double_stranded, k1, world)
    trace("route_lines")
    recs, _ = comm.all_to_all_rows(send, counts)
    del send
    trace("all_to_all lines")
    ops.relieve()
    ops.build_from_records(recs, k1)
    del recs
    n_shard = ops.n_distinct()
    trace("shard table build")
    tm["shard_build"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # 2. global components of the successor graph
    n_local = ops.cc_local()
    trace("cc_local")
    locs = [v[0] for v in comm.exchange_ints([n_local])]
    gid_base = sum(locs[:rank])
    n_super = sum(locs)
    trace("exchange n_local")
    q, qcounts = ops.cc_cross(world, rank, gid_base, k1)
    trace("cc_cross")
    rq, _ = comm.all_to_all_rows(q, qcounts)
    st["cross_queries"] = int(q.shape[0])
    del q
    trace("all_to_all queries")
    edges = ops.cc_resolve(rq, gid_base)
    del rq
    trace("cc_resolve")
    all_edges, ecounts = comm.all_gather_rows(edges)
    st["cross_edges"] = int(sum(ecounts))
    del edges
    trace("all_gather edges")
    n_final = ops.cc_merge(all_edges, n_super)
    del all_edges
    trace("cc_merge")
    ops.relieve()
    sizes = comm.all_reduce_sum(ops.cc_sizes(gid_base, n_final))
    trace("cc_sizes + all_reduce")
    h_sizes = sizes.cpu().numpy()
    owner = assign_components(h_sizes, world)
    if len(h_sizes):
        st["largest_component"] = int(h_sizes.max())
        st["n_keys_global"] = int(h_sizes.sum())
        check_component_fits(ops, h_sizes, owner, world, k1)
    trace("assign_components")
    tm["components"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # 3. whole components to their ranks
    send, counts = ops.cc_route(torch.from_numpy(owner), gid_base, world, k1)
    ops.cc_free()
    trace("cc_route")
    recs, _ = comm.all_to_all_rows(send, counts)
    del send
    trace("all_to_all components")
    ops.relieve()
    ops.build_from_records(recs, k1)
    del recs
    ops.relieve()
    trace("component table build")
    tm["reshard"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    st.update(n_shard_keys=int(n_shard), n_local_comps=int(n_local), n_super=int(n_super),
              n_raw_comps_global=int(n_final), n_owned_keys=int(ops.n_distinct()))
    # 4. per-rank walks
    ops.l3_walks(min_weight, min_length)
    trace("l3_walks")
    tm["walks"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # 5. merge candidates, replicated filter stage
    w, line, offs, codes = ops.cand_export()
    st["n_local_candidates"] = int(w.shape[0])
    mcodes, moffs, n_cand = merge_candidates(comm, w, line, offs, codes)
    trace("merge candidates")
    sz = ops.l3_filter(mcodes, moffs, n_cand)
    trace("l3_filter")
    aw = ops.allowed_weights(sz["n_allowed"])
    comm.all_reduce_sum(aw)
    ops.set_allowed_weights(aw)
    tm["filter"] = time.perf_counter() - t0
    n_loaded = comm.exchange_ints([n_shard])
    return sum(v[0] for v in n_loaded)


def merge_partitions_device(offs_list, idx_list, n_comps, device):
    """[(comp_offsets, record_idx)] of the ranks in rank order -> one partition; inside a component
    the records of rank 0 come first (ranks own ascending record ranges: the reference's append
    order, kmers_for_component.py:345-351).  Tensors on `device`."""
    sizes = torch.stack([o[1:] - o[:-1] for o in offs_list])             # [world, n_comps]
    total = sizes.sum(dim=0)
    goff = torch.zeros(n_comps + 1, dtype=torch.int64, device=device)
    goff[1:] = torch.cumsum(total, 0)
    before = torch.cumsum(sizes, 0) - sizes
    out = torch.empty(int(goff[-1].item()), dtype=idx_list[0].dtype, device=device)
    comp_ids = torch.arange(n_comps, dtype=torch.int64, device=device)
    for r, (offs, idx) in enumerate(zip(offs_list, idx_list)):
        n = idx.shape[0]
        if n == 0:
            continue
        comp = torch.repeat_interleave(comp_ids, sizes[r])
        dst = goff[comp] + before[r][comp] + (torch.arange(n, dtype=torch.int64, device=device) - offs[comp])
        out[dst] = idx
    return goff, out


def frontend_sharded(comm, ops, ctx, d_keys, d_counts, n_lines, first_line, k1, mates, rec_lo, paired,
                     min_weight=3, min_length=75, partition_size=500, double_stranded=False):
    """The whole hot path (what shannon.py:459+467 compute) over the ranks of `comm`; the sharded
    twin of pipeline.frontend_in_memory.  mates: this rank's record range [rec_lo, rec_lo + n) of
    the read files as [(bases, offsets, n, on_device)].  Returns (cor, comp_offsets, record_idx,
    stats); the partition (global record indices, numpy) is returned on rank 0, None elsewhere."""
    tm, st = {}, {}
    pipeline.upload_reads_early(ctx, mates)      # host-resident reads: the H2D copy overlaps the L3 stage
    n_loaded = correct_sharded(comm, ops, d_keys, d_counts, n_lines, first_line, k1, double_stranded,
                               min_weight, min_length, tm, st)
    t0 = time.perf_counter()
    pipeline.load_reads(ctx, mates, staged=True)
    cor = pipeline.collect_correction(ctx, k1, n_loaded, tm, fetch_allowed=False)
    comp_of_contig, n_comps, pk = pipeline.component_ids(cor, partition_size)
    ctx.l4_map_add_l3_contigs(comp_of_contig[1:], True)
    ctx.l4_map_set_weights(None, None)
    n_assign, n_lookups, n_valid = ctx.l4_assign(paired, k1)
    dev = ops.device
    t_offs, t_idx = ops.assignments(n_comps, n_assign, rec_lo)
    all_offs, _ = comm.all_gather_rows(t_offs.view(1, -1))
    all_idx, icounts = comm.all_gather_rows(t_idx)
    tot = comm.exchange_ints([n_assign, n_lookups, n_valid])
    comp_offs = rec_idx = None
    if comm.rank == 0:
        bounds = np.concatenate([[0], np.cumsum(icounts)])
        goff, merged = merge_partitions_device(
            [all_offs[r] for r in range(comm.world)],
            [all_idx[int(bounds[r]):int(bounds[r + 1])] for r in range(comm.world)], n_comps, dev)
        comp_offs = goff.cpu().numpy()
        rec_idx = ops.to_host(merged).view(np.uint32)
    tm["partition_reads"] = time.perf_counter() - t0
    stats = {"assignments": sum(v[0] for v in tot), "lookups": sum(v[1] for v in tot),
             "valid_records": sum(v[2] for v in tot)}
    stats["host_timings_ms"] = dict((k, 1000.0 * v) for k, v in tm.items())
    stats.update(cor.sizes)
    stats.update(st)
    stats["n_loaded"] = n_loaded
    stats["n_partitions"] = n_comps
    stats["n_singles"] = len(pk.singles)
    return cor, comp_offs, rec_idx, stats
