"""The front end as in-memory building blocks on one GPU context.  The two drop-in modules
(extension_correction.py, kmers_for_component.py) are these blocks plus the reference's file
formats; bench.py times the same blocks without touching the disk.
"""
import numpy as np

from . import _lib

_CODE_TO_ASCII = np.frombuffer(b"AGCT", dtype=np.uint8)


def decode_kmers(keys, k1):
    """uint64 packed keys -> (n, k1) uint8 ASCII matrix."""
    keys = np.asarray(keys, dtype=np.uint64)
    shifts = (2 * (k1 - 1 - np.arange(k1))).astype(np.uint64)
    codes = ((keys[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    return _CODE_TO_ASCII[codes]


def contig_adjacency(n_contigs, a, b, w, fp):
    """contig_connections (extension_correction.py:372-389) in the reference's dict insertion
    order, from the GPU's distinct edge list (a < b, multiplicity w, fp = first C-mer position in
    b shared with a): node x first gets its earlier neighbours ordered by (fp, id) -- they are
    connected while x itself is being indexed -- then later contigs in ascending id."""
    adj = [[] for _ in range(n_contigs + 1)]
    if len(a):
        a = a.astype(np.int64)
        b = b.astype(np.int64)
        wl = w.tolist()
        al, bl = a.tolist(), b.tolist()
        lower = np.lexsort((a, fp.astype(np.int64), b))      # by b, then fp, then a
        for e in lower.tolist():
            adj[bl[e]].append((al[e], wl[e]))
        higher = np.lexsort((b, a))                          # by a, then b
        for e in higher.tolist():
            adj[al[e]].append((bl[e], wl[e]))
    return adj


def dfs_components(n_contigs, adj):
    """extension_correction.py:417-434: iterative DFS in ascending contig index; member order is
    the pop order.  Returns ({root: [members]} in insertion order, comp_of[contig])."""
    comp_of = [0] * (n_contigs + 1)
    seen = [False] * (n_contigs + 1)
    component2contig = {}
    for root in range(1, n_contigs + 1):
        if comp_of[root]:
            continue
        if not adj[root]:                       # isolated contig: a singleton component
            comp_of[root] = root
            seen[root] = True
            component2contig[root] = [root]
            continue
        members = component2contig[root] = []
        stack = [root]
        seen[root] = True
        while stack:
            cur = stack.pop()
            comp_of[cur] = root
            members.append(cur)
            for nb, _ in adj[cur]:
                if not seen[nb]:
                    stack.append(nb)
                    seen[nb] = True
    return component2contig, comp_of


class Correction(object):
    """Result of the L3 stage (run_correction's in-memory state after :450)."""
    __slots__ = ("k1", "n_loaded", "sizes", "contigs", "allowed_keys", "allowed_weights", "adj",
                 "component2contig", "comp_of", "n_edges")


def correct(ctx, keys, counts, k1, double_stranded, min_weight, min_length, on_device=False,
            n=None, timings=None, fetch_allowed=True, after_table_build=None):
    """load_kmers .. DFS (extension_correction.py:317-450) on the GPU.  keys/counts: host numpy
    arrays, or device pointers with on_device=True and n given."""
    import time
    tm = timings if timings is not None else {}
    t0 = time.perf_counter()
    ctx.table_build(keys, counts, k1, double_stranded, on_device=on_device, n=n)
    tm["table_build"] = time.perf_counter() - t0
    if after_table_build is not None:
        after_table_build()      # e.g. start the read upload now that the copy engine is free
    cor = Correction()
    cor.k1 = k1
    cor.n_loaded = ctx.table_stats()["n_distinct"]
    if cor.n_loaded == 0:
        raise StopIteration("no K1-mers loaded")  # the reference fails on next(iter(kmers)) here
    t0 = time.perf_counter()
    cor.sizes = ctx.l3_run(min_weight, min_length)
    tm["l3_run"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_contigs = cor.sizes["n_contigs"]
    bases, offs = ctx.l3_contigs()
    text = bases.tobytes().decode()
    o = offs.tolist()
    cor.contigs = [None] + [text[o[i]:o[i + 1]] for i in range(n_contigs)]
    # allowed_kmer_dict (:404-408); left on the device when the caller feeds it straight into L4
    cor.allowed_keys, cor.allowed_weights = ctx.l3_allowed() if fetch_allowed else (None, None)
    ea, eb, ew, efp = ctx.l3_edges()
    cor.adj = contig_adjacency(n_contigs, ea, eb, ew, efp)
    cor.component2contig, cor.comp_of = dfs_components(n_contigs, cor.adj)
    labels = ctx.l3_labels()
    if n_contigs and not np.array_equal(labels[1:], np.asarray(cor.comp_of[1:], dtype=np.uint32)):
        raise _lib.ShnError("internal error: GPU component labels disagree with the DFS partition")
    cor.n_edges = dict((c, 0) for c in cor.component2contig)
    for x in ea.tolist():
        cor.n_edges[cor.comp_of[x]] += 1
    tm["l3_host_order"] = time.perf_counter() - t0
    return cor


class Packing(object):
    """How run_correction distributes components over its output files (:458-513)."""
    __slots__ = ("singles", "remaining", "big")


def pack_components(cor, comp_size_threshold):
    p = Packing()
    p.singles, p.remaining, p.big = [], [[]], []
    cur = 0
    for root, members in cor.component2contig.items():
        if len(members) == 1:
            p.singles.append(members[0])
        elif len(members) > comp_size_threshold:
            p.big.append((root, members))
        else:
            p.remaining[-1].extend(members)
            cur += len(members)
            if cur > comp_size_threshold:
                p.remaining.append([])
                cur = 0
    return p


def contig_arrays(contig_strings):
    """list of str -> (uint8 bases, uint64 offsets)."""
    text = "".join(contig_strings)
    bases = np.frombuffer(text.encode(), dtype=np.uint8)
    offs = np.zeros(len(contig_strings) + 1, dtype=np.uint64)
    if contig_strings:
        offs[1:] = np.cumsum([len(c) for c in contig_strings])
    return bases, offs


def build_component_map(ctx, ctg_bases, ctg_offs, ctg_comp, k1, dict_keys, dict_weights):
    """k1mers2component (kmers_for_component.py:239-305) on the device."""
    lens = np.diff(ctg_offs.astype(np.int64))
    total_windows = int(np.maximum(lens - k1 + 1, 0).sum())
    ctx.l4_map_add_contigs(ctg_bases, ctg_offs, ctg_comp, k1, True, total_windows)
    ctx.l4_map_set_weights(dict_keys, dict_weights)
    return total_windows


def upload_reads_early(ctx, mates):
    """Start the H2D copy of host-resident read files now (it overlaps the L3 stage)."""
    for m, (bases, offs, n, on_dev) in enumerate(mates):
        if not on_dev:
            ctx.l4_upload_reads_async(m, bases, offs)


def partition_reads(ctx, mates, paired, k1, n_comps, staged=False):
    """get_comps over all records (kmers_for_component.py:322-423).  mates: list of
    (bases, offsets, n, on_device).  Returns (comp_offsets, record_idx, stats)."""
    for m, (bases, offs, n, on_dev) in enumerate(mates):
        if staged and not on_dev:
            ctx.l4_load_reads_staged(m)
        else:
            ctx.l4_load_reads(m, bases, offs, n=n, on_device=on_dev)
    n_assign, n_lookups, n_valid = ctx.l4_assign(paired, k1)
    comp_offs, rec_idx = ctx.l4_assignments(n_comps, n_assign)
    return comp_offs.astype(np.int64), rec_idx, {"assignments": n_assign, "lookups": n_lookups,
                                                  "valid_records": n_valid}


def frontend_in_memory(ctx, keys, counts, k1, mates, paired, min_weight=3, min_length=75,
                       partition_size=500, on_device=False, n_kmers=None):
    """Whole hot path without files: what shannon.py:459+467 compute, for bench.py.  Oversized
    components (more than partition_size contigs) are split into contiguous blocks, the rule of
    the gpmetis stand-in."""
    import time
    tm = {}
    cor = correct(ctx, keys, counts, k1, False, min_weight, min_length, on_device, n_kmers, tm,
                  fetch_allowed=False, after_table_build=lambda: upload_reads_early(ctx, mates))
    t0 = time.perf_counter()
    pk = pack_components(cor, partition_size)
    entries, comp_ids = [], []
    n_comps = 0
    for _, members in pk.big:
        parts = min(-(-len(members) // partition_size), 100)
        block = -(-len(members) // parts)
        for i, c in enumerate(members):
            entries.append(cor.contigs[c])
            comp_ids.append(n_comps + min(i // block, parts - 1))
        n_comps += parts
    for group in pk.remaining:
        if not group and len(pk.remaining) > 1 and group is pk.remaining[-1]:
            continue
        for c in group:
            entries.append(cor.contigs[c])
            comp_ids.append(n_comps)
        n_comps += 1
    bases, offs = contig_arrays(entries)
    tm["pack_host"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    build_component_map(ctx, bases, offs, np.asarray(comp_ids, dtype=np.uint32), k1,
                        cor.allowed_keys, cor.allowed_weights)
    tm["map_build"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    comp_offs, rec_idx, stats = partition_reads(ctx, mates, paired, k1, n_comps, staged=True)
    tm["partition_reads"] = time.perf_counter() - t0
    stats["host_timings_ms"] = dict((k, 1000.0 * v) for k, v in tm.items())
    stats.update(cor.sizes)
    stats["n_loaded"] = cor.n_loaded
    stats["n_partitions"] = n_comps
    stats["n_singles"] = len(pk.singles)
    return cor, comp_offs, rec_idx, stats
