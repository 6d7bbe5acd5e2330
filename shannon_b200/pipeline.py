"""The front end as in-memory building blocks on one GPU context.  The two drop-in modules
(extension_correction.py, kmers_for_component.py) are these blocks plus the reference's file
formats; bench.py times the same blocks without touching the disk.
"""
import numpy as np

from . import _lib

_CODE_TO_ASCII = np.frombuffer(b"AGCT", dtype=np.uint8)


def decode_kmers(keys, k1):
    """packed keys ((n,) uint64, or (n, 2) uint64 low word first for k1 = 33) -> (n, k1) uint8
    ASCII matrix."""
    keys = np.asarray(keys, dtype=np.uint64)
    shifts = 2 * (k1 - 1 - np.arange(k1))
    if keys.ndim == 1:
        codes = (keys[:, None] >> shifts.astype(np.uint64)[None, :]) & np.uint64(3)
    else:
        word = (shifts >= 64).astype(np.int64)                  # pairs never straddle a word
        sh = (shifts - 64 * word).astype(np.uint64)
        codes = (keys[:, word] >> sh[None, :]) & np.uint64(3)
    return _CODE_TO_ASCII[codes.astype(np.uint8)]


def encode_kmer(s):
    """str -> packed integer key (arbitrary precision), A=0 G=1 C=2 T=3, first base most significant."""
    x = 0
    for ch in s:
        x = (x << 2) | "AGCT".index(ch)
    return x


def keys_array(values, k1):
    """list of integer keys -> array in the layout the library uses ((n,) or (n, 2) uint64)."""
    if k1 <= 32:
        return np.asarray(values, dtype=np.uint64).reshape(-1)
    lo = np.asarray([v & 0xFFFFFFFFFFFFFFFF for v in values], dtype=np.uint64)
    hi = np.asarray([v >> 64 for v in values], dtype=np.uint64)
    return np.stack([lo, hi], axis=1) if len(values) else np.empty((0, 2), dtype=np.uint64)


def keys_as_ints(keys):
    """array of packed keys -> list of Python ints."""
    keys = np.asarray(keys, dtype=np.uint64)
    if keys.ndim == 1:
        return keys.tolist()
    return [(int(h) << 64) | int(l) for l, h in keys.tolist()]


def contig_adjacency_csr(n_contigs, a, b, w, fp):
    """contig_connections (extension_correction.py:372-389) in the reference's dict insertion
    order, from the GPU's distinct edge list (a < b, multiplicity w, fp = first C-mer position in
    b shared with a), as CSR arrays (indptr, neighbour, weight) over contigs 0..n_contigs:
    node x first gets its earlier neighbours ordered by (fp, id) -- they are connected while x
    itself is being indexed -- then later contigs in ascending id."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    w = np.asarray(w, dtype=np.int64)
    fp = np.asarray(fp, dtype=np.int64)
    # one row per (node, neighbour): phase 0 = earlier neighbours keyed (fp, a), phase 1 = later (b, 0)
    node = np.concatenate([b, a])
    nbr = np.concatenate([a, b])
    wt = np.concatenate([w, w])
    phase = np.concatenate([np.zeros(len(a), np.int64), np.ones(len(a), np.int64)])
    k1 = np.concatenate([fp, b])
    k2 = np.concatenate([a, np.zeros(len(a), np.int64)])
    order = np.lexsort((k2, k1, phase, node))
    node, nbr, wt = node[order], nbr[order], wt[order]
    indptr = np.zeros(n_contigs + 2, dtype=np.int64)
    np.add.at(indptr, node + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, nbr, wt


def contig_adjacency(n_contigs, a, b, w, fp):
    """Same as lists of (neighbour, weight) per contig (index 0 unused)."""
    indptr, nbr, wt = contig_adjacency_csr(n_contigs, a, b, w, fp)
    ip, nl, wl = indptr.tolist(), nbr.tolist(), wt.tolist()
    return [list(zip(nl[ip[x]:ip[x + 1]], wl[ip[x]:ip[x + 1]])) for x in range(n_contigs + 1)]


def dfs_components_csr(n_contigs, indptr, nbr):
    """extension_correction.py:417-434: iterative DFS in ascending contig index; member order is
    the pop order.  Returns ({root: [members]} in insertion order, comp_of[contig])."""
    ip, nl = indptr.tolist(), nbr.tolist()
    comp_of = [0] * (n_contigs + 1)
    seen = [False] * (n_contigs + 1)
    component2contig = {}
    for root in range(1, n_contigs + 1):
        if comp_of[root]:
            continue
        if ip[root] == ip[root + 1]:            # isolated contig: a singleton component
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
            for nb in nl[ip[cur]:ip[cur + 1]]:
                if not seen[nb]:
                    stack.append(nb)
                    seen[nb] = True
    return component2contig, comp_of


def dfs_components(n_contigs, adj):
    """Same on adjacency lists (as returned by contig_adjacency)."""
    indptr = np.zeros(n_contigs + 2, dtype=np.int64)
    indptr[1:] = np.cumsum([len(x) for x in adj])
    nbr = np.asarray([nb for x in adj for nb, _ in x], dtype=np.int64)
    return dfs_components_csr(n_contigs, indptr, nbr)


class ContigStore(object):
    """Accepted contigs as one ASCII byte array + offsets (1-based access like the reference's
    `contigs` list, extension_correction.py:341); strings are only made when asked for."""

    def __init__(self, bases, offs):
        self.bases = bases
        self.offs = np.asarray(offs, dtype=np.int64)
        self._text = None

    def __len__(self):
        return len(self.offs)            # index 0 unused, like ["buffer"] + contigs

    def __getitem__(self, i):
        if self._text is None:
            self._text = self.bases.tobytes().decode()
        return self._text[self.offs[i - 1]:self.offs[i]]

    def strings(self):
        return [self[i] for i in range(1, len(self))]

    def gather(self, ids):
        """(bases, offsets) of the contigs `ids` (1-based), concatenated in that order."""
        ids = np.asarray(ids, dtype=np.int64)
        lens = self.offs[ids] - self.offs[ids - 1]
        out_offs = np.zeros(len(ids) + 1, dtype=np.uint64)
        out_offs[1:] = np.cumsum(lens)
        total = int(out_offs[-1])
        if total == 0:
            return np.empty(0, dtype=np.uint8), out_offs
        start = np.repeat(self.offs[ids - 1] - out_offs[:-1].astype(np.int64), lens)
        return self.bases[start + np.arange(total, dtype=np.int64)], out_offs


class Correction(object):
    """Result of the L3 stage (run_correction's in-memory state after :450)."""
    __slots__ = ("k1", "n_loaded", "sizes", "contigs", "allowed_keys", "allowed_weights", "adj_csr",
                 "component2contig", "comp_of", "n_edges")

    def neighbours(self, contig):
        """[(neighbour, weight)] of a contig in the reference's dict order."""
        indptr, nbr, wt = self.adj_csr
        lo, hi = int(indptr[contig]), int(indptr[contig + 1])
        return list(zip(nbr[lo:hi].tolist(), wt[lo:hi].tolist()))


def correct(ctx, keys, counts, k1, double_stranded, min_weight, min_length, on_device=False,
            n=None, timings=None, fetch_allowed=True, after_table_build=None, after_l3_run=None):
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
    if after_l3_run is not None:
        after_l3_run()           # e.g. queue the read packing: it runs under the host-side ordering
    return collect_correction(ctx, k1, cor.n_loaded, tm, fetch_allowed, cor)


def collect_correction(ctx, k1, n_loaded, timings=None, fetch_allowed=True, cor=None):
    """The contig-level results of the L3 stage this context holds (after shn_l3_run, or after the
    sharded path's shn_l3_filter), ordered like the reference's dicts: contigs, allowed set, contig
    adjacency in insertion order, DFS components (extension_correction.py:403-450)."""
    import time
    tm = timings if timings is not None else {}
    if cor is None:
        cor = Correction()
        cor.k1 = k1
        cor.n_loaded = n_loaded
        cor.sizes = ctx.l3_sizes()
    t0 = time.perf_counter()
    n_contigs = cor.sizes["n_contigs"]
    bases, offs = ctx.l3_contigs()
    cor.contigs = ContigStore(bases, offs)
    # allowed_kmer_dict (:404-408); left on the device when the caller feeds it straight into L4
    cor.allowed_keys, cor.allowed_weights = ctx.l3_allowed() if fetch_allowed else (None, None)
    ea, eb, ew, efp = ctx.l3_edges()
    cor.adj_csr = contig_adjacency_csr(n_contigs, ea, eb, ew, efp)
    cor.component2contig, cor.comp_of = dfs_components_csr(n_contigs, cor.adj_csr[0], cor.adj_csr[1])
    labels = ctx.l3_labels()
    if n_contigs and not np.array_equal(labels[1:], np.asarray(cor.comp_of[1:], dtype=np.uint32)):
        raise _lib.ShnError("internal error: GPU component labels disagree with the DFS partition")
    cor.n_edges = dict((c, 0) for c in cor.component2contig)
    if len(ea):
        roots, cnt = np.unique(np.asarray(cor.comp_of, dtype=np.int64)[ea.astype(np.int64)],
                               return_counts=True)
        cor.n_edges.update(zip(roots.tolist(), cnt.tolist()))
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


def load_reads(ctx, mates, staged=False):
    """2-bit packing of the read files on the device (asynchronous on the context's stream)."""
    for m, (bases, offs, n, on_dev) in enumerate(mates):
        if staged and not on_dev:
            ctx.l4_load_reads_staged(m)
        else:
            ctx.l4_load_reads(m, bases, offs, n=n, on_device=on_dev)


def partition_reads(ctx, mates, paired, k1, n_comps, staged=False, loaded=False, pinned=False):
    """get_comps over all records (kmers_for_component.py:322-423).  mates: list of
    (bases, offsets, n, on_device).  Returns (comp_offsets, record_idx, stats)."""
    if not loaded:
        load_reads(ctx, mates, staged)
    n_assign, n_lookups, n_valid = ctx.l4_assign(paired, k1)
    comp_offs, rec_idx = ctx.l4_assignments(n_comps, n_assign, pinned=pinned)
    return comp_offs.astype(np.int64), rec_idx, {"assignments": n_assign, "lookups": n_lookups,
                                                  "valid_records": n_valid}


def component_ids(cor, partition_size):
    """Partition id of every accepted contig for the in-memory pipelines (index 0 unused;
    0xFFFFFFFF = single-contig component, not partitioned): the parts of every oversized component
    (contiguous blocks of its DFS order, the rule of the gpmetis stand-in), then the
    remaining_contigs groups in file order.  Returns (comp_of_contig, n_comps, packing)."""
    pk = pack_components(cor, partition_size)
    n_contigs = cor.sizes["n_contigs"]
    comp_of_contig = np.full(n_contigs + 1, 0xFFFFFFFF, dtype=np.uint32)   # singles stay NONE
    n_comps = 0
    for _, members in pk.big:
        parts = min(-(-len(members) // partition_size), 100)
        block = -(-len(members) // parts)
        comp_of_contig[members] = n_comps + np.minimum(np.arange(len(members)) // block, parts - 1)
        n_comps += parts
    for group in pk.remaining:
        if not group and len(pk.remaining) > 1 and group is pk.remaining[-1]:
            continue
        comp_of_contig[group] = n_comps
        n_comps += 1
    return comp_of_contig, n_comps, pk


def frontend_in_memory(ctx, keys, counts, k1, mates, paired, min_weight=3, min_length=75,
                       partition_size=500, on_device=False, n_kmers=None):
    """Whole hot path without files: what shannon.py:459+467 compute, for bench.py.  Oversized
    components (more than partition_size contigs) are split into contiguous blocks, the rule of
    the gpmetis stand-in.  The returned record_idx is a view of a page-locked buffer that the
    next call re-uses: copy it to keep it."""
    import time
    tm = {}
    loaded = []   # set once the read packing has been queued (under the host-side ordering)
    cor = correct(ctx, keys, counts, k1, False, min_weight, min_length, on_device, n_kmers, tm,
                  fetch_allowed=False, after_table_build=lambda: upload_reads_early(ctx, mates),
                  after_l3_run=lambda: loaded.append(load_reads(ctx, mates, staged=True)))
    t0 = time.perf_counter()
    comp_of_contig, n_comps, pk = component_ids(cor, partition_size)
    tm["pack_host"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # the accepted contigs and the allowed set are still on the device: no round trip
    ctx.l4_map_add_l3_contigs(comp_of_contig[1:], True)
    ctx.l4_map_set_weights(None, None)
    tm["map_build"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    comp_offs, rec_idx, stats = partition_reads(ctx, mates, paired, k1, n_comps, staged=True, loaded=bool(loaded),
                                                pinned=True)
    tm["partition_reads"] = time.perf_counter() - t0
    stats["host_timings_ms"] = dict((k, 1000.0 * v) for k, v in tm.items())
    stats.update(cor.sizes)
    stats["n_loaded"] = cor.n_loaded
    stats["n_partitions"] = n_comps
    stats["n_singles"] = len(pk.singles)
    return cor, comp_offs, rec_idx, stats
