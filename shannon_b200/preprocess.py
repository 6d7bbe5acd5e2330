"""The two driver steps in front of the hot path, on the B200:

* ``rc_double``       -- reverse-complement doubling of the read files (shannon.py:395-424, rc_gnu.py,
                         rc_s.py): same output files, byte for byte, for the FASTA the driver handles
                         (one sequence line per header);
* ``jellyfish_count`` -- ``jellyfish count -m K+1`` + ``jellyfish dump -c -t -L cutoff`` (shannon.py:
                         439-441): writes ``k1mer.dict_org``; the jellyfish binary is absent here and
                         its line order is unpinned, so the documented order of the stand-in counter
                         (ascending ASCII, oracle/kmer_count.py) is produced;
* ``frontend_from_fasta`` -- the same two steps feeding the hot path WITHOUT the text round trips:
                         reads go to the device once, are doubled there, counted there, and the
                         counted K1-mers go straight into ``shn_table_build`` and the reads into the
                         read partition.

Reads have any length and may contain N (K1-mer windows with a non-ACGT character are not counted;
reads with one are not partitioned, kmers_for_component.py:336).
"""
import shutil

import numpy as np

from . import pipeline

COUNT_CHUNK_BASES = 1 << 30     # ASCII bases per counting chunk (1 GB on the device at a time)


def _rc_file(ctx, src, dst):
    """rc_s.py: headers kept (stripped), every sequence line reverse-complemented."""
    names, noffs, bases, offs = ctx.load_fasta_named(src)
    rc = ctx.revcomp_var(bases, offs)
    ctx.write_fasta_named(dst, False, names, noffs, rc, offs)
    n = len(offs) - 1
    return n, (float(offs[-1]) / n if n else 0.0)


def _cat(parts, dst):
    with open(dst, "wb") as out:
        for p in parts:
            with open(p, "rb") as f:
                shutil.copyfileobj(f, out, 1 << 24)


def rc_double(ctx, reads_files, kmer_directory, paired_end, double_stranded):
    """shannon.py:395-424.  Returns (reads_files for the rest of the pipeline, N, L)."""
    import os
    d = kmer_directory
    if not paired_end:
        if not double_stranded:
            names, noffs, bases, offs = ctx.load_fasta_named(reads_files[0])   # rc_gnu.find_L
            n = len(offs) - 1
            return list(reads_files), n, (float(offs[-1]) / n if n else 0.0)
        rc = d + "/rc.fasta"
        n, l = _rc_file(ctx, reads_files[0], rc)
        _cat([reads_files[0], rc], d + "/reads.fasta")
        os.remove(rc)
        return [d + "/reads.fasta"], n, l
    if not double_stranded:
        n, l = _rc_file(ctx, reads_files[1], d + "/rc_2.fasta")
        return [reads_files[0], d + "/rc_2.fasta"], n, l
    rc1, rc2 = d + "/rc_1.fasta", d + "/rc_2.fasta"
    _rc_file(ctx, reads_files[0], rc1)
    n, l = _rc_file(ctx, reads_files[1], rc2)
    _cat([reads_files[0], rc2], d + "/reads_1.fasta")
    _cat([rc1, reads_files[1]], d + "/reads_2.fasta")
    os.remove(rc1)
    os.remove(rc2)
    return [d + "/reads_1.fasta", d + "/reads_2.fasta"], n, l


def _count_arrays(ctx, arrays, k1, expected=None):
    """Counts the K1-mers of [(bases uint8, offsets uint64)] host arrays in chunks; the counting
    table is re-sized and the pass repeated if the first estimate was too small."""
    from . import _lib
    total_windows = sum(int(np.maximum(np.diff(o.astype(np.int64)) - k1 + 1, 0).sum()) for _, o in arrays)
    expected = expected or max(1 << 16, total_windows // 4)
    while True:
        try:
            ctx.count_begin(k1, min(expected, max(total_windows, 1)))
            for bases, offs in arrays:
                offs = np.asarray(offs, dtype=np.uint64)
                lo = 0
                while lo < len(offs) - 1:
                    hi = int(np.searchsorted(offs, offs[lo] + np.uint64(COUNT_CHUNK_BASES), side="right")) - 1
                    hi = min(max(hi, lo + 1), len(offs) - 1)
                    ctx.count_add_reads(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo])
                    lo = hi
            return
        except _lib.ShnError as e:
            if "table full" not in str(e) or expected >= total_windows:
                raise
            expected = min(2 * expected, total_windows)


def jellyfish_count(ctx, reads_files, K, out_path=None, cutoff=1):
    """shannon.py:439-441 for FASTA read files.  Returns (d_keys, d_counts, n): device arrays owned by
    the context (valid until the next count); writes ``k1mer.dict_org`` when out_path is given."""
    k1 = K + 1
    arrays = []
    for f in reads_files:
        _, _, bases, offs = ctx.load_fasta_named(f)
        arrays.append((bases, offs))
    _count_arrays(ctx, arrays, k1)
    d_keys, d_counts, n = ctx.count_finish(cutoff)
    if out_path is not None:
        kw = 2 if k1 > 32 else 1
        keys = ctx.d2h(np.empty(n * kw, dtype=np.uint64), d_keys) if n else np.empty(0, np.uint64)
        counts = ctx.d2h(np.empty(n, dtype=np.uint32), d_counts) if n else np.empty(0, np.uint32)
        ctx.write_kmer_file(out_path, keys, counts, k1)
    return d_keys, d_counts, n


def doubled_reads(ctx, reads_files, paired_end, double_stranded):
    """The read files the rest of the pipeline sees (shannon.py:395-424) as host arrays, with the
    reverse complements made on the device: [(bases, offsets)] per mate file."""
    raw = []
    for f in reads_files:
        _, _, bases, offs = ctx.load_fasta_named(f)
        raw.append((bases, offs.astype(np.uint64)))

    def cat(a, b):
        return (np.concatenate([a[0], b[0]]),
                np.concatenate([a[1], a[1][-1] + b[1][1:]]).astype(np.uint64))

    def rc(a):
        return (ctx.revcomp_var(a[0], a[1]), a[1])
    if not paired_end:
        return [cat(raw[0], rc(raw[0]))] if double_stranded else [raw[0]]
    if not double_stranded:
        return [raw[0], rc(raw[1])]
    return [cat(raw[0], rc(raw[1])), cat(rc(raw[0]), raw[1])]


def frontend_from_fasta(ctx, reads_files, K, paired_end, double_stranded=True, min_weight=3,
                        min_length=75, partition_size=500, cutoff=1):
    """FASTA read files -> contigs, components and the read partition with no intermediate text:
    RC doubling and K1-mer counting on the device, the counted K1-mers handed to the table build as
    device arrays (what shannon.py:395-467 computes through reads*.fasta and k1mer.dict_org).
    Returns (cor, comp_offsets, record_idx, stats, mates) like pipeline.frontend_in_memory."""
    k1 = K + 1
    mates = doubled_reads(ctx, reads_files, paired_end, double_stranded)
    _count_arrays(ctx, mates, k1)
    d_keys, d_counts, n = ctx.count_finish(cutoff)
    cor, comp_offs, rec_idx, stats = pipeline.frontend_in_memory(
        ctx, d_keys, d_counts, k1, [(b, o, None, False) for b, o in mates], paired_end, min_weight,
        min_length, partition_size, on_device=True, n_kmers=n)
    return cor, comp_offs, rec_idx, stats, mates
