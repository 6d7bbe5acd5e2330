"""Drop-in replacement of the reference's ``kmers_for_component.py`` (same signature, return value
and output files; kmers_for_component.py:144-558) with the K1-mer -> component map, the read
partition and the per-component K1-mer weights computed on the B200 (libshannon_b200.so).
"""
import math
import os
import time

import numpy as np

from . import _lib
from .extension_correction import AllowedKmerDict, get_context
from .pipeline import build_component_map, contig_arrays, partition_reads
from .weight_updated_graph import weight_updated_graph


LAST_TIMINGS = {}     # wall-clock seconds of the sections of the last kmers_for_component call


def run_cmd(s1):
    print(s1)
    os.system(s1)


def _count_files(pattern):
    n = 0
    while os.path.exists(pattern % (n + 1)):
        n += 1
    return n


def _dict_arrays(ctx, k1mer_dictionary, k1):
    """(packed keys, weights) of the caller's k1mer_dictionary (any Mapping[str,int])."""
    if isinstance(k1mer_dictionary, AllowedKmerDict):
        if k1mer_dictionary.k1 == k1 or len(k1mer_dictionary) == 0:
            return k1mer_dictionary.keys_packed, k1mer_dictionary.weights
        return np.empty(0, np.uint64), np.empty(0, np.uint32)
    ks, ws = [], []
    for k, w in k1mer_dictionary.items():
        if len(k) == k1 and not k.strip("ACGT"):    # other keys can never equal a contig window
            ks.append(k)
            ws.append(int(w))
    if not ks:
        return np.empty(0, np.uint64), np.empty(0, np.uint32)
    keys = ctx.pack_kmers("".join(ks).encode(), len(ks), k1)
    return keys, np.asarray(ws, dtype=np.uint32)


NR = 10000000     # reads per chunk of the reference's partition loop (kmers_for_component.py:322)


def _gather_reads(bases, offs, idx):
    """(bases, offsets) of the reads `idx`, concatenated in that order."""
    offs = np.asarray(offs, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    lens = offs[idx + 1] - offs[idx]
    out_offs = np.zeros(len(idx) + 1, dtype=np.uint64)
    out_offs[1:] = np.cumsum(lens)
    total = int(out_offs[-1])
    if total == 0:
        return np.empty(0, dtype=np.uint8), out_offs
    start = np.repeat(offs[idx] - out_offs[:-1].astype(np.int64), lens)
    return bases[start + np.arange(total, dtype=np.int64)], out_offs


def _valid_reads(bases, offs):
    """not read.strip('ACTG') per read (:336,376)"""
    offs = np.asarray(offs, dtype=np.int64)
    bad = np.concatenate([[0], np.cumsum(~np.isin(bases, np.frombuffer(b"ACGT", dtype=np.uint8)))])
    return bad[offs[1:]] == bad[offs[:-1]]


def _double_strand(ctx, read_bases, read_offs, paired_end):
    """The read list the reference's loop sees with double_stranded=True and nJobs = 1: per chunk of
    NR valid records, SE: the reads then their reverse complements (rc, :36-42); PE: the pairs
    (r1, rc(r2)) then the pairs (r2, rc(r1)) (rc_mate_ds, :44-52).  Reverse complements on the GPU."""
    valid = _valid_reads(read_bases[0], read_offs[0])
    if paired_end:
        valid &= _valid_reads(read_bases[1], read_offs[1])
    v = np.nonzero(valid)[0]
    out_b = [[] for _ in read_bases]
    out_n = [[] for _ in read_bases]
    for lo in range(0, max(len(v), 1), NR):
        sel = v[lo:lo + NR]
        m = [_gather_reads(read_bases[k], read_offs[k], sel) for k in range(len(read_bases))]
        rc = [(ctx.revcomp_var(b, o), o) for b, o in m]
        if paired_end:
            parts = [[m[0], m[1]], [rc[1], rc[0]]]
        else:
            parts = [[m[0], rc[0]]]
        for k, seq in enumerate(parts):
            for b, o in seq:
                out_b[k].append(b)
                out_n[k].append(np.diff(o.astype(np.int64)))
    bases, offs = [], []
    for k in range(len(read_bases)):
        bases.append(np.concatenate(out_b[k]) if out_b[k] else np.empty(0, np.uint8))
        lens = np.concatenate(out_n[k]) if out_n[k] else np.empty(0, np.int64)
        o = np.zeros(len(lens) + 1, dtype=np.uint64)
        o[1:] = np.cumsum(lens)
        offs.append(o)
    return bases, offs


def kmers_for_component(k1mer_dictionary, kmer_directory, reads, reads_files, directory_name,
                        contig_file_extension, get_partition_k1mers, double_stranded=True,
                        paired_end=False, repartition=False, partition_size=500, overload=1.5,
                        K=24, gpmetis_path='gpmetis', penalty=5, only_reads=False, inMem=False,
                        nJobs=1, ctx=None):
    """See the reference docstring (kmers_for_component.py:145-155).  ``kmer_directory`` and
    ``reads`` are unused there as well.  ``double_stranded=True`` (never passed by shannon.py:427,467)
    follows every chunk of NR valid reads by its reverse complements like the reference with
    nJobs = 1 (:36-52,117-141,341,381); with nJobs > 1 the reference concatenates its workers'
    results in queue-arrival order, here the result does not depend on nJobs."""
    if not get_partition_k1mers:
        return None
    ctx = ctx or get_context()
    k1 = K + 1
    LAST_TIMINGS.clear()
    t_sec = time.perf_counter()

    def section(name):
        nonlocal t_sec
        now = time.perf_counter()
        LAST_TIMINGS[name] = LAST_TIMINGS.get(name, 0.0) + now - t_sec
        t_sec = now
    log_path = directory_name + "/before_sp_log.txt"
    f_log = open(log_path, 'a' if os.path.exists(log_path) else 'w')

    def write_log(s):
        f_log.write(s + "\n")
        print(s)

    n_components = _count_files(directory_name + "/component%dcontigs.txt")
    n_remaining = _count_files(directory_name + "/remaining_contigs%d.txt")

    # ---- gpmetis on the oversized components (external binary, :207-237) ----------------------
    ufactor = int(1000.0 * overload - 1000.0)
    components_broken = {}
    temp_string = ""
    for i in range(n_components):
        base = directory_name + "/component" + str(i + 1)
        with open(base + contig_file_extension, 'r') as f:
            num_contigs = len(f.readlines())
        partitions = min(int(math.ceil(float(num_contigs) / float(partition_size))), 100)
        components_broken[i] = partitions
        temp_string += "Component " + str(i) + ": " + str(partitions) + " partitions, "
        if num_contigs >= 2:
            run_cmd(gpmetis_path + " -ufactor=" + str(ufactor) + " " + base + ".txt " + str(partitions))
            if repartition:
                write_log(str(time.asctime()) + ": " + "Creating graph for repartition ")
                weight_updated_graph(directory_name,
                                     "/component" + str(i + 1) + ".txt.part." + str(partitions),
                                     "/component" + str(i + 1) + ".txt",
                                     "/component" + str(i + 1) + "r2.txt",
                                     "/component" + str(i + 1) + contig_file_extension,
                                     "/component" + str(i + 1) + contig_file_extension,
                                     penalty, False)
                write_log(str(time.asctime()) + ": " + "Created graph for repartition ")
                run_cmd(gpmetis_path + " -ufactor=" + str(ufactor) + " " + base + "r2.txt " + str(partitions))
    write_log(str(time.asctime()) + ": " + "gpmetis for partitioning is complete \n " + temp_string)

    section("kfc_gpmetis")
    # ---- component membership of every contig (:244-305), host side: contig-level text -------
    new_components = {}          # name -> [contig strings], insertion ordered
    comp_index = {}              # name -> dense id
    entries = []                 # (contig string, comp id) for the device map
    comp_contig_ids = {}         # name -> indices into `entries`

    def add(comp, contig):
        if comp not in new_components:
            new_components[comp] = []
            comp_index[comp] = len(comp_index)
            comp_contig_ids[comp] = []
        new_components[comp].append(contig)
        comp_contig_ids[comp].append(len(entries))
        entries.append((contig, comp_index[comp]))

    for i in components_broken:
        base = directory_name + "/component" + str(i + 1)
        with open(base + contig_file_extension, 'r') as f:
            contig_lines = f.readlines()
        passes = [('c', ".txt.part.")]
        if repartition:
            passes.append(('r2_c', "r2.txt.part."))
        for prefix, ext in passes:
            with open(base + ext + str(components_broken[i]), 'r') as f_component:
                for j, line in enumerate(f_component):
                    add(prefix + str(i + 1) + "_" + line.split()[0], contig_lines[j].split()[0])
    for i in range(n_remaining):
        with open(directory_name + "/remaining_contigs" + str(i + 1) + ".txt", 'r') as f:
            for line in f.readlines():
                add("cremaining" + str(i + 1), line.split()[0])

    # ---- k1mers2component on the device (a10) -------------------------------------------------
    ctg_bases, ctg_offs = contig_arrays([c for c, _ in entries])
    ctg_comp = np.asarray([cid for _, cid in entries], dtype=np.uint32)
    d_keys, d_w = _dict_arrays(ctx, k1mer_dictionary, k1)
    build_component_map(ctx, ctg_bases, ctg_offs, ctg_comp, k1, d_keys, d_w)
    write_log(str(time.asctime()) + ": " + "k1mers2component dictionary created ")
    section("kfc_component_map")

    # ---- read partition (a11) ---------------------------------------------------------------
    n_files = 2 if paired_end else 1
    suffix = ["_1", "_2"] if paired_end else [""]
    rb0, ro0 = ctx.load_fasta(reads_files[0])
    n_records = len(ro0) - 1
    read_bases, read_offs = [rb0], [ro0]
    if paired_end:
        rb1, ro1 = ctx.load_fasta(reads_files[1], n_records)
        read_bases.append(rb1)
        read_offs.append(ro1)
    if double_stranded:
        read_bases, read_offs = _double_strand(ctx, read_bases, read_offs, paired_end)
    section("kfc_read_fasta")
    n_comps = len(new_components)
    comp_offs, rec_idx, _ = partition_reads(
        ctx, [(b, o, None, False) for b, o in zip(read_bases, read_offs)], paired_end, k1, n_comps)

    section("kfc_partition_gpu")
    part = [dict() for _ in range(n_files)]
    read_text = [b.tobytes().decode() for b in read_bases] if inMem else None
    read_offs_l = [o.tolist() for o in read_offs] if inMem else None
    for comp, cid in comp_index.items():
        sel = rec_idx[comp_offs[cid]:comp_offs[cid + 1]]
        for m in range(n_files):
            if inMem:
                txt, o = read_text[m], read_offs_l[m]
                part[m][comp] = [txt[o[r]:o[r + 1]] for r in sel.tolist()]
            else:
                ctx.write_fasta_subset(directory_name + "/reads" + str(comp) + suffix[m] + ".fasta",
                                       True, read_bases[m], read_offs[m], sel, 0, suffix[m])
    write_log(str(time.asctime()) + ": " + "reads partititoned ")
    section("kfc_write_reads")

    # ---- per-component K1-mer files (a12) -----------------------------------------------------
    contig_weights = {}
    if not only_reads:
        write_log(str(time.asctime()) + ": Writing k1mers to file")
        win_w, win_off = ctx.l4_map_window_weights(ctg_bases, ctg_offs, k1)
        for comp in new_components:
            path = directory_name + "/component" + comp + "k1mers_allowed.dict"
            ids = np.asarray(comp_contig_ids[comp], dtype=np.uint32)
            contig_weights[comp] = []
            if inMem:
                open(path, 'w').close()
                for e in ids.tolist():
                    contig_weights[comp].append(
                        win_w[int(win_off[e]):int(win_off[e + 1])].astype(np.int64).tolist())
            else:
                ctx.write_k1mer_windows(path, ctg_bases, ctg_offs, ids, k1, win_w, win_off)
        write_log(str(time.asctime()) + ": " + "k1mers written to file ")
    section("kfc_write_k1mers")
    write_log(str(time.asctime()) + ": " + "kmers written to file " + "\n")
    f_log.close()

    if inMem:
        new_comps = new_components
        if paired_end:
            rps = dict((c, [[part[0][c]], [part[1][c]]]) for c in new_components)
        else:
            rps = dict((c, [part[0][c]]) for c in new_components)
    else:
        new_comps = [c for c in new_components]
        contig_weights = []
        rps = {}
    return [components_broken, new_comps, contig_weights, rps]
