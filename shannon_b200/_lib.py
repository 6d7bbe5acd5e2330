"""ctypes binding of libshannon_b200.so (include/shannon_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
available, every entry point raises -- the product never silently computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshannon_b200.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p


class ShnError(RuntimeError):
    pass


class L3Sizes(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_seeds", "n_raw_comps", "n_walks", "n_traversed", "n_candidates", "n_contigs",
        "contig_bases", "n_allowed", "n_edges", "dup_rounds", "walk_rounds", "n_spec_comps",
        "spec_windows")]


# name -> (restype, argtypes); must list every symbol declared in include/shannon_b200.h
SIGNATURES = {
    "shn_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "shn_destroy": (None, [vp]),
    "shn_last_error": (C.c_char_p, [vp]),
    "shn_version": (C.c_char_p, []),
    "shn_device_info": (C.c_int, [vp, C.POINTER(C.c_int), u64p, u64p]),
    "shn_dev_alloc": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
    "shn_dev_free": (C.c_int, [vp, vp]),
    "shn_host_alloc_pinned": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
    "shn_host_free_pinned": (C.c_int, [vp, vp]),
    "shn_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "shn_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "shn_sync": (C.c_int, [vp]),
    "shn_memcpy_d2d": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "shn_use_stream": (C.c_int, [vp, vp]),
    "shn_l3_walks": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
    "shn_l3_cand_sizes": (C.c_int, [vp, u64p, u64p]),
    "shn_l3_cand_export": (C.c_int, [vp, vp, vp, vp, vp]),
    "shn_l3_filter": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int]),
    "shn_l3_allowed_copy": (C.c_int, [vp, vp, vp]),
    "shn_l3_set_allowed_weights": (C.c_int, [vp, vp]),
    "shn_route_lines": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint32,
                                  u64p, vp]),
    "shn_table_build_records": (C.c_int, [vp, vp, C.c_uint64, C.c_int]),
    "shn_trim": (C.c_int, [vp]),
    "shn_cc_local": (C.c_int, [vp, u64p]),
    "shn_cc_cross": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, u64p, vp]),
    "shn_cc_resolve": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, u64p]),
    "shn_cc_merge": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, u64p]),
    "shn_cc_sizes": (C.c_int, [vp, C.c_uint64, vp]),
    "shn_cc_route": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, u64p, vp]),
    "shn_cc_free": (C.c_int, [vp]),
    "shn_timer_start": (C.c_int, [vp]),
    "shn_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "shn_prof_enable": (C.c_int, [vp, C.c_int]),
    "shn_prof_get": (C.c_int, [vp, C.c_char_p, C.POINTER(C.c_double), u64p]),
    "shn_prof_dump": (C.c_int, [vp, C.c_char_p, C.c_uint64]),
    "shn_launch_count": (C.c_uint64, [vp]),
    "shn_flush_l2": (C.c_int, [vp]),
    "shn_parse_kmer_file": (C.c_int, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp), u64p,
                                      C.POINTER(C.c_int)]),
    "shn_host_free": (None, [vp]),
    "shn_load_fasta": (C.c_int, [vp, C.c_char_p, C.c_int64, C.POINTER(vp), C.POINTER(vp), u64p]),
    "shn_write_fasta_subset": (C.c_int, [vp, C.c_char_p, C.c_int, vp, vp, vp, C.c_uint64,
                                         C.c_uint64, C.c_char_p]),
    "shn_write_kmer_file": (C.c_int, [vp, C.c_char_p, vp, vp, C.c_uint64, C.c_int]),
    "shn_count_release": (C.c_int, [vp]),
    "shn_condense_run": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_int, u64p, u64p, u64p, u64p]),
    "shn_condense_get": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "shn_find_reps": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_int, vp]),
    "shn_revcomp_var": (C.c_int, [vp, vp, vp, C.c_uint64, vp, C.c_int]),
    "shn_count_begin": (C.c_int, [vp, C.c_int, C.c_uint64]),
    "shn_count_add_reads": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int]),
    "shn_count_finish": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), C.POINTER(vp), u64p]),
    "shn_load_fasta_named": (C.c_int, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(vp), u64p]),
    "shn_write_fasta_named": (C.c_int, [vp, C.c_char_p, C.c_int, vp, vp, vp, vp, C.c_uint64]),
    "shn_write_k1mer_windows": (C.c_int, [vp, C.c_char_p, vp, vp, vp, C.c_uint64, C.c_int, vp, vp]),
    "shn_pack_kmers": (C.c_int, [vp, vp, C.c_uint64, C.c_int, vp, C.c_int]),
    "shn_table_build": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "shn_table_build_indexed": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_int]),
    "shn_route_plan": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, vp, u64p]),
    "shn_permute": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp]),
    "shn_table_stats": (C.c_int, [vp, u64p, u64p, u64p, C.POINTER(C.c_int)]),
    "shn_table_lookup": (C.c_int, [vp, vp, C.c_uint64, vp, vp, C.c_int]),
    "shn_table_dump": (C.c_int, [vp, vp, vp, vp]),
    "shn_l3_run": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
    "shn_l3_get_sizes": (C.c_int, [vp, C.POINTER(L3Sizes)]),
    "shn_l3_get_walks": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "shn_l3_get_contigs": (C.c_int, [vp, vp, vp]),
    "shn_l3_get_allowed": (C.c_int, [vp, vp, vp]),
    "shn_l3_get_edges": (C.c_int, [vp, vp, vp, vp, vp]),
    "shn_l3_get_labels": (C.c_int, [vp, vp]),
    "shn_l4_map_add_contigs": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_uint64]),
    "shn_l4_map_add_l3_contigs": (C.c_int, [vp, vp, C.c_uint64, C.c_int]),
    "shn_l4_map_set_weights": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "shn_l4_map_window_weights": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, vp]),
    "shn_l4_load_reads": (C.c_int, [vp, C.c_int, vp, vp, C.c_uint64, C.c_int]),
    "shn_l4_upload_reads_async": (C.c_int, [vp, C.c_int, vp, vp, C.c_uint64]),
    "shn_l4_load_reads_staged": (C.c_int, [vp, C.c_int]),
    "shn_l4_assign": (C.c_int, [vp, C.c_int, C.c_int, u64p, u64p, u64p]),
    "shn_l4_get_assignments": (C.c_int, [vp, C.c_uint32, vp, vp]),
    "shn_l4_assignments_dev": (C.c_int, [vp, C.c_uint32, C.c_uint64, vp, vp]),
    "shn_synth_pairs": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.c_int, C.c_int, C.c_uint32, vp, vp]),
    "shn_revcomp_reads": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int]),
    "shn_count_k1mers": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                   C.POINTER(vp), C.POINTER(vp), u64p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ShnError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShnError("%s not found: build it with `python -m shannon_b200.build` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def key_words(k1):
    """uint64 words per packed K1-mer at the C-ABI: 1 for k1 <= 32, 2 (low word first) for k1 = 33."""
    return 2 if k1 > 32 else 1


def shape_keys(flat, k1):
    """flat uint64 array from the library -> (n,) for one-word keys, (n, 2) for two-word keys."""
    return flat.reshape(-1, 2) if key_words(k1) == 2 else flat


def empty_keys(n, k1):
    return np.empty((n, 2) if key_words(k1) == 2 else (n,), dtype=np.uint64)


def ptr(a):
    """void* of a numpy array / int device pointer / None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(a)), C.c_void_p)
    raise TypeError(type(a))


class HostIO(object):
    """The native host-side text IO of the library; needs no GPU (ctx handle may be NULL)."""

    h = None

    def __init__(self):
        self.lib = load()

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise ShnError("%s: %s" % (name, self.lib.shn_last_error(self.h).decode()))

    # ---- host IO --------------------------------------------------------------------------
    def parse_kmer_file(self, path):
        keys, counts = vp(), vp()
        n, k1 = C.c_uint64(), C.c_int()
        self.call("shn_parse_kmer_file", os.fsencode(path), C.byref(keys), C.byref(counts),
                  C.byref(n), C.byref(k1))
        kw = key_words(k1.value)
        try:
            k = np.ctypeslib.as_array(C.cast(keys, u64p), shape=(max(n.value, 1) * kw,))[:n.value * kw].copy()
            k = shape_keys(k, k1.value)
            c = np.ctypeslib.as_array(C.cast(counts, u32p), shape=(max(n.value, 1),))[:n.value].copy()
        finally:
            self.lib.shn_host_free(keys)
            self.lib.shn_host_free(counts)
        return k, c, k1.value

    def load_fasta(self, path, n_fixed=-1):
        bases, offs = vp(), vp()
        n = C.c_uint64()
        self.call("shn_load_fasta", os.fsencode(path), C.c_int64(n_fixed), C.byref(bases),
                  C.byref(offs), C.byref(n))
        try:
            o = np.ctypeslib.as_array(C.cast(offs, u64p), shape=(n.value + 1,)).copy()
            nb = int(o[-1])
            b = np.ctypeslib.as_array(C.cast(bases, u8p), shape=(max(nb, 1),))[:nb].copy()
        finally:
            self.lib.shn_host_free(bases)
            self.lib.shn_host_free(offs)
        return b, o

    def load_fasta_named(self, path):
        """(names uint8, name_offsets, bases uint8, offsets) of a FASTA file read like rc_s.py does."""
        names, noffs, bases, offs = vp(), vp(), vp(), vp()
        n = C.c_uint64()
        self.call("shn_load_fasta_named", os.fsencode(path), C.byref(names), C.byref(noffs),
                  C.byref(bases), C.byref(offs), C.byref(n))
        try:
            no = np.ctypeslib.as_array(C.cast(noffs, u64p), shape=(n.value + 1,)).copy()
            o = np.ctypeslib.as_array(C.cast(offs, u64p), shape=(n.value + 1,)).copy()
            nn, nb = int(no[-1]), int(o[-1])
            nm = np.ctypeslib.as_array(C.cast(names, u8p), shape=(max(nn, 1),))[:nn].copy()
            b = np.ctypeslib.as_array(C.cast(bases, u8p), shape=(max(nb, 1),))[:nb].copy()
        finally:
            for x in (names, noffs, bases, offs):
                self.lib.shn_host_free(x)
        return nm, no, b, o

    def write_fasta_named(self, path, append, names, name_offsets, bases, offsets):
        self.call("shn_write_fasta_named", os.fsencode(path), int(bool(append)), ptr(names),
                  ptr(name_offsets), ptr(bases), ptr(offsets), C.c_uint64(len(offsets) - 1))

    def write_fasta_subset(self, path, append, bases, offsets, read_idx, first_index, suffix):
        read_idx = np.ascontiguousarray(read_idx, dtype=np.uint32)
        self.call("shn_write_fasta_subset", os.fsencode(path), int(bool(append)), ptr(bases),
                  ptr(offsets), ptr(read_idx), C.c_uint64(len(read_idx)), C.c_uint64(first_index),
                  suffix.encode())

    def write_kmer_file(self, path, keys, counts, k1):
        """k1mer.dict_org (`KMER\\tcount` lines) from packed keys."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self.call("shn_write_kmer_file", os.fsencode(path), ptr(keys), ptr(counts),
                  C.c_uint64(len(counts)), int(k1))

    def write_k1mer_windows(self, path, bases, offsets, contig_ids, k1, weights, win_off):
        contig_ids = np.ascontiguousarray(contig_ids, dtype=np.uint32)
        self.call("shn_write_k1mer_windows", os.fsencode(path), ptr(bases), ptr(offsets),
                  ptr(contig_ids), C.c_uint64(len(contig_ids)), int(k1), ptr(weights), ptr(win_off))


class Context(HostIO):
    """One GPU context (shn_ctx).  Not thread-safe; re-entrant across instances."""

    def __init__(self, device=0):
        self.lib = load()
        h = vp()
        rc = self.lib.shn_create(int(device), C.byref(h))
        if rc != 0:
            raise ShnError(self.lib.shn_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            for hp in getattr(self, "_pinned", []):
                self.lib.shn_host_free_pinned(self.h, vp(hp))
            self._pinned = []
            self.lib.shn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- memory / timing ----------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = vp()
        self.call("shn_dev_alloc", C.c_uint64(int(nbytes)), C.byref(p))
        return p.value or 0

    def dev_free(self, dptr):
        self.call("shn_dev_free", vp(dptr))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.call("shn_memcpy_h2d", vp(dptr), ptr(arr), C.c_uint64(arr.nbytes))

    def d2h(self, arr, dptr):
        self.call("shn_memcpy_d2h", ptr(arr), vp(dptr), C.c_uint64(arr.nbytes))
        return arr

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.dev_alloc(max(arr.nbytes, 1))
        if arr.nbytes:
            self.h2d(d, arr)
        return d

    def pinned_empty(self, n, dtype):
        """numpy array over page-locked host memory (falls back to pageable memory)."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 1)
        p = vp()
        try:
            self.call("shn_host_alloc_pinned", C.c_uint64(nbytes), C.byref(p))
        except ShnError:
            return np.empty(int(n), dtype=dtype)
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p.value)
        return arr

    def pinned_scratch(self, name, n, dtype):
        """View of a cached page-locked buffer (grown on demand, re-used by the next call with
        the same name: consume the contents before calling again)."""
        dtype = np.dtype(dtype)
        cache = self.__dict__.setdefault("_pinned_cache", {})
        buf = cache.get(name)
        need = max(int(n), 1) * dtype.itemsize
        if buf is None or buf.nbytes < need:
            buf = cache[name] = self.pinned_empty(need + need // 2, np.uint8)
        return buf[:int(n) * dtype.itemsize].view(dtype)

    def sync(self):
        self.call("shn_sync")

    def use_stream(self, cuda_stream):
        """cuda_stream: raw cudaStream_t as int (torch.cuda.current_stream().cuda_stream), 0/None =
        the context's own stream."""
        self.call("shn_use_stream", vp(cuda_stream or None))

    def d2d(self, d_dst, d_src, nbytes):
        self.call("shn_memcpy_d2d", vp(d_dst), vp(d_src), C.c_uint64(int(nbytes)))

    def timer_start(self):
        self.call("shn_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self.call("shn_timer_stop", C.byref(ms))
        return ms.value

    def prof_enable(self, on=True):
        self.call("shn_prof_enable", int(bool(on)))

    def prof(self):
        buf = C.create_string_buffer(1 << 16)
        self.call("shn_prof_dump", buf, C.c_uint64(len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split("\t")
            out[name] = (float(ms), int(n))
        return out

    def launch_count(self):
        return int(self.lib.shn_launch_count(self.h))

    def flush_l2(self):
        self.call("shn_flush_l2")

    def device_info(self):
        sm = C.c_int()
        fr, tot = C.c_uint64(), C.c_uint64()
        self.call("shn_device_info", C.byref(sm), C.byref(fr), C.byref(tot))
        return sm.value, fr.value, tot.value

    # ---- table ------------------------------------------------------------------------------
    def pack_kmers(self, ascii_bytes, n, k1):
        a = np.frombuffer(ascii_bytes, dtype=np.uint8) if not isinstance(ascii_bytes, np.ndarray) \
            else ascii_bytes
        keys = empty_keys(n, k1)
        self.call("shn_pack_kmers", ptr(np.ascontiguousarray(a)), C.c_uint64(n), int(k1), ptr(keys), 0)
        return keys

    def table_build(self, keys, counts, k1, double_stranded=False, on_device=False, n=None):
        if not on_device:
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            counts = np.ascontiguousarray(counts, dtype=np.uint32)
            n = len(keys)
        self.call("shn_table_build", ptr(keys), ptr(counts), C.c_uint64(n), int(k1),
                  int(bool(double_stranded)), int(bool(on_device)))

    def table_build_indexed(self, d_keys, d_counts, d_line_idx, n, k1):
        self.call("shn_table_build_indexed", vp(d_keys), vp(d_counts), vp(d_line_idx), C.c_uint64(n),
                  int(k1))

    def route_plan(self, d_keys, n, nranks, d_perm):
        counts = (C.c_uint64 * nranks)()
        self.call("shn_route_plan", vp(d_keys), C.c_uint64(n), C.c_uint32(nranks), vp(d_perm), counts)
        return [int(x) for x in counts]

    def permute(self, d_src, d_perm, n, elem_bytes, d_dst, scatter=False):
        self.call("shn_permute", vp(d_src), vp(d_perm), C.c_uint64(n), int(elem_bytes),
                  int(bool(scatter)), vp(d_dst))

    def table_stats(self):
        nd, nl, ns = C.c_uint64(), C.c_uint64(), C.c_uint64()
        k1 = C.c_int()
        self.call("shn_table_stats", C.byref(nd), C.byref(nl), C.byref(ns), C.byref(k1))
        return {"n_distinct": nd.value, "n_lowcomplexity": nl.value, "n_slots": ns.value,
                "k1": k1.value}

    def table_lookup(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        w = np.empty(len(keys), dtype=np.uint32)
        f = np.empty(len(keys), dtype=np.uint8)
        self.call("shn_table_lookup", ptr(keys), C.c_uint64(len(keys)), ptr(w), ptr(f), 0)
        return w, f

    def table_lookup_dev(self, d_keys, n, d_weights, d_found):
        self.call("shn_table_lookup", vp(d_keys), C.c_uint64(n), vp(d_weights), vp(d_found), 1)

    def table_dump(self):
        st = self.table_stats()
        n = st["n_distinct"]
        keys = empty_keys(n, st["k1"])
        w = np.empty(n, dtype=np.uint32)
        idx = np.empty(n, dtype=np.uint32)
        self.call("shn_table_dump", ptr(keys), ptr(w), ptr(idx))
        return keys, w, idx

    # ---- L3 -----------------------------------------------------------------------------------
    def l3_run(self, min_weight, min_length):
        self.call("shn_l3_run", C.c_uint32(min_weight), C.c_uint32(min_length))
        return self.l3_sizes()

    # the two phases of l3_run and the candidate exchange of the sharded path (device pointers)
    def l3_walks_phase(self, min_weight, min_length):
        self.call("shn_l3_walks", C.c_uint32(min_weight), C.c_uint32(min_length))

    def l3_cand_sizes(self):
        n, nb = C.c_uint64(), C.c_uint64()
        self.call("shn_l3_cand_sizes", C.byref(n), C.byref(nb))
        return n.value, nb.value

    def l3_cand_export(self, d_weight, d_line, d_offs, d_codes):
        self.call("shn_l3_cand_export", vp(d_weight), vp(d_line), vp(d_offs), vp(d_codes))

    def l3_filter_phase(self, d_codes=None, d_offs=None, n_cand=0, external=False, allow_missing=False):
        self.call("shn_l3_filter", vp(d_codes or None), vp(d_offs or None), C.c_uint64(int(n_cand)),
                  int(bool(external)), int(bool(allow_missing)))
        return self.l3_sizes()

    def l3_allowed_copy(self, d_keys, d_weights):
        self.call("shn_l3_allowed_copy", vp(d_keys or None), vp(d_weights or None))

    def l3_set_allowed_weights(self, d_weights):
        self.call("shn_l3_set_allowed_weights", vp(d_weights or None))

    # ---- sharded tables ------------------------------------------------------------------------
    def _route(self, name, nranks, counts, d_send, *args):
        arr = (C.c_uint64 * nranks)(*([0] * nranks if counts is None else [int(x) for x in counts]))
        self.call(name, *args, arr, vp(d_send or None))
        return [int(x) for x in arr]

    def route_lines(self, d_keys, d_counts, n, first_line, double_stranded, k1, nranks, counts=None,
                    d_send=None):
        return self._route("shn_route_lines", nranks, counts, d_send, vp(d_keys or None),
                           vp(d_counts or None), C.c_uint64(int(n)), C.c_uint64(int(first_line)),
                           int(bool(double_stranded)), int(k1), C.c_uint32(nranks))

    def table_build_records(self, d_recs, n, k1):
        self.call("shn_table_build_records", vp(d_recs or None), C.c_uint64(int(n)), int(k1))

    def trim(self):
        self.call("shn_trim")

    def cc_local(self):
        n = C.c_uint64()
        self.call("shn_cc_local", C.byref(n))
        return n.value

    def cc_cross(self, nranks, rank, gid_base, counts=None, d_send=None):
        return self._route("shn_cc_cross", nranks, counts, d_send, C.c_uint32(nranks), C.c_uint32(rank),
                           C.c_uint64(int(gid_base)))

    def cc_resolve(self, d_recs, n, gid_base, d_edges):
        ne = C.c_uint64()
        self.call("shn_cc_resolve", vp(d_recs or None), C.c_uint64(int(n)), C.c_uint64(int(gid_base)),
                  vp(d_edges or None), C.byref(ne))
        return ne.value

    def cc_merge(self, d_edges, n_edges, n_super):
        nf = C.c_uint64()
        self.call("shn_cc_merge", vp(d_edges or None), C.c_uint64(int(n_edges)), C.c_uint64(int(n_super)),
                  C.byref(nf))
        return nf.value

    def cc_sizes(self, gid_base, d_sizes):
        self.call("shn_cc_sizes", C.c_uint64(int(gid_base)), vp(d_sizes))

    def cc_route(self, d_owner_of_final, gid_base, nranks, counts=None, d_send=None):
        return self._route("shn_cc_route", nranks, counts, d_send, vp(d_owner_of_final),
                           C.c_uint64(int(gid_base)), C.c_uint32(nranks))

    def cc_free(self):
        self.call("shn_cc_free")

    def l3_sizes(self):
        s = L3Sizes()
        self.call("shn_l3_get_sizes", C.byref(s))
        return dict((f, getattr(s, f)) for f, _ in L3Sizes._fields_)

    def l3_walks(self):
        n = self.l3_sizes()["n_walks"]
        seed = empty_keys(n, self.table_stats()["k1"])
        nl = np.empty(n, dtype=np.uint32)
        nr = np.empty(n, dtype=np.uint32)
        tot = np.empty(n, dtype=np.uint64)
        flags = np.empty(n, dtype=np.uint8)
        self.call("shn_l3_get_walks", ptr(seed), ptr(nl), ptr(nr), ptr(tot), ptr(flags))
        return seed, nl, nr, tot, flags

    def l3_contigs(self):
        sz = self.l3_sizes()
        bases = np.empty(max(sz["contig_bases"], 1), dtype=np.uint8)
        offs = np.empty(sz["n_contigs"] + 1, dtype=np.uint64)
        self.call("shn_l3_get_contigs", ptr(bases), ptr(offs))
        return bases[:sz["contig_bases"]], offs

    def l3_allowed(self):
        n = self.l3_sizes()["n_allowed"]
        keys = empty_keys(n, self.table_stats()["k1"])
        w = np.empty(n, dtype=np.uint32)
        self.call("shn_l3_get_allowed", ptr(keys), ptr(w))
        return keys, w

    def l3_edges(self):
        n = self.l3_sizes()["n_edges"]
        a = np.empty(n, dtype=np.uint32)
        b = np.empty(n, dtype=np.uint32)
        w = np.empty(n, dtype=np.uint32)
        fp = np.empty(n, dtype=np.uint32)
        self.call("shn_l3_get_edges", ptr(a), ptr(b), ptr(w), ptr(fp))
        return a, b, w, fp

    def l3_labels(self):
        n = self.l3_sizes()["n_contigs"]
        lab = np.empty(n + 1, dtype=np.uint32)
        self.call("shn_l3_get_labels", ptr(lab))
        return lab

    # ---- L4 -----------------------------------------------------------------------------------
    def l4_map_add_contigs(self, bases, offsets, comp_of_contig, k1, reset, expected_total):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        comp = np.ascontiguousarray(comp_of_contig, dtype=np.uint32)
        self.call("shn_l4_map_add_contigs", ptr(bases), ptr(offsets), ptr(comp),
                  C.c_uint64(len(comp)), int(k1), int(bool(reset)), C.c_uint64(int(expected_total)))

    def l4_map_add_l3_contigs(self, comp_of_contig, reset=True):
        comp = np.ascontiguousarray(comp_of_contig, dtype=np.uint32)
        self.call("shn_l4_map_add_l3_contigs", ptr(comp), C.c_uint64(len(comp)), int(bool(reset)))

    def l4_map_set_weights(self, keys, weights):
        if keys is None:   # use the allowed set of this context's last l3_run, on the device
            self.call("shn_l4_map_set_weights", None, None, C.c_uint64(0))
            return
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        weights = np.ascontiguousarray(weights, dtype=np.uint32)
        self.call("shn_l4_map_set_weights", ptr(keys), ptr(weights), C.c_uint64(len(keys)))

    def l4_map_window_weights(self, bases, offsets, k1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lens = np.diff(offsets.astype(np.int64))
        nwin = np.maximum(lens - k1 + 1, 0)
        win_off = np.zeros(len(nwin) + 1, dtype=np.uint64)
        win_off[1:] = np.cumsum(nwin)
        w = np.empty(max(int(win_off[-1]), 1), dtype=np.uint32)
        self.call("shn_l4_map_window_weights", ptr(bases), ptr(offsets), C.c_uint64(len(nwin)),
                  int(k1), ptr(w))
        return w[:int(win_off[-1])], win_off

    def l4_load_reads(self, mate, bases, offsets, n=None, on_device=False):
        if not on_device:
            bases = np.ascontiguousarray(bases, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            n = len(offsets) - 1
        self.call("shn_l4_load_reads", int(mate), ptr(bases), ptr(offsets), C.c_uint64(n),
                  int(bool(on_device)))

    def l4_upload_reads_async(self, mate, bases, offsets):
        """bases/offsets: C-contiguous host arrays that stay alive until l4_load_reads_staged."""
        assert bases.dtype == np.uint8 and offsets.dtype == np.uint64
        self.call("shn_l4_upload_reads_async", int(mate), ptr(bases), ptr(offsets),
                  C.c_uint64(len(offsets) - 1))

    def l4_load_reads_staged(self, mate):
        self.call("shn_l4_load_reads_staged", int(mate))

    def l4_assign(self, paired, k1):
        na, nl, nv = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.call("shn_l4_assign", int(bool(paired)), int(k1), C.byref(na), C.byref(nl), C.byref(nv))
        return na.value, nl.value, nv.value

    def l4_assignments(self, n_comps, n_assign, pinned=False):
        """pinned=True: record_idx is a view of a cached page-locked buffer (valid until the next
        call), which makes the device-to-host copy of large partitions several times faster."""
        offs = np.empty(n_comps + 1, dtype=np.uint64)
        if pinned:
            idx = self.pinned_scratch("l4_assignments", max(n_assign, 1), np.uint32)
        else:
            idx = np.empty(max(n_assign, 1), dtype=np.uint32)
        self.call("shn_l4_get_assignments", C.c_uint32(n_comps), ptr(offs), ptr(idx))
        return offs, idx[:n_assign]

    def l4_assignments_dev(self, n_comps, first_record, d_offs, d_idx):
        self.call("shn_l4_assignments_dev", C.c_uint32(n_comps), C.c_uint64(int(first_record)),
                  vp(d_offs), vp(d_idx or None))

    # ---- inputs of the path -----------------------------------------------------------------
    def synth_pairs(self, d_tx, d_tx_offs, d_thr, n_tx, n_pairs, first_pair, seed, read_len,
                    frag_len, err_thr, d_m1, d_m2):
        self.call("shn_synth_pairs", vp(d_tx), vp(d_tx_offs), vp(d_thr), C.c_uint64(n_tx),
                  C.c_uint64(n_pairs), C.c_uint64(first_pair), C.c_uint64(seed), int(read_len),
                  int(frag_len), C.c_uint32(err_thr), vp(d_m1), vp(d_m2))

    def revcomp_reads(self, d_in, d_out, n_reads, read_len):
        self.call("shn_revcomp_reads", vp(d_in), vp(d_out), C.c_uint64(n_reads), int(read_len))

    def count_release(self):
        self.call("shn_count_release")

    def condense(self, prefix_kmers, suffix_kmers, prevalence, K):
        """unitig graph of the de Bruijn graph whose edges are the given K1-mer lines; returns a dict
        of numpy arrays (bases, offsets, count, prevalence, edge_src, edge_dst, edge_copy_count)."""
        pre = np.ascontiguousarray(prefix_kmers, dtype=np.uint64)
        suf = np.ascontiguousarray(suffix_kmers, dtype=np.uint64)
        prev = np.ascontiguousarray(prevalence, dtype=np.uint32)
        nu, nb, ne, ncyc = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.call("shn_condense_run", ptr(pre), ptr(suf), ptr(prev), C.c_uint64(len(prev)), int(K),
                  C.byref(nu), C.byref(nb), C.byref(ne), C.byref(ncyc))
        out = {"bases": np.empty(max(nb.value, 1), np.uint8), "offsets": np.zeros(nu.value + 1, np.uint64),
               "count": np.empty(max(nu.value, 1), np.uint32), "prevalence": np.empty(max(nu.value, 1), np.uint64),
               "edge_src": np.empty(max(ne.value, 1), np.uint32), "edge_dst": np.empty(max(ne.value, 1), np.uint32),
               "edge_copy_count": np.empty(max(ne.value, 1), np.uint32)}
        self.call("shn_condense_get", ptr(out["bases"]), ptr(out["offsets"]), ptr(out["count"]),
                  ptr(out["prevalence"]), ptr(out["edge_src"]), ptr(out["edge_dst"]), ptr(out["edge_copy_count"]))
        out["bases"] = out["bases"][:nb.value]
        for k in ("count", "prevalence"):
            out[k] = out[k][:nu.value]
        for k in ("edge_src", "edge_dst", "edge_copy_count"):
            out[k] = out[k][:ne.value]
        out["n_cycle_nodes"] = ncyc.value
        return out

    def find_reps(self, bases, offsets, name_rank, double_stranded):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        name_rank = np.ascontiguousarray(name_rank, dtype=np.uint32)
        dup = np.zeros(max(len(name_rank), 1), dtype=np.uint8)
        self.call("shn_find_reps", ptr(bases), ptr(offsets), ptr(name_rank), C.c_uint64(len(name_rank)),
                  int(bool(double_stranded)), ptr(dup))
        return dup[:len(name_rank)]

    def revcomp_var(self, bases, offsets):
        """reverse complement of every read of (bases uint8, offsets uint64), host arrays"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.empty(max(len(bases), 1), dtype=np.uint8)
        self.call("shn_revcomp_var", ptr(bases), ptr(offsets), C.c_uint64(len(offsets) - 1), ptr(out), 0)
        return out[:len(bases)]

    def revcomp_var_dev(self, d_bases, d_offsets, n_reads, d_out):
        self.call("shn_revcomp_var", vp(d_bases), vp(d_offsets), C.c_uint64(n_reads), vp(d_out), 1)

    def count_begin(self, k1, expected_distinct):
        self.call("shn_count_begin", int(k1), C.c_uint64(int(expected_distinct)))

    def count_add_reads(self, bases, offsets, n=None, on_device=False):
        if not on_device:
            bases = np.ascontiguousarray(bases, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            n = len(offsets) - 1
        self.call("shn_count_add_reads", ptr(bases), ptr(offsets), C.c_uint64(n), int(bool(on_device)))

    def count_finish(self, min_count=1):
        keys, counts = vp(), vp()
        n = C.c_uint64()
        self.call("shn_count_finish", C.c_uint32(min_count), C.byref(keys), C.byref(counts), C.byref(n))
        return keys.value or 0, counts.value or 0, n.value

    def count_k1mers(self, d_arrays, n_reads, read_len, k1, expected_distinct):
        na = len(d_arrays)
        arr = (vp * na)(*[vp(a) for a in d_arrays])
        nr = (C.c_uint64 * na)(*[int(x) for x in n_reads])
        keys, counts = vp(), vp()
        n = C.c_uint64()
        self.call("shn_count_k1mers", arr, nr, na, int(read_len), int(k1),
                  C.c_uint64(int(expected_distinct)), C.byref(keys), C.byref(counts), C.byref(n))
        return keys.value or 0, counts.value or 0, n.value
